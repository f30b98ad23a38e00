"""Multi-trait scan (BASELINE configs[4] in miniature): T traits share one eigendecomposition and ONE rotation
(store epilogue + fp64 DMMA contraction)."""
import os, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import gapit_b200 as gb
from gapit_b200 import synth

n = int(os.environ.get("PERF_N", 5000)); m = int(os.environ.get("PERF_M", 100000)); T = int(os.environ.get("PERF_T", 100))
G = synth.make_genotypes(n, m, seed=3, threads=16)
Y = synth.make_phenotypes(G[:4096], ntraits=T)
X0 = np.ones((n, 1))
ctx = gb.Context(0)
geno = gb.Genotypes(G, ctx=ctx, marker_major=True)
K = gb.GAPIT_kinship_VanRaden(geno, ctx=ctx, verbose=False)
eL = gb.emma_eigen_L_wo_Z(K, ctx=ctx)
eR = gb.emma_eigen_R_wo_Z(K, X0, ctx=ctx)
t0 = time.perf_counter()
remls = [gb.GAPIT_emma_REMLE(Y[:, t], X0, K, eig_R=eR, ctx=ctx) for t in range(T)]
print(f"REML for {T} traits: {1e3*(time.perf_counter()-t0):.1f} ms")
es = gb.EigenSystem(eL["values"], eL["vectors"], ctx=ctx)
d = [r["delta"] for r in remls]; v = [r["vg"] for r in remls]
for rep in range(3):
    t0 = time.perf_counter()
    arrs, _ = gb.p3d_scan(geno, X0, Y, eigsys=es, delta=d, vg=v, ctx=ctx, pinned=True)
    wall = (time.perf_counter() - t0) * 1e3
    st = ctx.stats()
    print(f"T={T}: kernel {st['kernel_ms']:.1f} ms, finalize {st['post_ms']:.1f} ms, d2h {st['d2h_ms']:.1f} ms, wall {wall:.1f} ms "
          f"=> {m*T/wall/1e3:.2f} M SNP x trait tests/s", flush=True)
t0 = time.perf_counter()
one, _ = gb.p3d_scan(geno, X0, Y[:, :1], eigsys=es, delta=d[:1], vg=v[:1], ctx=ctx, pinned=True)
print(f"T=1: kernel {ctx.stats()['kernel_ms']:.1f} ms, wall {1e3*(time.perf_counter()-t0):.1f} ms")
print("trait 0 equal to single-trait run (rtol 1e-8):", np.allclose(arrs['pvalue'][0], one['pvalue'][0], rtol=1e-8, equal_nan=True))
