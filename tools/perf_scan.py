"""Timing of the single-trait scan kernel per digit-plane count, with the merged and the unmerged digit epilogue
(GAPIT_EPI_PAIR, the test hook of the n >= 32,640 path); kernel_ms from the library's own CUDA events."""
import os, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import gapit_b200 as gb
from gapit_b200 import synth

n = int(os.environ.get("PERF_N", 5000)); m = int(os.environ.get("PERF_M", 200000))
G = synth.make_genotypes(n, m, seed=3, threads=16)
Y = synth.make_phenotypes(G[:4096])
X0 = np.ones((n, 1))
ctx = gb.Context(0)
geno = gb.Genotypes(G, ctx=ctx, marker_major=True)
K = gb.GAPIT_kinship_VanRaden(geno, ctx=ctx, verbose=False)
print("kinship kernel_ms", ctx.stats())
eL = gb.emma_eigen_L_wo_Z(K, ctx=ctx)
reml = gb.GAPIT_emma_REMLE(Y[:, 0], X0, K, ctx=ctx)
variants = [v.split(",") for v in os.environ.get("PERF_VARIANTS", "GAPIT_EPI_PAIR=1;GAPIT_EPI_PAIR=0").split(";")]
for S in [int(x) for x in os.environ.get("PERF_SLICES", "5,6").split(",")]:
    ctx.set_slices(S)
    es = gb.EigenSystem(eL["values"], eL["vectors"], ctx=ctx)
    for rep in range(3):
        for var in variants:
            for kv in var:
                k, v = kv.split("=")
                os.environ[k] = v
            t = []
            for _ in range(3):
                t0 = time.perf_counter()
                gb.p3d_scan(geno, X0, Y[:, :1], eigsys=es, delta=[reml["delta"]], vg=[reml["vg"]], ctx=ctx, pinned=True)
                wall = (time.perf_counter() - t0) * 1e3
                t.append((ctx.stats()["kernel_ms"], wall))
            km = min(x[0] for x in t)
            print(f"S={S} {var} kernel_ms={km:.3f} wall_ms={min(x[1] for x in t):.3f} int8 TOP/s={2.0*n*n*m*S/km/1e9:.0f}", flush=True)
    es.close()
for _ in range(3):
    t0 = time.perf_counter()
    gb.p3d_scan(geno, X0, Y[:, :1], glm=True, ctx=ctx, pinned=True)
    print("glm kernel_ms", ctx.stats()["kernel_ms"], "wall_ms", (time.perf_counter() - t0) * 1e3)
