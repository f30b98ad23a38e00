"""Kinship at scale (BASELINE configs[3] shape: 50,000 taxa): timing of the gram kernel + exact checksums."""
import os, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import gapit_b200 as gb
from gapit_b200 import synth, _lib

n = int(os.environ.get("PERF_N", 50000)); m = int(os.environ.get("PERF_M", 250000))
t0 = time.time()
G = synth.make_genotypes(n, m, seed=9, threads=min(32, os.cpu_count() or 8))
print(f"generated {n} x {m} in {time.time()-t0:.1f}s", flush=True)
ctx = gb.Context(0)
lib = ctx.lib
t0 = time.time()
geno = gb.Genotypes(G, ctx=ctx, marker_major=True)
print(f"pack (H2D + colstats) {time.time()-t0:.2f}s", flush=True)
ldg = lib.gapit_gram_ld(n)
G_t = torch.empty((ldg, ldg), dtype=torch.int32, device="cuda")
sums = torch.empty(3, dtype=torch.int64, device="cuda")
K_t = torch.empty((n, n), dtype=torch.float64, device="cuda")
for rep in range(3):
    t0 = time.time()
    _lib.check(lib.gapit_vanraden_gram(ctx.h, geno.h, G_t.data_ptr(), sums.data_ptr()))
    st = ctx.stats()
    t1 = time.time()
    _lib.check(lib.gapit_vanraden_finish(ctx.h, n, G_t.data_ptr(), sums.data_ptr(), K_t.data_ptr(), None, None))
    t2 = time.time()
    ops = float(n) * (n + 1) * m
    print(f"rep {rep}: transpose {st['pack_ms']:.1f} ms, gram {st['kernel_ms']:.1f} ms = {ops/st['kernel_ms']/1e9:.0f} TOP/s (algorithmic), "
          f"gram call {1e3*(t1-t0):.1f} ms, epilogue {1e3*(t2-t1):.1f} ms", flush=True)
cs, cq = geno.colsums()
c = cs.astype(np.int64)
tr = int(torch.diagonal(G_t[:n, :n]).to(torch.int64).sum().item())
tot = 2 * int(torch.tril(G_t[:n, :n]).sum(dtype=torch.int64).item()) - tr      # only the lower triangle is formed
print("sum G == sum c^2:", tot == int((c * c).sum()), " trace G == sum x^2:", tr == int(cq.astype(np.int64).sum()))
sym = bool(torch.equal(K_t[:4096, :4096], K_t[:4096, :4096].T))
sub = G[:, -300:].astype(np.float64)   # last taxa: exercises the last tile band (fp64 BLAS: exact below 2^53)
ref = (sub.T @ sub).astype(np.int64)
got = G_t[n - 300:n, n - 300:n].cpu().numpy()
print("K symmetric block:", sym, " last 300x300 block (lower) exact:", np.array_equal(np.tril(got), np.tril(ref)))
rs = K_t.sum(dim=1).abs().max().item()
print("max |K 1| =", rs, " K[0,0] =", K_t[0, 0].item())
