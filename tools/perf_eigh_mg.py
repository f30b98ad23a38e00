"""cuSOLVER multi-GPU syevd (gapit_eigh_mg -> libgapitcuda_mg.so) against the one-GPU syevd (gapit_eigh) on a VanRaden
kinship.  Runs WITHOUT torch in the process: libcusolverMg needs the CUDA toolkit's own libcublas (see the header).
    python tools/perf_eigh_mg.py [n] [ngpus] [m]"""
import ctypes as C, os, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import gapit_b200 as gb
from gapit_b200 import _lib, synth

n = int(sys.argv[1]) if len(sys.argv) > 1 else 8000
ngpus = int(sys.argv[2]) if len(sys.argv) > 2 else 0
m = int(sys.argv[3]) if len(sys.argv) > 3 else 2 * n
assert "torch" not in sys.modules
ctx = gb.Context(0)
G = synth.make_genotypes(n, m, seed=11, threads=min(32, os.cpu_count() or 8))
K = gb.GAPIT_kinship_VanRaden(gb.Genotypes(G, ctx=ctx, marker_major=True), ctx=ctx, verbose=False)
del G
P = lambda a: a.ctypes.data_as(C.c_void_p)
out = {}
for name, k in (("one GPU (cusolverDnXsyevd)", 1), (f"{ngpus or 'all'} GPUs (cusolverMgSyevd)", ngpus)):
    vals, vecs = np.empty(n), np.empty((n, n), order="F")
    t0 = time.perf_counter()
    if k == 1:
        _lib.check(ctx.lib.gapit_eigh(ctx.h, P(K), n, P(vals), P(vecs)))
    else:
        _lib.check(ctx.lib.gapit_eigh_mg(ctx.h, P(K), n, P(vals), P(vecs), k))
    dt = time.perf_counter() - t0
    idx = np.linspace(0, n - 1, 9).astype(int)           # residual on a few eigenpairs (a full check is n^3 on the host)
    res = max(np.abs(K @ vecs[:, i] - vals[i] * vecs[:, i]).max() for i in idx)
    orth = np.abs(vecs[:, idx].T @ vecs[:, idx] - np.eye(len(idx))).max()
    print(f"n={n} {name}: {dt:.3f} s incl. the two n^2 host copies, |K v - l v|max {res:.2e}, |V'V - I|max {orth:.2e}, "
          f"lambda[0] {vals[0]:.6f}, lambda[-1] {vals[-1]:.3e}", flush=True)
    out[k == 1] = vals
print("eigenvalues agree to", float(np.max(np.abs(out[True] - out[False]))))
