"""A/B timing of the in-place device eigendecomposition (gapit_eigh_dev): 64-bit cusolverDnXsyevd vs legacy Dsyevd."""
import os, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import gapit_b200 as gb
ctx = gb.Context(0)
for n in [int(x) for x in os.environ.get("EIG_N", "5000,20000").split(",")]:
    g = torch.Generator(device="cuda"); g.manual_seed(n)
    X = torch.randn((n, 2 * n), generator=g, device="cuda", dtype=torch.float32)
    K0 = (X @ X.T).to(torch.float64) / (2 * n)
    del X
    for mode in ("x64", "legacy", "x64", "legacy"):
        os.environ["GAPIT_SYEVD"] = mode
        K = K0.clone()
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        vals = gb.eigh_dev(K.data_ptr(), n, ctx=ctx)
        dt = time.perf_counter() - t0
        V = K.T            # column k = eigenvector k
        res = (K0 @ V[:, :4] - V[:, :4] * torch.from_numpy(vals[:4]).cuda()).abs().max().item()
        print(f"n={n} {mode}: {dt*1e3:.1f} ms, lambda_max={vals[0]:.6f}, residual(top 4)={res:.2e}", flush=True)
        del K
