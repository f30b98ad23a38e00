"""A/B timing of the Gram kernel under environment switches (kernel_ms from the library's own CUDA events)."""
import os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import gapit_b200 as gb
from gapit_b200 import _lib

n = int(os.environ.get("PERF_N", 5000)); m = int(os.environ.get("PERF_M", 500000))
g0 = torch.Generator(device="cuda"); g0.manual_seed(1)
blk = (torch.rand((m, n), generator=g0, device="cuda") < 0.3).to(torch.int8) + (torch.rand((m, n), generator=g0, device="cuda") < 0.3).to(torch.int8)
torch.cuda.synchronize()
ctx = gb.Context(0)
geno = gb.Genotypes.from_device(blk.data_ptr(), n, m, ctx=ctx)
del blk
ldg = ctx.lib.gapit_gram_ld(n)
G_t = torch.empty((ldg, ldg), dtype=torch.int32, device="cuda")
s_t = torch.empty(3, dtype=torch.int64, device="cuda")
variants = [v.split(",") for v in os.environ.get("PERF_VARIANTS", "GAPIT_ROUND_WAIT=0;GAPIT_ROUND_WAIT=100000").split(";")]
ref = None
for rep in range(3):
    for var in variants:
        for kv in var:
            k, v = kv.split("=")
            os.environ[k] = v
        t = []
        for _ in range(3):
            _lib.check(ctx.lib.gapit_vanraden_gram(ctx.h, geno.h, G_t.data_ptr(), s_t.data_ptr()))
            t.append(ctx.stats()["kernel_ms"])
        cs = int(G_t.sum(dtype=torch.int64).item())
        ref = cs if ref is None else ref
        print(f"{var} gram kernel_ms={min(t):.3f} TOP/s={float(n)*(n+1)*m/min(t)/1e9:.0f} checksum_same={cs == ref}", flush=True)
