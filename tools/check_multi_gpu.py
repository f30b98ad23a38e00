"""Multi-GPU parity check (torchrun, one rank per GPU): the marker-sharded pipeline against the same pipeline on one
GPU, on the same inputs (integer cross-products bit-identical for any rank count, scan rows
identical wherever a marker sits).  Prints one JSON line {"ok": true, ...} on rank 0; exits non-zero on a mismatch.

    python -m torch.distributed.run --nproc-per-node 2 --master-addr 127.0.0.1 tools/check_multi_gpu.py
"""
import json, os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import torch.distributed as dist
import gapit_b200 as gb
from gapit_b200 import _lib, shard, synth

rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
dist.init_process_group("nccl", device_id=dev)
n, m = int(os.environ.get("CHK_N", 700)), int(os.environ.get("CHK_M", 6001))
G = synth.make_genotypes(n, m, seed=23, threads=4)          # every rank generates the same matrix ...
G[5, :] = 2
G[m // world, :] = G[m // world - 1, :]                      # a duplicated marker straddling the first shard boundary
lo, hi = shard.marker_shard(m, rank, world)                  # ... and owns one column block of it
Y = synth.make_phenotypes(G, ntraits=2, seed=23)
X0 = synth.make_covariates(G, npc=2)
ctx = gb.Context(local)
stream = torch.cuda.Stream(device=dev)
torch.cuda.set_stream(stream)
ctx.set_stream(stream.cuda_stream)
lib = ctx.lib
geno = gb.Genotypes(G[lo:hi], ctx=ctx, marker_major=True)
ldg = lib.gapit_gram_ld(n)
G_t = torch.empty((ldg, ldg), dtype=torch.int32, device=dev)
s_t = torch.empty(3, dtype=torch.int64, device=dev)
_lib.check(lib.gapit_vanraden_gram(ctx.h, geno.h, G_t.data_ptr(), s_t.data_ptr()))
dist.all_reduce(G_t)
dist.all_reduce(s_t)
K = np.empty((n, n), order="F")
Gi = np.empty((n, n), dtype=np.int32)
_lib.check(lib.gapit_vanraden_finish(ctx.h, n, G_t.data_ptr(), s_t.data_ptr(), None, K.ctypes.data, Gi.ctypes.data))
# ---- the same sums with the exchange fused into the Gram kernel (bulk tensor reduce-adds over NVLink), no collective ----
fused = {}
try:
    import ctypes
    import time
    import torch.distributed._symmetric_memory as symm_mem
    flat = symm_mem.empty(ldg * ldg + 8, dtype=torch.int32, device=dev)   # Gram buffer + 3 int64 mailbox for the scalar sums
    hdl = symm_mem.rendezvous(flat, dist.group.WORLD)
    G_sym = flat[:ldg * ldg].view(ldg, ldg)
    box = flat[ldg * ldg:].view(torch.int64)
    peers = (ctypes.c_void_p * world)(*[int(p) for p in hdl.buffer_ptrs])
    boxes = (ctypes.c_void_p * world)(*[int(p) + ldg * ldg * 4 for p in hdl.buffer_ptrs])
    s_f = torch.empty(3, dtype=torch.int64, device=dev)
    ref_t = torch.tril(G_t[:n, :n])
    # root = -1: reduce-scatter by block-row owner, then the finishing rank (every rank in turn here) gathers the rows
    # the others own; root = 0: reduce to one rank
    for root in (-1, 0):
        flat.zero_()
        torch.cuda.synchronize()
        hdl.barrier()
        t0 = time.perf_counter()
        _lib.check(lib.gapit_vanraden_gram_fused(ctx.h, geno.h, peers, world, rank, root, s_f.data_ptr(), boxes))
        torch.cuda.synchronize()
        hdl.barrier()
        torch.cuda.synchronize()
        dt = time.perf_counter() - t0
        same = bool(torch.equal(box[:3], s_t))          # every rank's mailbox holds the all-reduced scalar sums
        if root >= 0:
            full = hdl.get_buffer(root, (ldg * ldg + 8,), torch.int32)[:ldg * ldg].view(ldg, ldg).clone()
            torch.cuda.synchronize()
            same = same and bool(torch.equal(ref_t, torch.tril(full[:n, :n])))
            hdl.barrier()
        else:
            for gatherer in range(world):     # one rank at a time pulls the others' block-rows into its own buffer
                if rank == gatherer:
                    keep = G_sym.clone()
                    _lib.check(lib.gapit_vanraden_gram_gather(ctx.h, n, peers, world, rank))
                    torch.cuda.synchronize()
                    same = same and bool(torch.equal(ref_t, torch.tril(G_sym[:n, :n])))
                    G_sym.copy_(keep)         # leave only the owned rows for the next gatherer
                    torch.cuda.synchronize()
                hdl.barrier()
        flag = torch.tensor([1 if same else 0], device=dev)
        dist.all_reduce(flag, op=dist.ReduceOp.MIN)
        fused[f"root{root}"] = {"bit_identical": bool(flag.item()), "call_ms": dt * 1e3}
except Exception as exc:   # no peer access / symmetric memory on this box: the NCCL route stays available
    fused = {"unavailable": repr(exc)[:200]}
vec = torch.empty((n, n), dtype=torch.float64, device=dev)
val = torch.empty(n, dtype=torch.float64, device=dev)
dv = torch.empty((2, 2), dtype=torch.float64, device=dev)
if rank == 0:
    eL = gb.emma_eigen_L_wo_Z(K, ctx=ctx)
    r = [gb.GAPIT_emma_REMLE(Y[:, t], X0, K, eig_L=eL, ctx=ctx) for t in range(2)]
    vec.copy_(torch.from_numpy(np.ascontiguousarray(eL["vectors"].T)))
    val.copy_(torch.from_numpy(eL["values"]))
    dv.copy_(torch.tensor([[x["delta"] for x in r], [x["vg"] for x in r]], dtype=torch.float64))
dist.broadcast(vec, 0)
dist.broadcast(val, 0)
dist.broadcast(dv, 0)
vectors = np.asfortranarray(vec.cpu().numpy().T)
values = val.cpu().numpy()
delta, vg = dv.cpu().numpy()
es = gb.EigenSystem(values, vectors, ctx=ctx)
arrs, _ = gb.p3d_scan(geno, X0, Y, eigsys=es, delta=delta, vg=vg, ctx=ctx)
glm, _ = gb.p3d_scan(geno, X0, Y, glm=True, ctx=ctx)
got = {}
for name, a in (("mlm", arrs), ("glm", glm)):
    for key in ("effect", "se", "pvalue"):
        for t in range(2):
            got[(name, key, t)] = shard.gather_markers(np.nan_to_num(a[key][t], nan=-7.0), m, rank, world, dst=0, device=dev)
ok = True
report = {"n": n, "m": m, "world": world}
if rank == 0:
    full = gb.Genotypes(G, ctx=ctx, marker_major=True)
    K1, G1 = gb.GAPIT_kinship_VanRaden(full, ctx=ctx, return_gram=True, verbose=False)
    report["gram_bit_identical"] = bool(np.array_equal(G1, Gi))
    report["kinship_bit_identical"] = bool(np.array_equal(K1, K))
    a1, _ = gb.p3d_scan(full, X0, Y, eigsys=es, delta=delta, vg=vg, ctx=ctx)
    g1, _ = gb.p3d_scan(full, X0, Y, glm=True, ctx=ctx)
    worst = 0.0
    for name, a in (("mlm", a1), ("glm", g1)):
        for key in ("effect", "se", "pvalue"):
            for t in range(2):
                ref = np.nan_to_num(a[key][t], nan=-7.0)
                d = np.abs(got[(name, key, t)] - ref) / np.maximum(np.abs(ref), 1e-300)
                worst = max(worst, float(d.max()))
    report["scan_max_rel_diff"] = worst
    report["fused_exchange"] = fused
    ok = report["gram_bit_identical"] and report["kinship_bit_identical"] and worst <= 1e-13
    ok = ok and all(v["bit_identical"] for k, v in fused.items() if isinstance(v, dict))
    report["ok"] = bool(ok)
    print(json.dumps(report), flush=True)
dist.barrier()
dist.destroy_process_group()
sys.exit(0 if ok else 1)
