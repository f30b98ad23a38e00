"""Full-size runs of BASELINE.json configs[2..4] (C3, C4, C5) -- used by ``bench.py`` (strong-scaling legs) and
``tools/run_config.py``.  One process per GPU (``torch.distributed`` already initialised by the caller for N > 1);
genotypes are generated on the device block by block, so no 20-50 GB host matrix is needed.

    C3  20,000 taxa x 1,000,000 SNPs, 1 trait:  kinship + eigh + REML + MLM scan, markers sharded over the ranks
    C4  50,000 taxa x 1,000,000 SNPs:           VanRaden kinship only
    C5  10,000 taxa x 2,000,000 SNPs x 100 traits sharing one eigendecomposition (one rotation for all traits)

Markers are split into contiguous blocks; rank r owns blocks r, r+N, ...  Kinship: every block's cross-product kernel
reduce-scatters its tiles straight into the block-row owners' Gram buffers over NVLink (bulk tensor reduce-adds,
``gapit_vanraden_gram_fused`` with root = -1; buffers from torch symmetric memory), rank 0 gathers the rows it does not
own (``gapit_vanraden_gram_gather``) and runs the exact epilogue.  Rank 0 runs the eigendecomposition (cuSOLVER, timed
separately) and REML, V / lambda / delta / vg are broadcast, every rank scans its blocks, rank 0 gathers p-values.
"""
from __future__ import annotations

import ctypes
import time

import numpy as np

CONFIGS = {"c3": dict(n=20000, m=1_000_000, T=1, scan=True), "c4": dict(n=50000, m=1_000_000, T=0, scan=False),
           "c5": dict(n=10000, m=2_000_000, T=100, scan=True)}


def gen_block(n, mb, seed, dev, pop):
    """int8 (mb, n) on the device: 3 sub-populations, p_k ~ U(0.05, 0.5) with an F_ST = 0.05 spread, Binomial(2, p)."""
    import torch
    g = torch.Generator(device=dev)
    g.manual_seed(seed)
    out = torch.empty((mb, n), dtype=torch.int8, device=dev)
    step = 4096   # fixed sub-block: the first rows of a block do not depend on how many rows are asked for
    for k0 in range(0, mb, step):
        k1 = min(mb, k0 + step)
        p = torch.rand(k1 - k0, generator=g, device=dev) * 0.45 + 0.05
        sd = torch.sqrt(0.05 * p * (1 - p))
        pp = (p[None, :] + sd[None, :] * torch.randn(3, k1 - k0, generator=g, device=dev)).clamp_(0.01, 0.99)
        P = pp[pop].T.contiguous()                      # (block, n)
        x = (torch.rand(P.shape, generator=g, device=dev) < P).to(torch.int8)
        x += (torch.rand(P.shape, generator=g, device=dev) < P).to(torch.int8)
        out[k0:k1] = x
    return out


def run_config(name, ctx, dev, rank=0, world=1, n=None, m=None, T=None, blocks=8, slices=None, exchange="fused"):
    """Run one full-size configuration; returns the report dict (meaningful on rank 0).  ``ctx`` must already share the
    current torch stream (``ctx.set_stream``)."""
    import torch
    import torch.distributed as dist
    import gapit_b200 as gb
    from gapit_b200 import _lib

    cfg = dict(CONFIGS[name])
    n = n or cfg["n"]
    m = m or cfg["m"]
    T = cfg["T"] if T is None else T
    lib = ctx.lib
    nblocks = max(blocks, world)
    bounds = [(b * m) // nblocks for b in range(nblocks + 1)]
    mine = [b for b in range(nblocks) if b % world == rank]
    g0 = torch.Generator(device=dev)
    g0.manual_seed(7)
    pop = torch.randint(0, 3, (n,), generator=g0, device=dev)

    def sync():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
            torch.cuda.synchronize()

    T_ = {}
    # ---------------- kinship ----------------
    ldg = lib.gapit_gram_ld(n)
    s_acc = torch.zeros(3, dtype=torch.int64, device=dev)
    s_t = torch.empty(3, dtype=torch.int64, device=dev)
    hdl = None
    if world > 1 and exchange == "fused":
        try:
            import torch.distributed._symmetric_memory as symm_mem
            G_acc = symm_mem.empty((ldg, ldg), dtype=torch.int32, device=dev)
            hdl = symm_mem.rendezvous(G_acc, dist.group.WORLD)
            peers = (ctypes.c_void_p * world)(*[int(p) for p in hdl.buffer_ptrs])
        except Exception as exc:   # no peer access on this box
            print("symmetric memory unavailable, using ncclAllReduce:", repr(exc)[:200], flush=True)
            hdl = None
    if hdl is None:
        G_acc = torch.zeros((ldg, ldg), dtype=torch.int32, device=dev)
        peers = (ctypes.c_void_p * 1)(G_acc.data_ptr())
    fused = hdl is not None or world == 1
    G_t = None if fused else torch.empty((ldg, ldg), dtype=torch.int32, device=dev)
    gram_ms = 0.0
    keep = {}
    K_dev = torch.empty((n, n), dtype=torch.float64, device=dev)     # the result buffer exists before the clock starts
    sync()
    t_gen = 0.0
    t0 = time.perf_counter()
    if fused:
        G_acc.zero_()
        if hdl is not None:
            hdl.barrier()
    for b in mine:
        tg = time.perf_counter()
        blk = gen_block(n, bounds[b + 1] - bounds[b], 1000 + b, dev, pop)
        torch.cuda.synchronize()
        t_gen += time.perf_counter() - tg
        geno = gb.Genotypes.from_device(blk.data_ptr(), n, blk.shape[0], ctx=ctx)
        del blk
        if fused:   # blocks of one rank accumulate in place; across ranks the tiles go to their block-row owners
            _lib.check(lib.gapit_vanraden_gram_fused(ctx.h, geno.h, peers, world if hdl is not None else 1,
                                                     rank if hdl is not None else 0, -1 if hdl is not None else 0,
                                                     s_t.data_ptr()))
        else:
            _lib.check(lib.gapit_vanraden_gram(ctx.h, geno.h, G_t.data_ptr(), s_t.data_ptr()))
            G_acc += G_t
        gram_ms += ctx.stats()["kernel_ms"]
        s_acc += s_t
        if cfg["scan"] and len(mine) == 1:
            keep[b] = geno           # one block per rank: keep it resident for the scan
        else:
            geno.close()
    if world > 1:
        ta = time.perf_counter()
        if hdl is not None:
            hdl.barrier()
            if rank == 0:
                _lib.check(lib.gapit_vanraden_gram_gather(ctx.h, n, peers, world, 0))
                T_["gather_rows_ms"] = ctx.stats()["post_ms"]
        else:
            dist.all_reduce(G_acc)
        dist.all_reduce(s_acc)
        torch.cuda.synchronize()
        T_["exchange_tail_s"] = time.perf_counter() - ta
    T_["exchange"] = ("fused: tiles reduce-scattered over NVLink inside the Gram kernel, rank 0 gathers the rows"
                      if hdl is not None else ("nccl all-reduce" if world > 1 else "none"))
    del G_t
    tf = time.perf_counter()
    if rank == 0 or hdl is None:
        _lib.check(lib.gapit_vanraden_finish(ctx.h, n, G_acc.data_ptr(), s_acc.data_ptr(), K_dev.data_ptr(), None, None))
    torch.cuda.synchronize()
    T_["kinship_epilogue_s"] = time.perf_counter() - tf
    sync()
    T_["kinship_total_s"] = time.perf_counter() - t0 - t_gen
    T_["generate_s"] = t_gen
    gram_ms_t = torch.tensor([gram_ms], device=dev, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(gram_ms_t, op=dist.ReduceOp.MAX)
    T_["gram_kernel_s_max_rank"] = float(gram_ms_t[0]) * 1e-3
    kin_ops = float(n) * (n + 1) * m
    out = {"config": name, "n": n, "m": m, "traits": T, "n_gpus": world, "blocks": nblocks,
           "kinship_tops_kernel": kin_ops / T_["gram_kernel_s_max_rank"] / 1e12,
           "kinship_tops_end_to_end": kin_ops / T_["kinship_total_s"] / 1e12,
           "G_checksum": int(torch.tril(G_acc[:n, :n]).sum(dtype=torch.int64).item()) if rank == 0 else None,
           "K_trace": float(torch.trace(K_dev).item()) if rank == 0 else None}
    del G_acc
    if hdl is not None:
        del hdl
    if not cfg["scan"]:
        out["timings_s"] = T_
        out["end_to_end_s"] = T_["kinship_total_s"]
        del K_dev
        torch.cuda.empty_cache()
        return out
    # ---------------- spectra + REML on rank 0, broadcast (everything stays on the device) ----------------
    val_t = torch.empty(n, dtype=torch.float64, device=dev)
    dv_t = torch.empty((2, max(T, 1)), dtype=torch.float64, device=dev)
    # phenotypes: 10 QTN per trait drawn from block 0 (same on every rank), h2 = 0.5
    blk0 = gen_block(n, min(4096, bounds[1]), 1000, dev, pop).to(torch.float64)
    g1 = torch.Generator(device=dev)
    g1.manual_seed(99)
    Yt = torch.empty((T, n), dtype=torch.float64, device=dev)
    for t in range(T):
        q = torch.randint(0, blk0.shape[0], (10,), generator=g1, device=dev)
        gv = torch.randn(10, generator=g1, device=dev, dtype=torch.float64) @ blk0[q]
        Yt[t] = 10.0 + gv + torch.randn(n, generator=g1, device=dev, dtype=torch.float64) * gv.std()
    del blk0
    Y = np.asfortranarray(Yt.cpu().numpy().T)
    X0 = np.ones((n, 1))
    sync()
    if rank == 0:
        t1 = time.perf_counter()
        lam = gb.eigh_dev(K_dev.data_ptr(), n, ctx=ctx)      # K_dev now holds the eigenvectors (row r = vector r)
        T_["eigh_s"] = time.perf_counter() - t1
        val_t.copy_(torch.from_numpy(lam))
    sync()
    t1 = time.perf_counter()
    if world > 1:
        dist.broadcast(K_dev, 0)
        dist.broadcast(val_t, 0)
    sync()
    T_["broadcast_s"] = time.perf_counter() - t1
    t1 = time.perf_counter()
    es = gb.EigenSystem.from_device(val_t.cpu().numpy(), K_dev.data_ptr(), ctx=ctx)
    del K_dev
    T_["eigsys_slice_s"] = time.perf_counter() - t1
    if rank == 0:
        t1 = time.perf_counter()
        proj = es.project(np.column_stack([X0, Y]))          # V'[X0 | Y] on the device
        rs = gb.reml_from_projections_multi(es.values, proj[:, 1:], proj[:, :1])   # all traits, host threads
        dl, vg = np.array([r["delta"] for r in rs]), np.array([r["vg"] for r in rs])
        T_["reml_all_traits_s"] = time.perf_counter() - t1
        dv_t[0].copy_(torch.from_numpy(dl))
        dv_t[1].copy_(torch.from_numpy(vg))
    if world > 1:
        dist.broadcast(dv_t, 0)
    delta, vg = dv_t[0].cpu().numpy(), dv_t[1].cpu().numpy()
    # ---------------- the scan ----------------
    scan_kernel_ms = scan_d2h_ms = 0.0
    minp = 1.0
    # result arrays are page-locked once, before the clock starts (1.2 GB per 250k-marker block at 100 traits)
    sizes = {bounds[b + 1] - bounds[b] for b in mine}
    arrs = gb.alloc_scan_out(T, sizes.pop(), pinned=True) if len(sizes) == 1 else None
    if world > 1:   # first use of a collective sets up its NCCL channels (~1 s): not part of the scan
        w = torch.zeros(8, device=dev, dtype=torch.float64)
        dist.gather(w, [torch.empty_like(w) for _ in range(world)] if rank == 0 else None, dst=0)
    sync()
    t_gen = 0.0
    t1 = time.perf_counter()
    for b in mine:
        if b in keep:
            geno = keep[b]
        else:
            tg = time.perf_counter()
            blk = gen_block(n, bounds[b + 1] - bounds[b], 1000 + b, dev, pop)
            geno = gb.Genotypes.from_device(blk.data_ptr(), n, blk.shape[0], ctx=ctx)
            del blk
            torch.cuda.synchronize()
            t_gen += time.perf_counter() - tg
        arrs, _ = gb.p3d_scan(geno, X0, Y, eigsys=es, delta=delta, vg=vg, ctx=ctx, pinned=True, out=arrs)
        st = ctx.stats()
        scan_kernel_ms += st["kernel_ms"]
        scan_d2h_ms += st["d2h_ms"]
        geno.close()
    if world > 1:   # rank 0 gathers every shard's p-values of trait 0 (marker order = block order): device to device
        p = torch.from_numpy(np.ascontiguousarray(arrs["pvalue"][0])).to(dev, non_blocking=True)
        bufs = [torch.empty_like(p) for _ in range(world)] if rank == 0 else None
        dist.gather(p, bufs, dst=0)
    sync()
    T_["scan_total_s"] = time.perf_counter() - t1 - t_gen
    minp = float(np.nanmin(arrs["pvalue"]))      # of the last block scanned (a consumer's read, outside the clock)
    T_["generate_scan_blocks_s"] = t_gen
    sk = torch.tensor([scan_kernel_ms], device=dev, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(sk, op=dist.ReduceOp.MAX)
    T_["scan_kernels_s_max_rank"] = float(sk[0]) * 1e-3
    T_["scan_d2h_tail_s"] = scan_d2h_ms * 1e-3
    S = slices or ctx.slices
    # the rotation is counted ONCE whatever the number of traits (int8-native 2 n^2 m S)
    out.update({
        "snps_per_s": m / T_["scan_total_s"], "snp_trait_tests_per_s": m * max(T, 1) / T_["scan_total_s"],
        "scan_int8_tops_per_gpu_rotation_once": 2.0 * n * n * (m / world) * S / T_["scan_kernels_s_max_rank"] / 1e12,
        "min_p_last_block": minp, "delta_first": float(delta[0]),
        "end_to_end_s": T_["kinship_total_s"] + T_.get("eigh_s", 0.0) + T_.get("reml_all_traits_s", 0.0) +
                        T_["broadcast_s"] + T_["eigsys_slice_s"] + T_["scan_total_s"],
        "timings_s": T_})
    es.close()
    del arrs
    torch.cuda.empty_cache()
    return out
