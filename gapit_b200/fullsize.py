"""Full-size runs of BASELINE.json configs[2..4] (C3, C4, C5) -- used by ``bench.py`` (strong-scaling legs) and
``tools/run_config.py``.  One process per GPU (``torch.distributed`` already initialised by the caller for N > 1);
genotypes are generated on the device block by block, so no 20-50 GB host matrix is needed.

    C3  20,000 taxa x 1,000,000 SNPs, 1 trait:  kinship + eigh + REML + MLM scan, markers sharded over the ranks
    C4  50,000 taxa x 1,000,000 SNPs:           VanRaden kinship only
    C5  10,000 taxa x 2,000,000 SNPs x 100 traits sharing one eigendecomposition (one rotation for all traits)

Markers are split into contiguous blocks; rank r owns blocks r, r+N, ... and holds them in ONE resident handle, generated
straight into the library's device layout (zero-copy `gapit_geno_wrap_device`).  Kinship: every rank's cross-product kernel
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
    # ---------------- genotypes: every block of this rank, generated straight into the library's device layout ----------
    # (one resident handle per rank: one cross-product launch and one scan call, however many blocks the rank owns;
    # block b is generated from its own seed, so the data -- and every checksum -- do not depend on the rank count)
    t_g = time.perf_counter()
    ldn = (n + 127) // 128 * 128
    m_rank = sum(bounds[b + 1] - bounds[b] for b in mine)
    big = torch.zeros((m_rank, ldn), dtype=torch.int8, device=dev)
    off = 0
    for b in mine:
        mb = bounds[b + 1] - bounds[b]
        big[off:off + mb, :n] = gen_block(n, mb, 1000 + b, dev, pop)
        off += mb
    geno = gb.Genotypes.wrap_device(big.data_ptr(), n, m_rank, ctx=ctx, keepalive=big)
    torch.cuda.synchronize()
    T_["generate_s"] = time.perf_counter() - t_g
    # ---------------- kinship ----------------
    ldg = lib.gapit_gram_ld(n)
    s_t = torch.zeros(3, dtype=torch.int64, device=dev)
    hdl = None
    if world > 1 and exchange == "fused":
        try:
            import torch.distributed._symmetric_memory as symm_mem
            flat = symm_mem.empty(ldg * ldg + 8, dtype=torch.int32, device=dev)   # Gram buffer + mailbox of the 3 scalar sums
            hdl = symm_mem.rendezvous(flat, dist.group.WORLD)
            G_acc = flat[:ldg * ldg].view(ldg, ldg)
            box = flat[ldg * ldg:].view(torch.int64)
            peers = (ctypes.c_void_p * world)(*[int(p) for p in hdl.buffer_ptrs])
            boxes = (ctypes.c_void_p * world)(*[int(p) + ldg * ldg * 4 for p in hdl.buffer_ptrs])
        except Exception as exc:   # no peer access on this box
            print("symmetric memory unavailable, using ncclAllReduce:", repr(exc)[:200], flush=True)
            hdl = None
    if hdl is None:
        G_acc = torch.zeros((ldg, ldg), dtype=torch.int32, device=dev)
    K_dev = torch.empty((n, n), dtype=torch.float64, device=dev)     # the result buffer exists before the clock starts
    sync()
    t0 = time.perf_counter()
    if hdl is not None:   # tiles go to their block-row owners inside the kernel, rank 0 gathers the rows
        flat.zero_()
        hdl.barrier()
        _lib.check(lib.gapit_vanraden_gram_fused(ctx.h, geno.h, peers, world, rank, -1, s_t.data_ptr(), boxes))
        gram_ms = ctx.stats()["kernel_ms"]
        ta = time.perf_counter()
        hdl.barrier()
        s_t = box[:3]                 # the scalar sums of all ranks, added into every rank's mailbox by the same call
        if rank == 0:
            _lib.check(lib.gapit_vanraden_gram_gather(ctx.h, n, peers, world, 0))
            T_["gather_rows_ms"] = ctx.stats()["post_ms"]
        torch.cuda.synchronize()
        T_["exchange_tail_s"] = time.perf_counter() - ta
    else:
        _lib.check(lib.gapit_vanraden_gram(ctx.h, geno.h, G_acc.data_ptr(), s_t.data_ptr()))
        gram_ms = ctx.stats()["kernel_ms"]
        if world > 1:
            ta = time.perf_counter()
            dist.all_reduce(G_acc)
            dist.all_reduce(s_t)
            torch.cuda.synchronize()
            T_["exchange_tail_s"] = time.perf_counter() - ta
    T_["exchange"] = ("fused: tiles reduce-scattered over NVLink inside the Gram kernel, rank 0 gathers the rows"
                      if hdl is not None else ("nccl all-reduce" if world > 1 else "none"))
    tf = time.perf_counter()
    if rank == 0 or hdl is None:
        _lib.check(lib.gapit_vanraden_finish(ctx.h, n, G_acc.data_ptr(), s_t.data_ptr(), K_dev.data_ptr(), None, None))
    torch.cuda.synchronize()
    T_["kinship_epilogue_s"] = time.perf_counter() - tf
    sync()
    T_["kinship_total_s"] = time.perf_counter() - t0
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
        geno.close()
        del K_dev, big
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
    sliced = world > 1 and n >= 2048 * world
    if sliced:
        # every rank solves for its own slice of the spectrum (cuSOLVER syevdx by index range): the tridiagonalisation
        # -- 87 % of syevd at n = 20,000 and bound by HBM, profiles/r2_syevd_breakdown.log -- is repeated on every
        # GPU, the divide-and-conquer back-transformation and the eigenvector traffic are split
        t1 = time.perf_counter()
        dist.broadcast(K_dev, 0)
        lo_r = [n * r // world for r in range(world + 1)]
        V_t = torch.empty((n, n), dtype=torch.float64, device=dev)      # row r = eigenvector r (descending)
        lam = gb.eigh_dev_range(K_dev.data_ptr(), n, lo_r[rank], lo_r[rank + 1] - lo_r[rank],
                                V_t[lo_r[rank]:lo_r[rank + 1]].data_ptr(), ctx=ctx)
        val_t[lo_r[rank]:lo_r[rank + 1]].copy_(torch.from_numpy(lam))
        torch.cuda.synchronize()
        mine_s = torch.tensor([time.perf_counter() - t1], device=dev, dtype=torch.float64)
        all_s = [torch.empty_like(mine_s) for _ in range(world)]
        dist.all_gather(all_s, mine_s)
        sync()
        T_["eigh_s"] = time.perf_counter() - t1
        T_["eigh_s_per_rank"] = [round(float(x[0]), 3) for x in all_s]    # the slowest rank sets eigh_s
        T_["eigh_route"] = f"spectrum sliced over {world} GPUs (K broadcast + syevdx per rank)"
        t1 = time.perf_counter()
        for r in range(world):
            dist.broadcast(V_t[lo_r[r]:lo_r[r + 1]], r)
            dist.broadcast(val_t[lo_r[r]:lo_r[r + 1]], r)
        K_dev = V_t
        del V_t
    else:
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
    if rank == 0:   # 64 eigenvectors spread over the spectrum (and over the ranks' slices): orthonormal?
        pick = K_dev[torch.linspace(0, n - 1, 64, device=dev).long()]
        out["eigvec_orthonormality_err_sample"] = float((pick @ pick.T - torch.eye(64, dtype=torch.float64, device=dev)).abs().max())
        del pick
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
    # ---------------- the scan: one call per rank over its resident markers ----------------
    # result arrays are page-locked once, before the clock starts (9.6 GB for 100 traits x 2,000,000 markers)
    arrs = gb.alloc_scan_out(T, m_rank, pinned=True)
    if world > 1:   # first use of a collective sets up its NCCL channels (~1 s): not part of the scan
        w = torch.zeros(8, device=dev, dtype=torch.float64)
        dist.gather(w, [torch.empty_like(w) for _ in range(world)] if rank == 0 else None, dst=0)
    sync()
    t1 = time.perf_counter()
    arrs, _ = gb.p3d_scan(geno, X0, Y, eigsys=es, delta=delta, vg=vg, ctx=ctx, pinned=True, out=arrs)
    st = ctx.stats()
    scan_kernel_ms, scan_d2h_ms = st["kernel_ms"] + st["post_ms"], st["d2h_ms"]
    if world > 1:   # rank 0 gathers every shard's p-values of trait 0 (marker order = rank order within a block round)
        mx = max((bounds[b + 1] - bounds[b]) for b in range(nblocks)) * ((nblocks + world - 1) // world)
        p = torch.zeros(mx, dtype=torch.float64, device=dev)
        p[:m_rank].copy_(torch.from_numpy(arrs["pvalue"][0]), non_blocking=True)
        bufs = [torch.empty_like(p) for _ in range(world)] if rank == 0 else None
        dist.gather(p, bufs, dst=0)
    sync()
    T_["scan_total_s"] = time.perf_counter() - t1
    minp = float(np.nanmin(arrs["pvalue"][0]))    # a consumer's read, outside the clock
    geno.close()
    del big
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
        "min_p_trait0_this_rank": minp, "delta_first": float(delta[0]),
        "end_to_end_s": T_["kinship_total_s"] + T_.get("eigh_s", 0.0) + T_.get("reml_all_traits_s", 0.0) +
                        T_["broadcast_s"] + T_["eigsys_slice_s"] + T_["scan_total_s"],
        "timings_s": T_})
    es.close()
    del arrs
    torch.cuda.empty_cache()
    return out
