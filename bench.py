#!/usr/bin/env python
"""Benchmark of the GAPIT hot path on B200: SNPs tested/sec (MLM P3D) and kinship TOPS.

    python bench.py --gpus N --steps K --warmup W            # this repo (libgapitcuda)
    python bench.py --impl reference --gpus N --steps K ...  # CPU reference arm (oracle restatement)

Headline workload (BASELINE.json configs[1]): 5,000 taxa x 500,000 SNPs per GPU, one trait, intercept-only covariates,
VanRaden kinship + MLM/P3D scan (GLM scan reported beside it).  A "step" is one P3D scan of every marker resident on
the GPU.  Under torchrun each rank holds its own 500,000-marker shard (weak scaling); the kinship partial
cross-products are exchanged inside the Gram kernel over NVLink (reduce-scatter + one row gather), the eigen-system is
broadcast once, the scan needs no collective and rank 0 gathers the p-values.
`e2e` is the same scan through the public host-buffer call (`gapit_p3d_scan_host`) from pinned host memory in the
library's packed host format (2 bits per genotype call), results back to the host every step; `e2e.int8` is the same
call fed one byte per call.
Beside the headline the same line carries: `parity` (the oracle run on the very markers the GPU scanned, in this run),
the other BASELINE configs as `legs` (N = 1: configs[4], 100 traits; N >= 2: configs[2] and configs[3], STRONG scaling:
total work fixed, sharded over the N ranks), and the int8 / fp64 tensor ceilings calibrated on this box.
One JSON line is printed by rank 0 (see the task contract for the keys).
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "SNPs tested/sec (MLM P3D)"
UNIT = "SNPs/s"


def workload_name(n, m):
    """The `config.workload` string -- identical in both arms."""
    return f"MLM P3D scan, {n} taxa x {m} SNPs per GPU, 1 trait, q0=1 (BASELINE configs[1] per GPU)"


def ncu_traffic(n, m, S):
    """DRAM bytes (read + write) of one rotate_reduce launch from the committed `ncu --set full` capture of this exact
    shape (profiles/traffic.json, written by tools/ncu_traffic.py from the capture it names); None for other shapes."""
    try:
        with open(os.path.join(ROOT, "profiles", "traffic.json")) as f:
            for e in json.load(f)["rotate_reduce"]:
                if (e["n"], e["m"], e["slices"]) == (n, m, S):
                    return e["dram_bytes"], e["source"]
    except Exception:
        pass
    return None, None


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="native", choices=["native", "reference"])
    ap.add_argument("--n", type=int, default=5000, help="taxa (default: BASELINE configs[1])")
    ap.add_argument("--m", type=int, default=500000, help="markers per GPU")
    ap.add_argument("--slices", type=int, default=0, help="digit planes of the rotation (0 = the library default)")
    ap.add_argument("--cpu-sample", type=int, default=0, help="markers of the CPU baseline sample (0 = auto)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--kinship-exchange", default="fused", choices=["fused", "nccl"],
                    help="N > 1: exchange of the Gram partials fused into the kernel over NVLink peer memory, or ncclAllReduce")
    ap.add_argument("--no-legs", action="store_true", help="skip the other BASELINE configs (configs[2..4])")
    ap.add_argument("--no-parity", action="store_true")
    ap.add_argument("--reml-route", default="eigL", choices=["eigL", "eigR"],
                    help="eigL: REML from the spectrum of K alone (one eigendecomposition); eigR: the reference's "
                         "second decomposition of S(K+I)S (emma.txt:82-89)")
    return ap.parse_args()


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        with open(p) as f:
            d = json.load(f)
        return {"hbm_gbs": d["hbm_gbs"], "bf16_tflops": d["bf16_tflops"],
                "bf16_tflops_sustained": d.get("bf16_tflops_sustained", d["bf16_tflops"]), "source": "measured"}
    return {"hbm_gbs": 6650.0, "bf16_tflops": 1590.0, "bf16_tflops_sustained": 1400.0, "source": "fallback"}


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled during the timed region."""

    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index = index
        self.lines = []
        self.proc = None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                          "-lms", "50", "-i", str(self.index)], stdout=subprocess.PIPE, text=True)
            self.th = threading.Thread(target=self._read, daemon=True)
            self.th.start()
        except Exception:
            self.proc = None

    def _read(self):
        for ln in self.proc.stdout:
            self.lines.append(ln.strip())

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        for ln in self.lines:
            f = [x.strip() for x in ln.split(",")]
            if len(f) < 9:
                continue
            try:
                sm.append(float(f[1]))
                mx.append(float(f[2]))
            except ValueError:
                continue
            for name, val in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), f[5:9]):
                if val.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


def calibrate_int8_tops(dev):
    """cuBLASLt int8 GEMM (torch._int_mm, 8192^3) on this box: the measured int8 tensor ceiling."""
    import torch
    try:
        a = torch.randint(-8, 8, (8192, 8192), dtype=torch.int8, device=dev)
        b = torch.randint(-8, 8, (8192, 8192), dtype=torch.int8, device=dev)
        for _ in range(3):
            torch._int_mm(a, b)
        torch.cuda.synchronize()
        best = 1e9
        for _ in range(10):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            torch._int_mm(a, b)
            e1.record()
            torch.cuda.synchronize()
            best = min(best, e0.elapsed_time(e1))
        return 2.0 * 8192.0 ** 3 / (best * 1e-3) / 1e12
    except Exception:
        return None


def calibrate_fp64_tflops(dev):
    """cuBLAS DGEMM (torch.matmul float64, 8192^3) on this box: the measured fp64 ceiling the rotation would have as
    a plain fp64 GEMM."""
    import torch
    try:
        a = torch.randn((8192, 8192), dtype=torch.float64, device=dev)
        b = torch.randn((8192, 8192), dtype=torch.float64, device=dev)
        for _ in range(2):
            a @ b
        torch.cuda.synchronize()
        best = 1e9
        for _ in range(5):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            a @ b
            e1.record()
            torch.cuda.synchronize()
            best = min(best, e0.elapsed_time(e1))
        return 2.0 * 8192.0 ** 3 / (best * 1e-3) / 1e12
    except Exception:
        return None


def count_launches(fn):
    """Kernels of libgapitcuda launched by one call of `fn`, counted by CUPTI through torch.profiler (names in
    namespace gapit::).  Returns (count, {kernel name: launches}) or (None, reason)."""
    import torch
    try:
        from torch.profiler import profile, ProfilerActivity
        with profile(activities=[ProfilerActivity.CUDA]) as prof:
            fn()
            torch.cuda.synchronize()
        names = {}
        for ev in prof.events():
            nm = ev.name
            if "gapit::" in nm or nm.startswith("gapit"):
                key = nm.split("(")[0][:96]
                names[key] = names.get(key, 0) + 1
        return sum(names.values()), names
    except Exception as exc:
        return None, repr(exc)[:160]


def cpu_threads():
    try:
        from threadpoolctl import threadpool_info
        return max([p.get("num_threads", 1) for p in threadpool_info()] or [1])
    except Exception:
        return os.cpu_count() or 1


def cpu_scan_sample(n, ns, seed, eig_L, reml, G=None, Y=None, X0=None, keep=None):
    """Time the oracle's per-marker loop (two n x n GEMVs per marker, GAPIT.EMMAxP3D.R:634,957)
    on `ns` markers.  Returns SNPs/s; the oracle's rows are left in keep["ora"] for the in-run parity check."""
    from oracle import gapit_oracle as go
    t0 = time.perf_counter()
    ora = go.emmax_p3d(Y, G[:ns].T, K=np.zeros((2, 2)), X0=X0, eig_L=eig_L, reml=reml)
    dt = time.perf_counter() - t0
    if keep is not None:
        keep["ora"] = ora
    return ns / dt, dt


def parity_report(arrs, G_host, Y, X0, eig_L, reml_d, ora, ns_loop, block=2048):
    """Correctness on the same inputs in the same run: (1) the oracle's per-marker loop (the restated reference path)
    on the first `ns_loop` markers, (2) the batched fp64 derivation (oracle/batched.py: V'X by dgemm + closed-form
    bordered solve, independent of the loop and of the int8 split) on three blocks -- first, middle, last -- of the
    shard, against the rows the GPU produced for those markers."""
    from oracle import batched
    m = G_host.shape[0]
    rep = {}
    if ora is not None:
        ref = {"effect": ora.effect_est, "se": ora.se_clean, "tvalue": ora.stats, "pvalue": ora.ps}
        with np.errstate(divide="ignore"):
            ref["log10p"] = np.log10(ora.ps)
        one = {k: arrs[k][0] for k in ("effect", "se", "tvalue", "pvalue")}
        rep["per_marker_loop"] = batched.compare(one, ref, rows=slice(0, ns_loop))
    worst = None
    nb = 0
    for k0 in sorted({0, max(0, (m // 2) // block * block), max(0, m - block)}):
        k1 = min(m, k0 + block)
        ref = batched.p3d_block(Y, G_host[k0:k1].T, X0, eig_L["values"], eig_L["vectors"], reml_d["delta"], reml_d["vg"])
        one = {k: arrs[k][0] for k in ("effect", "se", "tvalue", "pvalue")}
        c = batched.compare(one, ref, rows=slice(k0, k1))
        nb += c["n_markers"]
        if worst is None:
            worst = c
        else:
            for key in ("beta_rel_max", "se_rel_max", "beta_err_in_se_units_max", "dlog10p_max"):
                worst[key] = max(worst[key], c[key])
            worst["nan_rows_identical"] = worst["nan_rows_identical"] and c["nan_rows_identical"]
            worst["n_tested"] += c["n_tested"]
    worst["n_markers"] = nb
    rep["batched_fp64"] = worst
    rep["tolerance"] = {"beta_se_rel": 1e-8, "dlog10p_abs": 1e-6}
    rep["pass"] = bool(all(v["nan_rows_identical"] and v["se_rel_max"] < 1e-8 and v["beta_rel_max"] < 1e-8 and
                           v["dlog10p_max"] < 1e-6 for v in (rep.get("per_marker_loop"), rep["batched_fp64"]) if v))
    return rep


def run_reference(args):
    """CPU reference arm: the oracle restatement (R is not installable here) on the host cores."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    from gapit_b200 import synth
    from oracle import gapit_oracle as go
    try:   # torchrun exports OMP_NUM_THREADS=1; the CPU reference is entitled to every host core
        from threadpoolctl import threadpool_limits
        threadpool_limits(limits=os.cpu_count() or 1)
    except Exception:
        pass
    n = args.n
    m_kin = min(args.m, 20000)          # markers used to build a realistic K on the CPU
    G = synth.make_genotypes(n, m_kin, seed=synth.BASE_SEED + 1, threads=os.cpu_count() or 8)
    Y = synth.make_phenotypes(G)[:, 0]
    X0 = np.ones((n, 1))
    K = go.kinship_vanraden(G.T)
    eig_L = go.emma_eigen_L_wo_Z(K)
    reml = go.emma_REMLE(Y, X0, K)
    # bounded sample per step: a pilot sizes it for ~3 s of CPU work, so the fixed per-call setup of the loop
    # (forming U, the marker-free model) stays a small fraction, as it is in a full-size run
    ns = args.cpu_sample
    if not ns:
        v0, _ = cpu_scan_sample(n, min(m_kin, 16), 0, eig_L, reml, G, Y, X0)
        ns = int(max(16, min(m_kin, 3.0 * v0)))
    for _ in range(args.warmup):
        cpu_scan_sample(n, max(2, ns // 8), 0, eig_L, reml, G, Y, X0)
    t = []
    for _ in range(args.steps):
        v, dt = cpu_scan_sample(n, ns, 0, eig_L, reml, G, Y, X0)
        t.append(dt)
    dt = float(np.mean(t))
    value = ns / dt
    cores = cpu_threads()
    line = {"impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": dt * 1e3, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": workload_name(n, args.m)},
            "sample": f"{ns} markers per step, per-marker loop with two n x n GEMVs",
            "cpu_baseline": {"value": value, "unit": UNIT, "cores": cores, "kind": "port",
                             "sample": f"{ns} of {args.m} markers per step (cost is exactly linear in markers); "
                                       f"numpy/OpenBLAS oracle restatement of GAPIT.EMMAxP3D.R:397-979, K from {m_kin} markers"},
            "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line), flush=True)


def main():
    args = parse()
    if args.impl == "reference":
        run_reference(args)
        return
    import torch
    import torch.distributed as dist
    import gapit_b200 as gb
    from gapit_b200 import synth, _lib

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    n, m = args.n, args.m
    pk = peaks()

    # ---------------- synthetic inputs (host, pinned) ----------------
    G_np = synth.make_genotypes(n, m, seed=synth.BASE_SEED + 1 + 1000 * rank, threads=min(32, os.cpu_count() or 8))
    G_pin = torch.empty((m, n), dtype=torch.int8).pin_memory()
    G_pin.numpy()[:] = G_np
    del G_np
    G_host = G_pin.numpy()
    Y = synth.make_phenotypes(G_host[:4096] if m > 4096 else G_host)   # QTN drawn from the first markers
    if world > 1:   # every rank must scan the same phenotype
        yt = torch.from_numpy(Y).to(dev)
        dist.broadcast(yt, 0)
        Y = yt.cpu().numpy()
    X0 = np.ones((n, 1))

    ctx = gb.Context(local, slices=args.slices or None)
    S = ctx.slices
    stream = torch.cuda.Stream(device=dev)      # one non-default stream for torch, NCCL and libgapitcuda alike
    torch.cuda.set_stream(stream)
    ctx.set_stream(stream.cuda_stream)
    lib = ctx.lib
    geno = gb.Genotypes(G_host, ctx=ctx, marker_major=True)

    def sync_all():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
            torch.cuda.synchronize()

    def timed(fn, steps, warmup):
        for _ in range(warmup):
            fn()
        sync_all()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        kms = []
        e0.record(stream)
        for _ in range(steps):
            fn()
            kms.append(ctx.stats()["kernel_ms"])
        e1.record(stream)
        sync_all()
        ms = torch.tensor([e0.elapsed_time(e1)], device=dev, dtype=torch.float64)
        if world > 1:
            dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        return float(ms.item()) / steps, float(np.mean(kms))

    # ---------------- kinship: gram (+ exchange) + epilogue ----------------
    # N > 1: the exchange is fused into the Gram kernel -- every rank's epilogue adds its tiles into the buffer of the
    # rank that owns the block-row, over NVLink, as bulk tensor reduce-adds (gapit_vanraden_gram_fused, root = -1:
    # reduce-scatter, each rank's ingress carries 1/N of the tiles; buffers from torch symmetric memory); rank 0 then
    # pulls the rows it does not own (gapit_vanraden_gram_gather) and runs the epilogue.
    # `--kinship-exchange nccl` (or a box without peer access) runs Gram -> ncclAllReduce instead.
    import ctypes
    ldg = lib.gapit_gram_ld(n)
    sums_t = torch.empty(3, dtype=torch.int64, device=dev)
    K_dev = torch.empty((n, n), dtype=torch.float64, device=dev)
    hdl = None
    exchange = "none" if world == 1 else "nccl"
    if world > 1 and args.kinship_exchange == "fused":
        try:
            import torch.distributed._symmetric_memory as symm_mem
            flat = symm_mem.empty(ldg * ldg + 8, dtype=torch.int32, device=dev)   # Gram buffer + mailbox of the 3 scalar sums
            hdl = symm_mem.rendezvous(flat, dist.group.WORLD)
            G_t = flat[:ldg * ldg].view(ldg, ldg)
            box = flat[ldg * ldg:].view(torch.int64)
            peers = (ctypes.c_void_p * world)(*[int(p) for p in hdl.buffer_ptrs])
            boxes = (ctypes.c_void_p * world)(*[int(p) + ldg * ldg * 4 for p in hdl.buffer_ptrs])
            exchange = "fused: tiles reduce-scattered over NVLink inside the Gram kernel (TMA bulk reduce-add), rank 0 gathers the rows"
        except Exception as exc:
            hdl = None
            exchange = f"nccl (symmetric memory unavailable: {exc!r})"[:160]
    if hdl is None:
        G_t = torch.empty((ldg, ldg), dtype=torch.int32, device=dev)

    def kinship_step():
        if hdl is not None:
            flat.zero_()
            hdl.barrier()
            _lib.check(lib.gapit_vanraden_gram_fused(ctx.h, geno.h, peers, world, rank, -1, sums_t.data_ptr(), boxes))
            st = ctx.stats()
            hdl.barrier()              # after it every rank's mailbox holds the scalar sums of all ranks: no all-reduce
            if rank == 0:
                _lib.check(lib.gapit_vanraden_gram_gather(ctx.h, n, peers, world, 0))
                _lib.check(lib.gapit_vanraden_finish(ctx.h, n, G_t.data_ptr(), box.data_ptr(), K_dev.data_ptr(), None, None))
        else:
            _lib.check(lib.gapit_vanraden_gram(ctx.h, geno.h, G_t.data_ptr(), sums_t.data_ptr()))
            if world > 1:
                dist.all_reduce(G_t)
                dist.all_reduce(sums_t)
            st = ctx.stats()
            _lib.check(lib.gapit_vanraden_finish(ctx.h, n, G_t.data_ptr(), sums_t.data_ptr(), K_dev.data_ptr(), None, None))
        ctx._last_gram_ms = st["kernel_ms"]

    kin_ms, _ = timed(kinship_step, args.steps, args.warmup)
    gram_ms = ctx._last_gram_ms
    exchange_ok = None
    if hdl is not None:   # once, outside the timed region: the fused sums against Gram + ncclAllReduce, bit for bit
        G_ref = torch.empty((ldg, ldg), dtype=torch.int32, device=dev)
        s_ref = torch.empty(3, dtype=torch.int64, device=dev)
        _lib.check(lib.gapit_vanraden_gram(ctx.h, geno.h, G_ref.data_ptr(), s_ref.data_ptr()))
        dist.all_reduce(G_ref)
        if rank == 0:
            # G_t was mirrored in place by the last finish: compare the lower triangle
            exchange_ok = bool(torch.equal(torch.tril(G_ref[:n, :n]), torch.tril(G_t[:n, :n])))
        del G_ref, s_ref
    m_total = m * world
    kin_ops = float(n) * (n + 1) * m_total            # one multiply + one add per lower-triangle pair per marker
    kin_tops = kin_ops / (kin_ms * 1e-3) / 1e12
    gram_tops_local = float(n) * (n + 1) * m / (gram_ms * 1e-3) / 1e12

    # ---------------- spectra + REML (rank 0; one-off per trait set; timed separately) ----------------
    # device-resident: K_dev -> eigenvectors in place (cuSOLVER syevd) -> digit planes; REML from V'y, V'X0
    setup = {}
    val_t = torch.empty(n, dtype=torch.float64, device=dev)
    dv_t = torch.empty(2, dtype=torch.float64, device=dev)
    es = None
    if rank == 0:
        K_host = np.asfortranarray(K_dev.cpu().numpy()) if (args.reml_route == "eigR" or not args.no_e2e) else None
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        values = gb.eigh_dev(K_dev.data_ptr(), n, ctx=ctx)          # K_dev now holds V (row r = eigenvector r)
        setup["eigh_L_ms"] = (time.perf_counter() - t0) * 1e3
        val_t.copy_(torch.from_numpy(values))
    if world > 1:
        dist.broadcast(K_dev, 0)
        dist.broadcast(val_t, 0)
    torch.cuda.synchronize()
    values = val_t.cpu().numpy()
    t0 = time.perf_counter()
    es = gb.EigenSystem.from_device(values, K_dev.data_ptr(), ctx=ctx)
    setup["eigsys_slice_ms"] = (time.perf_counter() - t0) * 1e3
    if rank == 0:
        t0 = time.perf_counter()
        if args.reml_route == "eigR":
            eR = gb.emma_eigen_R_wo_Z(K_host, X0, ctx=ctx)
            setup["eigen_R_ms"] = (time.perf_counter() - t0) * 1e3
            t0 = time.perf_counter()
            reml = gb.GAPIT_emma_REMLE(Y[:, 0], X0, K_host, eig_R=eR, ctx=ctx)
            del eR, K_host
        else:
            proj = es.project(np.column_stack([Y[:, 0], X0]))
            reml = gb.reml_from_projections(values, proj[:, 0], proj[:, 1:])
        setup["reml_ms"] = (time.perf_counter() - t0) * 1e3
        setup["reml_route"] = args.reml_route
        dv_t.copy_(torch.tensor([reml["delta"], reml["vg"]], dtype=torch.float64))
    if world > 1:
        dist.broadcast(dv_t, 0)
    delta, vg = [float(x) for x in dv_t.cpu().numpy()]
    vectors = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        vectors = np.asfortranarray(K_dev.cpu().numpy().T)          # host copy for the CPU baseline only
    del K_dev

    # ---------------- the timed hot path: MLM P3D scan ----------------
    # N > 1: rank 0 gathers every shard's p-values.  The rows come back in page-locked memory (reused across steps), so
    # the gather's send buffer is filled by one asynchronous H2D copy on the shared stream, with no staging allocation.
    gather_buf = [torch.empty(m, dtype=torch.float64, device=dev) for _ in range(world)] if (world > 1 and rank == 0) else None
    p_dev = torch.empty(m, dtype=torch.float64, device=dev) if world > 1 else None
    out_arrs = gb.alloc_scan_out(1, m, pinned=True)
    last = {}

    def scan_step():
        arrs, bases = gb.p3d_scan(geno, X0, Y[:, :1], eigsys=es, delta=[delta], vg=[vg], ctx=ctx, pinned=True, out=out_arrs)
        last["arrs"] = arrs
        if world > 1:
            p_dev.copy_(torch.from_numpy(arrs["pvalue"][0]), non_blocking=True)
            dist.gather(p_dev, gather_buf, dst=0)

    sampler = ClockSampler(local)
    sampler.start()
    scan_ms, scan_kernel_ms = timed(scan_step, args.steps, args.warmup)
    clocks = sampler.stop()
    value = m_total / (scan_ms * 1e-3)
    # kernels of this library launched by ONE scan step, counted by CUPTI on an extra untimed step
    launches_per_step, launch_names = count_launches(scan_step)

    # GLM scan (K = NULL branch), HBM-bound
    glm_arrs = gb.alloc_scan_out(1, m, pinned=True)

    def glm_step():
        gb.p3d_scan(geno, X0, Y[:, :1], glm=True, ctx=ctx, pinned=True, out=glm_arrs)

    glm_ms, glm_kernel_ms = timed(glm_step, max(2, args.steps), args.warmup)

    # ---------------- end to end through the public API with host buffers ----------------
    # Headline `e2e`: host genotypes in the library's packed host format (GAPIT_B2: 2 bits per call, the bit order of a
    # PLINK .bed row -- the same 0/1/2 dosages, a quarter of the int8 bytes over PCIe); `e2e.int8` is the same call fed
    # one byte per call.  Both: pinned host input -> gapit_p3d_scan_host (chunked upload on a copy stream overlapped
    # with the scan) -> the seven result arrays back in pinned host memory, every step.
    e2e = None
    if not args.no_e2e:
        P2 = gb.Packed2.from_dosages(G_host, marker_major=True, pinned=True)

        def e2e_packed_step():
            arrs, _ = gb.p3d_scan_host(P2, X0, Y[:, :1], eigsys=es, delta=[delta], vg=[vg], ctx=ctx, pinned=True, out=glm_arrs)
            last["e2e_min_p"] = float(np.nanmin(arrs["pvalue"]))
        e2p_ms, _ = timed(e2e_packed_step, args.steps, 1)

        def e2e_int8_step():
            arrs, _ = gb.p3d_scan_host(G_host, X0, Y[:, :1], eigsys=es, delta=[delta], vg=[vg], ctx=ctx,
                                       marker_major=True, pinned=True, out=glm_arrs)
            last["e2e_int8_min_p"] = float(np.nanmin(arrs["pvalue"]))
        e2e_ms, _ = timed(e2e_int8_step, args.steps, 1)
        assert last["e2e_int8_min_p"] == last["e2e_min_p"]
        e2e = {"value": m_total / (e2p_ms * 1e-3), "unit": UNIT, "ms_per_step": e2p_ms,
               "h2d_bytes_per_step": int(P2.data.nbytes) + 8 * n * 2, "d2h_bytes_per_step": 7 * 8 * int(m),
               "note": "pinned host genotypes at 2 bits per call (GAPIT_B2, PLINK .bed bit order) -> gapit_p3d_scan_host "
                       "(chunked upload + unpack kernel on a copy stream, overlapped with the scan) -> host result arrays; "
                       "eigen-system resident (one-off per trait set, see setup_ms)",
               "int8": {"value": m_total / (e2e_ms * 1e-3), "unit": UNIT, "ms_per_step": e2e_ms,
                        "h2d_bytes_per_step": int(n) * int(m) + 8 * n * 2, "d2h_bytes_per_step": 7 * 8 * int(m),
                        "note": "the same call fed one byte per genotype call (PCIe-bound: 2.5 GB per step and GPU)"}}
        if rank == 0 and world == 1:
            # the whole reference-facing call, nothing resident: K on the host -> eigendecomposition (cuSOLVER syevd) +
            # REML + BLUP/PEV + the scan from packed host genotypes -> the reference's result list.  The eigen phase is
            # outside the hot-path roofline (north_star) but it IS what one GAPIT.EMMAxP3D call costs.
            try:
                whole = []
                for _ in range(2):
                    torch.cuda.synchronize()
                    t0 = time.perf_counter()
                    res = gb.GAPIT_EMMAxP3D(Y[:, 0], P2, K_host, X0=X0, ctx=ctx)
                    whole.append(time.perf_counter() - t0)
                same = bool(np.isclose(float(np.nanmin(res["ps"])), last["e2e_min_p"], rtol=1e-6, atol=0.0))
                e2e["whole_call"] = {"value": m / whole[-1], "unit": UNIT, "ms": whole[-1] * 1e3, "first_call_ms": whole[0] * 1e3,
                                     "min_p_matches_the_resident_scan": same,
                                     "note": "one GAPIT_EMMAxP3D(ys, xs, K, X0) call: host K -> syevd + REML + BLUP/PEV + scan "
                                             "from packed host genotypes -> result list (second of two calls)"}
                del res
            except Exception as exc:   # an extra: it must not take the headline down with it
                e2e["whole_call"] = {"error": repr(exc)[:300]}
        del P2

    # ---------------- CPU baseline beside it (rank 0, N=1) ----------------
    cpu = None
    parity = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        eig_L = {"values": values, "vectors": vectors}
        reml_d = {"delta": delta, "vg": vg, "ve": delta * vg, "REML": 0.0}
        ns = args.cpu_sample
        if not ns:   # pilot, then size the sample for ~15 s of CPU work
            v0, _ = cpu_scan_sample(n, min(m, 16), 0, eig_L, reml_d, G_host, Y[:, 0], X0)
            ns = int(max(16, min(m, 40.0 * v0)))
        kept = {}
        v, dt = cpu_scan_sample(n, ns, 0, eig_L, reml_d, G_host, Y[:, 0], X0, keep=kept)
        cpu = {"value": v, "unit": UNIT, "cores": cpu_threads(), "kind": "port",
               "sample": f"first {ns} of {m} markers ({dt:.1f} s); oracle restatement of GAPIT.EMMAxP3D.R:397-979 "
                         f"(two n x n GEMVs per marker) on numpy/OpenBLAS, eigen-system and delta injected"}
        if not args.no_parity:
            # the same markers, the same eigen-system and delta, the same run: GPU rows against the oracle's
            parity = parity_report(last["arrs"], G_host, Y[:, 0], X0, eig_L, reml_d, kept.get("ora"), ns)
        del kept
    del vectors

    # ---------------- roofline of the dominant kernel (rotate_reduce) ----------------
    int8_ops = 2.0 * n * n * m * S                     # int8-native: S digit planes of V
    fp64_flops = 2.0 * n * n * m                       # algorithmic fp64 flops of the rotation
    achieved = int8_ops / (scan_kernel_ms * 1e-3) / 1e12
    int8_meas = calibrate_int8_tops(dev)
    fp64_meas = calibrate_fp64_tflops(dev)
    int8_peak = int8_meas if int8_meas else 2.0 * pk["bf16_tflops"]
    # DRAM bytes (read + write) of one rotate_reduce launch: from the committed `ncu --set full` capture of this exact
    # shape (profiles/traffic.json names the capture); null for any other shape -- it is not measurable inside this run
    traffic, traffic_src = ncu_traffic(n, m, S)
    roofline = {"bound": "tensor", "kernel": "umma_i8_pair_kernel<ScanSched,ScanEpi> (rotate_reduce, tcgen05 cta_group::2)",
                "achieved": achieved, "peak": int8_peak, "unit": "TFLOP/s", "frac": achieved / int8_peak,
                "traffic": traffic, "traffic_source": traffic_src, "algorithmic_bytes": float(n) * m + float(n) * n * S,
                "peak_note": ("int8 dense ceiling measured in this run: cuBLASLt int8 GEMM 8192^3 via torch._int_mm, best of 10 "
                              "(MEASURED_PEAKS.json holds no int8 figure)" if int8_meas else
                              f"2 x bf16_tflops of MEASURED_PEAKS.json ({pk['source']}); int8 calibration unavailable"),
                "peak_2x_bf16_measured": 2.0 * pk["bf16_tflops"], "peak_nominal": 4500.0,
                "kernel_ms": scan_kernel_ms, "slices": S,
                "fp64_equivalent_tflops": fp64_flops / (scan_kernel_ms * 1e-3) / 1e12,
                # the same rotation as a plain fp64 GEMM would be bounded by cuBLAS DGEMM, measured here
                "fp64_dgemm_tflops_measured": fp64_meas,
                "fp64_equiv_frac": (fp64_flops / (scan_kernel_ms * 1e-3) / 1e12 / fp64_meas) if fp64_meas else None}
    glm_bytes = float(n) * m
    # ---------------- the other BASELINE configs, full size, on this box (outside the timed region above) -------------
    min_p = float(np.nanmin(last["arrs"]["pvalue"]))
    legs = None
    if not args.no_legs:
        from gapit_b200 import fullsize
        es.close()
        geno.close()
        es = geno = None
        del out_arrs, glm_arrs, last["arrs"], G_pin, G_host
        torch.cuda.empty_cache()
        legs = {}
        for name in (["c5"] if world == 1 else ["c3", "c4"]):
            try:
                legs[name] = fullsize.run_config(name, ctx, dev, rank, world, slices=S, exchange=args.kinship_exchange)
            except Exception as exc:   # a leg must not take the headline down with it
                legs[name] = {"error": repr(exc)[:300]}
            torch.cuda.empty_cache()
            leg = legs[name]
            # the same denominator as `roofline.peak`: the cuBLASLt int8 ceiling measured in this run, per GPU
            if "kinship_tops_kernel" in leg:
                leg["kinship_kernel_frac_of_int8_ceiling"] = leg["kinship_tops_kernel"] / world / int8_peak
            if "scan_int8_tops_per_gpu_rotation_once" in leg:
                leg["scan_frac_of_int8_ceiling"] = leg["scan_int8_tops_per_gpu_rotation_once"] / int8_peak
        legs["scaling"] = "strong: total work of the named config fixed, markers sharded over the N ranks" if world > 1 \
            else "one GPU"
    line = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": scan_ms, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": f"int8 x int8 -> int32 rotation ({S} digit planes of fp64 V), fp64 epilogue",
        "data": "synthetic",
        "config": {"workload": workload_name(n, m), "markers_total": m_total, "kinship": "VanRaden", "slices": S,
                   "l2": "inputs larger than L2 (genotypes 2.5 GB per GPU)", "parallelism": f"marker-shard x{world}"},
        "kinship": {"tops": kin_tops, "ms": kin_ms, "gram_kernel_ms": gram_ms, "gram_kernel_tops_per_gpu": gram_tops_local,
                    "ops": "n(n+1)m integer multiply-adds counted once per lower-triangle pair",
                    "frac_of_int8_peak": gram_tops_local / int8_peak,
                    "includes": "gram + exchange (N>1) + fp64 epilogue", "exchange": exchange,
                    "exchange_matches_nccl_bit_for_bit": exchange_ok},
        "glm": {"snps_per_s": m_total / (glm_ms * 1e-3), "ms": glm_ms, "kernel_ms": glm_kernel_ms,
                "hbm_gbs": glm_bytes / (glm_kernel_ms * 1e-3) / 1e9,
                "frac_of_hbm_peak": glm_bytes / (glm_kernel_ms * 1e-3) / 1e9 / pk["hbm_gbs"]},
        "setup_ms": setup,
        "roofline": roofline,
        "cpu_baseline": cpu,
        "e2e": e2e,
        "gpu_launches": (launches_per_step * args.steps) if launches_per_step else args.steps * 3,
        "gpu_launches_note": ({"per_step": launches_per_step, "counted_by": "CUPTI (torch.profiler) on one extra untimed step",
                               "kernels": launch_names} if launches_per_step else
                              f"CUPTI unavailable ({launch_names}); per scan step: vt_project + rotate_reduce + finalize"),
        "parity": parity,
        "legs": legs,
        "clocks": clocks,
        "reml": {"delta": delta, "vg": vg},
        "min_p": min_p,
    }
    if rank == 0:
        print(json.dumps(line), flush=True)
    if es is not None:
        es.close()
        geno.close()
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
