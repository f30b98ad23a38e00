"""Multi-GPU parity (-m gpu, needs >= 2 GPUs; skipped on a one-GPU box): the marker-sharded pipeline under torchrun
with NCCL must reproduce the one-GPU results -- integer cross-products and kinship bit-identical, scan outputs within
1e-13.  The CPU-side plumbing of the same path is covered by tests/test_dist_cpu.py (gloo)."""
import json
import os
import socket
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.mark.timeout(600)
def test_sharded_pipeline_matches_one_gpu():
    import torch
    ngpu = torch.cuda.device_count()
    if ngpu < 2:
        pytest.skip("needs at least 2 GPUs")
    world = 2 if ngpu < 4 else 4
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={world}",
           "--master-addr", "127.0.0.1", "--master-port", str(port), os.path.join(ROOT, "tools", "check_multi_gpu.py")]
    out = subprocess.run(cmd, capture_output=True, text=True, timeout=580)
    assert out.returncode == 0, out.stdout[-2000:] + out.stderr[-2000:]
    line = [ln for ln in out.stdout.splitlines() if ln.startswith("{")][-1]
    rep = json.loads(line)
    assert rep["ok"] and rep["gram_bit_identical"] and rep["kinship_bit_identical"] and rep["scan_max_rel_diff"] <= 1e-13


@pytest.mark.parametrize("ngpus", [1, 2, 4])
def test_one_process_drives_every_gpu_through_the_c_abi(ngpus):
    """gapit_mctx / gapit_mgeno (what the R shim binds): one process, one host thread per GPU, no torch.distributed.
    Kinship bit-identical to the one-GPU call (integer sums), scan rows identical wherever a marker sits."""
    import numpy as np
    import torch
    import gapit_b200 as gb
    from gapit_b200 import synth
    if torch.cuda.device_count() < ngpus:
        pytest.skip(f"needs {ngpus} GPUs")
    n, m = 600, 5003
    G = synth.make_genotypes(n, m, seed=29, threads=4)
    Y = synth.make_phenotypes(G, ntraits=3, seed=29)
    X0 = synth.make_covariates(G, npc=1)
    ctx = gb.Context(0)
    one = gb.Genotypes(G, ctx=ctx, marker_major=True)
    K1, G1 = gb.GAPIT_kinship_VanRaden(one, ctx=ctx, return_gram=True, verbose=False)
    eL = gb.emma_eigen_L_wo_Z(K1, ctx=ctx)
    remls = [gb.GAPIT_emma_REMLE(Y[:, t], X0, K1, eig_L=eL, ctx=ctx) for t in range(3)]
    delta, vg = [r["delta"] for r in remls], [r["vg"] for r in remls]
    es = gb.EigenSystem(eL["values"], eL["vectors"], ctx=ctx)
    ref, refb = gb.p3d_scan(one, X0, Y, eigsys=es, delta=delta, vg=vg, ctx=ctx)
    ref1, _ = gb.p3d_scan(one, X0, Y[:, :1], eigsys=es, delta=delta[:1], vg=vg[:1], ctx=ctx)
    refg, _ = gb.p3d_scan(one, X0, Y, glm=True, ctx=ctx)

    mctx = gb.MultiContext(ngpus)
    assert mctx.ngpus == ngpus
    for src in (G, gb.Packed2.from_dosages(G, marker_major=True)):
        mg = gb.MultiGenotypes(src, mctx, marker_major=True)
        K, Gm = gb.multi_vanraden(mg, return_gram=True)
        assert np.array_equal(Gm, G1) and np.array_equal(K, K1)
        got, gotb = gb.multi_p3d_scan(mg, X0, Y, values=eL["values"], vectors=eL["vectors"], delta=delta, vg=vg)
        got1, _ = gb.multi_p3d_scan(mg, X0, Y[:, :1], values=eL["values"], vectors=eL["vectors"], delta=delta[:1], vg=vg[:1])
        gotg, _ = gb.multi_p3d_scan(mg, X0, Y, glm=True)
        for key in ("effect", "tvalue", "se", "pvalue", "rsquare", "loglm", "maf"):
            assert np.array_equal(got[key], ref[key], equal_nan=True), key
            assert np.array_equal(got1[key], ref1[key], equal_nan=True), key
            assert np.array_equal(gotg[key], refg[key], equal_nan=True), key
        assert gotb[0]["logL0"] == refb[0]["logL0"] and gotb[2]["rsquare_base"] == refb[2]["rsquare_base"]
        mg.close()
    mctx.close()


@pytest.mark.parametrize("ngpus", [2, 4])
def test_eigen_decomposition_sliced_over_the_gpus(ngpus):
    """gapit_multi_eigh (eigen(K) of the R shim): every GPU solves for n/ngpus eigenpairs by index range; side by side
    they are one orthonormal eigen-system with the one-GPU solver's values."""
    import numpy as np
    import torch
    import gapit_b200 as gb
    if torch.cuda.device_count() < ngpus:
        pytest.skip(f"needs {ngpus} GPUs")
    rng = np.random.default_rng(9)
    n = 1024 * ngpus + 77                         # the route needs >= 1,024 eigenpairs per GPU
    A = rng.normal(size=(n, n + 30))
    K = A @ A.T / (n + 30)
    mctx = gb.MultiContext(ngpus)
    got = gb.multi_eigh(K, mctx)
    ref = gb.emma_eigen_L_wo_Z(K, ctx=mctx.context(0))
    vals, V = got["values"], got["vectors"]
    assert np.all(np.diff(vals) <= 1e-12) and np.allclose(vals, ref["values"], rtol=0, atol=1e-12 * vals[0])
    assert np.abs(V.T @ V - np.eye(n)).max() < 1e-11
    assert np.abs(K @ V - V * vals).max() < 1e-11 * vals[0]
    mctx.close()


@pytest.mark.parametrize("ngpus", [2, 4])
def test_resident_eigen_system_on_every_gpu(ngpus):
    """gapit_meigsys (what one GAPIT.EMMAxP3D call of the R shim holds): K decomposed with the spectrum sliced over the
    GPUs, slices exchanged over NVLink, a sliced copy on every GPU; REML projections, BLUP/PEV and the scan read it.
    Against the one-GPU pipeline: same eigenvalues, same delta, same p-values (the eigenvectors may differ by sign)."""
    import numpy as np
    import torch
    import gapit_b200 as gb
    from gapit_b200 import synth
    if torch.cuda.device_count() < ngpus:
        pytest.skip(f"needs {ngpus} GPUs")
    n, m = 1024 * ngpus + 50, 4001
    G = synth.make_genotypes(n, m, seed=31, threads=4)
    Y = synth.make_phenotypes(G, ntraits=2, seed=31)
    X0 = synth.make_covariates(G, npc=1)
    ctx = gb.Context(0)
    one = gb.Genotypes(G, ctx=ctx, marker_major=True)
    K = gb.GAPIT_kinship_VanRaden(one, ctx=ctx, verbose=False)
    es1 = gb.EigenSystem.from_kinship(K, ctx=ctx)
    p1 = es1.project(np.column_stack([X0, Y]))
    rs1 = gb.reml_from_projections_multi(es1.values, p1[:, 2:], p1[:, :2])
    d1, v1 = [r["delta"] for r in rs1], [r["vg"] for r in rs1]
    ref, _ = gb.p3d_scan(one, X0, Y, eigsys=es1, delta=d1, vg=v1, ctx=ctx)
    b1 = gb.blup_pev(es1, X0, Y[:, 1], d1[1], v1[1], ctx=ctx)

    mctx = gb.MultiContext(ngpus)
    me = gb.MultiEigenSystem(mctx, K=K)
    assert np.allclose(me.values, es1.values, rtol=0, atol=1e-12 * es1.values[0])
    pm = me.part(0).project(np.column_stack([X0, Y]))
    rsm = gb.reml_from_projections_multi(me.values, pm[:, 2:], pm[:, :2])
    dm, vm = [r["delta"] for r in rsm], [r["vg"] for r in rsm]
    assert np.allclose(dm, d1, rtol=1e-9) and np.allclose(vm, v1, rtol=1e-9)
    mg = gb.MultiGenotypes(G, mctx, marker_major=True)
    got, _ = gb.multi_p3d_scan(mg, X0, Y, delta=d1, vg=v1, eigsys=me)
    for key in ("effect", "se", "rsquare", "maf"):
        assert np.allclose(got[key], ref[key], rtol=1e-8, atol=1e-12, equal_nan=True), key
    with np.errstate(divide="ignore"):
        assert np.nanmax(np.abs(np.log10(got["pvalue"]) - np.log10(ref["pvalue"]))) < 1e-7
    bm = gb.blup_pev(me.part(0), X0, Y[:, 1], d1[1], v1[1], ctx=mctx.context(0))
    for x, y in zip(bm, b1):
        assert np.allclose(x, y, rtol=1e-7, atol=1e-9)
    # the upload form is the host-eigenvector call, bit for bit
    eL = gb.emma_eigen_L_wo_Z(K, ctx=ctx)
    mu = gb.MultiEigenSystem(mctx, values=eL["values"], vectors=eL["vectors"])
    a, _ = gb.multi_p3d_scan(mg, X0, Y, delta=d1, vg=v1, eigsys=mu)
    b, _ = gb.multi_p3d_scan(mg, X0, Y, values=eL["values"], vectors=eL["vectors"], delta=d1, vg=v1)
    assert all(np.array_equal(a[k], b[k], equal_nan=True) for k in a)
    for h in (mu, me, mg, mctx):
        h.close()

