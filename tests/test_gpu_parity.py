"""GPU parity tests (-m gpu): libgapitcuda through its C ABI (ctypes) against the oracle.

Tolerances are BASELINE.json's north_star: integer cross-products bit-exact, kinship 1e-12
(max-abs relative to max|K|, against the exact-rational value), beta/SE 1e-8 relative,
-log10(p) 1e-6 absolute, marker order identical.
"""
import numpy as np
import pytest

from gapit_b200 import synth
from oracle import crosscheck as cc
from oracle import gapit_oracle as go
from conftest import c1_inputs

pytestmark = pytest.mark.gpu

BETA_RTOL = 1e-8
SE_RTOL = 1e-8
LOGP_ATOL = 1e-6
K_RTOL = 1e-12


def gb():
    import gapit_b200
    return gapit_b200


def case(n, m, npc=0, seed=1, edge=True):
    G = synth.make_genotypes(n, m, seed=seed, threads=4)
    if edge and m > 25:
        G[3, :] = 0
        G[7, :] = 2
        G[11, :] = 1
        G[20, :] = G[19, :]
    Y = synth.make_phenotypes(G, seed=seed)[:, 0]
    X0 = synth.make_covariates(G, npc=npc)
    return G, Y, X0


def rel_err(got, ref, floor):
    """max relative error over entries whose reference magnitude exceeds `floor` * max|ref|,
    and max absolute error (relative to max|ref|) over the rest."""
    ok = ~np.isnan(ref)
    scale = np.nanmax(np.abs(ref))
    big = ok & (np.abs(ref) > floor * scale)
    small = ok & ~big
    r = np.max(np.abs(got[big] - ref[big]) / np.abs(ref[big])) if big.any() else 0.0
    a = np.max(np.abs(got[small] - ref[small])) / scale if small.any() else 0.0
    return r, a


def check_scan(res, ora, glm=False):
    ps, stats = res["ps"][:, 0], res["stats"][:, 0]
    assert np.array_equal(np.isnan(stats), np.isnan(ora.stats)), "NA rows differ"
    const = np.isnan(ora.stats)
    assert np.all(ps[const] == 1.0)
    r, a = rel_err(res["effect.est"][:, 0], ora.effect_est, 1e-4)
    assert r < BETA_RTOL and a < BETA_RTOL * 1e-4 * 10, (r, a)
    r, a = rel_err(res["se"][:, 0], ora.se_clean, 0.0)
    assert r < SE_RTOL, r
    r, a = rel_err(stats, ora.stats, 1e-4)
    assert r < BETA_RTOL and a < BETA_RTOL, (r, a)
    with np.errstate(divide="ignore"):
        dlp = np.abs(np.log10(ps) - np.log10(ora.ps))
    dlp[(ps == 0) & (ora.ps == 0)] = 0.0
    assert np.nanmax(dlp) < LOGP_ATOL, np.nanmax(dlp)
    ok = ~const
    assert np.max(np.abs(res["rsquare"][ok, 0] - ora.rsquare[ok])) < 1e-9
    assert np.max(np.abs(res["logLM"][ok, 0] - ora.logLM[ok]) / np.abs(ora.logLM[ok])) < 1e-10
    assert np.max(np.abs(res["rsquare_base"][ok, 0] - ora.rsquare_base[ok])) < 1e-10
    assert np.array_equal(res["maf"][:, 0], ora.maf)
    assert np.array_equal(res["nobs"][:, 0], ora.nobs)
    assert np.array_equal(res["dfs"][ok, 0], ora.dfs[ok])


# ------------------------------------------------------------------ genotype ingestion
@pytest.mark.parametrize("n,m", [(5, 3), (127, 129), (281, 3093), (300, 1000)])
def test_pack_and_column_stats(gpu_ctx, n, m):
    G, _, _ = case(n, m, edge=False)
    x = G.astype(np.int64)
    for arr, mm in [(G, True), (G.T.astype(np.float64), False), (G.T.astype(np.int32), False),
                    (np.asfortranarray(G.T), False)]:
        g = gb().Genotypes(arr, ctx=gpu_ctx, marker_major=mm)
        cs, cq = g.colsums()
        assert (g.n, g.m) == (n, m)
        assert np.array_equal(cs, x.sum(1)) and np.array_equal(cq, (x * x).sum(1))
        g.close()


@pytest.mark.parametrize("n,m", [(5, 3), (127, 129), (281, 3093), (1001, 517)])
def test_packed_2bit_ingestion_is_bit_identical(gpu_ctx, n, m):
    """GAPIT_B2 host rows (4 genotypes per byte) must land on the device exactly like the int8 rows."""
    G, Y, X0 = case(n, m, edge=False)
    x = G.astype(np.int64)
    P = gb().Packed2.from_dosages(G, marker_major=True)
    assert P.data.shape == (m, (n + 3) // 4)
    g = gb().Genotypes(P, ctx=gpu_ctx)
    cs, cq = g.colsums()
    assert np.array_equal(cs, x.sum(1)) and np.array_equal(cq, (x * x).sum(1))
    K2, gram2 = gb().GAPIT_kinship_VanRaden(g, ctx=gpu_ctx, return_gram=True, verbose=False)
    g.close()
    K1, gram1 = gb().GAPIT_kinship_VanRaden(G.T, ctx=gpu_ctx, return_gram=True, verbose=False)
    assert np.array_equal(gram1, gram2) and np.array_equal(K1, K2)
    # NA code 3: rejected without a rule, imputed with one (GAPIT.EMMAxP3D.R:20-24); a padded row pitch is honoured
    na = G.astype(np.int64).copy()
    na[m // 2, n // 2] = -1
    Pn = gb().Packed2.from_dosages(na, marker_major=True)
    with pytest.raises(gb().GapitCudaError):
        gb().Genotypes(Pn, ctx=gpu_ctx)
    wide = np.full((m, Pn.data.shape[1] + 3), 0xFF, dtype=np.uint8)
    wide[:, :Pn.data.shape[1]] = Pn.data
    for rule, val in [("Middle", 1), ("Major", 2), ("Minor", 0)]:
        g = gb().Genotypes(gb().Packed2(wide, n), ctx=gpu_ctx, impute=rule)
        ref = x.copy()
        ref[m // 2, n // 2] = val
        assert np.array_equal(g.colsums()[0], ref.sum(1))
        g.close()


def test_packed_2bit_streaming_scan_equals_int8_scan(gpu_ctx, monkeypatch):
    G, Y, X0 = case(300, 2100, npc=2)
    K = go.kinship_vanraden(G.T)
    eL = go.emma_eigen_L_wo_Z(K)
    reml = go.emma_REMLE(Y, X0, K)
    es = gb().EigenSystem(eL["values"], eL["vectors"], ctx=gpu_ctx)
    monkeypatch.setenv("GAPIT_STREAM_CHUNK", "512")      # several blocks, ragged tail
    kw = dict(eigsys=es, delta=[reml["delta"]], vg=[reml["vg"]], ctx=gpu_ctx)
    a8, _ = gb().p3d_scan_host(G, X0, Y, marker_major=True, **kw)
    a2, _ = gb().p3d_scan_host(gb().Packed2.from_dosages(G, marker_major=True), X0, Y, **kw)
    for k in a8:
        assert np.array_equal(a8[k], a2[k], equal_nan=True), k
    g8, _ = gb().p3d_scan_host(G, X0, Y, marker_major=True, glm=True, ctx=gpu_ctx)
    g2, _ = gb().p3d_scan_host(gb().Packed2.from_dosages(G, marker_major=True), X0, Y, glm=True, ctx=gpu_ctx)
    for k in g8:
        assert np.array_equal(g8[k], g2[k], equal_nan=True), k
    es.close()


# ------------------------------------------------------------------ HapMap numericalization
@pytest.mark.parametrize("bit", [1, 2])
@pytest.mark.parametrize("maz", [False, True])
def test_hapmap_numericalization_matches_oracle(gpu_ctx, bit, maz):
    """The text of a HapMap table through gapit_hapmap_numericalize against GAPIT.Numericalization restated
    (oracle.hapmap), byte for byte, for every impute rule and effect model."""
    G, _, _ = case(203, 331, edge=False)
    text = synth.make_hapmap_text(G, bit=bit, missing=0.03, crlf=(bit == 1))
    rows = synth.hapmap_rows(text)
    for impute in ("None", "Middle", "Minor", "Major"):
        for effect in ("Add", "Dom", "Left", "Right"):
            ref = go.hapmap(rows, SNP_effect=effect, SNP_impute=impute, Major_allele_zero=maz)
            got = gb().GAPIT_HapMap(text, SNP_effect=effect, SNP_impute=impute, Major_allele_zero=maz, ctx=gpu_ctx)
            gd = got["GD"].astype(np.float64)
            gd[gd < 0] = np.nan
            assert np.array_equal(gd, ref["GD"], equal_nan=True), (impute, effect)
            assert got["GT"] == ref["GT"] and got["GI"] == ref["GI"]


def test_ingest_golden_fixture(gpu_ctx, ingest_golden):
    """The committed known-answer vectors of the widened rows: text -> dosages byte for byte, ZHANG within 1e-11, PCA
    variances within 1e-10 and leading scores up to sign."""
    g = ingest_golden
    for bit in (1, 2):
        text = g[f"text{bit}"].tobytes()
        for impute in ("None", "Middle", "Minor", "Major"):
            for maz in (False, True):
                got = gb().GAPIT_HapMap(text, SNP_impute=impute, Major_allele_zero=maz, ctx=gpu_ctx)["GD"]
                assert np.array_equal(got, g[f"GD{bit}_{impute}_{int(maz)}"]), (bit, impute, maz)
        dom = gb().GAPIT_HapMap(text, SNP_effect="Dom", SNP_impute="Middle", ctx=gpu_ctx)["GD"]
        assert np.array_equal(dom, g[f"GD{bit}_Dom"])
    X = g["GD2_Major_0"]
    Kz = gb().GAPIT_kinship_ZHANG(X, ctx=gpu_ctx, verbose=False)
    assert np.abs(Kz - g["zhang"]).max() <= 1e-11 * np.abs(g["zhang"]).max()
    pc = gb().GAPIT_PCA(X, PC_number=4, ctx=gpu_ctx, verbose=False)
    assert np.abs(pc["EV"] - g["pca_ev"]).max() <= 1e-10 * g["pca_ev"][0]
    assert np.abs(np.abs(pc["PCs"]) - g["pca_x_abs"]).max() <= 1e-8 * np.abs(g["pca_x_abs"]).max()


@pytest.mark.parametrize("n,m,bit", [(4100, 37, 2), (16500, 21, 2), (16500, 21, 1), (3, 9, 2)])
def test_hapmap_row_widths(gpu_ctx, n, m, bit):
    """CTA width and staging depend on the row width: one-warp CTAs (n <= 4096), 128 / 256 threads, and rows too wide
    for the shared-memory stage (read from global memory)."""
    rng = np.random.default_rng(n + bit)
    G = rng.integers(0, 3, size=(m, n)).astype(np.int8)
    text = synth.make_hapmap_text(G, bit=bit, missing=0.01, odd_rows=(n > 10))
    rows = synth.hapmap_rows(text)
    for impute, maz in (("Minor", True), ("None", False)):
        ref = go.hapmap(rows, SNP_impute=impute, Major_allele_zero=maz)
        got = gb().GAPIT_HapMap(text, SNP_impute=impute, Major_allele_zero=maz, ctx=gpu_ctx)
        gd = got["GD"].astype(np.float64)
        gd[gd < 0] = np.nan
        assert np.array_equal(gd, ref["GD"], equal_nan=True), (impute, maz)


def test_by_file_scan_from_hapmap_text(gpu_ctx):
    """GAPIT.EMMAxP3D's by-file branch (:262-330): fragments of the HapMap text are numericalized and scanned on the GPU;
    the result must equal the oracle's numericalization followed by the oracle's scan, fragment size notwithstanding."""
    G, Y, X0 = case(180, 700, npc=1, edge=False)
    text = synth.make_hapmap_text(G, bit=2, missing=0.02)
    GD = go.hapmap(synth.hapmap_rows(text), SNP_impute="Middle")["GD"]          # n x m, what the reference would scan
    K = go.kinship_vanraden(GD)
    eL = go.emma_eigen_L_wo_Z(K)
    reml = go.emma_REMLE(Y, X0, K)
    ora = go.emmax_p3d(Y, GD, K, X0, eig_L=eL, reml=reml)
    es = gb().EigenSystem(eL["values"], eL["vectors"], ctx=gpu_ctx)
    res = {}
    for frag in (97, 256, 100000):
        arrs, bases, GT, GI = gb().p3d_scan_hapmap(text, X0, Y, eigsys=es, delta=[reml["delta"]], vg=[reml["vg"]],
                                                   file_fragment=frag, ctx=gpu_ctx)
        res[frag] = arrs
        assert len(GT) == 180 and len(GI) == 700 and GI[5][0] == "rs5"
    for k in res[97]:
        assert np.array_equal(res[97][k], res[256][k], equal_nan=True) and np.array_equal(res[97][k], res[100000][k], equal_nan=True)
    a = res[256]
    assert np.array_equal(np.isnan(a["tvalue"][0]), np.isnan(ora.stats))
    r, _ = rel_err(a["effect"][0], ora.effect_est, 1e-4)
    assert r < BETA_RTOL
    with np.errstate(divide="ignore"):
        assert np.nanmax(np.abs(np.log10(a["pvalue"][0]) - np.log10(ora.ps))) < LOGP_ATOL
    es.close()


def test_hapmap_to_device_genotypes_and_single_marker_call(gpu_ctx):
    G, _, _ = case(150, 260, edge=False)
    text = synth.make_hapmap_text(G, bit=2, missing=0.02, heading=False)
    hm = gb().GAPIT_HapMap(text, heading=False, SNP_impute="Major", ctx=gpu_ctx, device=True)
    assert hm["GT"] is None and hm["GD"].shape == (150, 260)
    x = hm["GD"].astype(np.int64)
    cs, cq = hm["geno"].colsums()
    assert np.array_equal(cs, x.sum(0)) and np.array_equal(cq, (x * x).sum(0))
    K1 = gb().GAPIT_kinship_VanRaden(hm["geno"], ctx=gpu_ctx, verbose=False)
    K2 = gb().GAPIT_kinship_VanRaden(hm["GD"], ctx=gpu_ctx, verbose=False)
    assert np.array_equal(K1, K2)
    hm["geno"].close()
    # impute None leaves NA: fine for GD, an error for a device handle
    with pytest.raises(gb().GapitCudaError):
        gb().GAPIT_HapMap(text, heading=False, SNP_impute="None", ctx=gpu_ctx, device=True)
    # the per-marker entry point (dense code vector, GAPIT.HMP2Num.R:36)
    row = synth.hapmap_rows(text)[7][11:]
    for impute in ("None", "Minor"):
        got = gb().GAPIT_Numericalization(row, bit=2, impute=impute, ctx=gpu_ctx)
        assert got.shape == (150, 1)
        assert np.array_equal(got[:, 0], go.numericalization(row, impute=impute), equal_nan=True)


def test_device_resident_ingestion(gpu_ctx):
    import torch
    G, _, _ = case(131, 257, edge=False)
    t = torch.from_numpy(G).cuda()
    g = gb().Genotypes.from_device(t.data_ptr(), 131, 257, ctx=gpu_ctx)
    x = G.astype(np.int64)
    cs, cq = g.colsums()
    assert np.array_equal(cs, x.sum(1)) and np.array_equal(cq, (x * x).sum(1))
    K1 = gb().GAPIT_kinship_VanRaden(g, ctx=gpu_ctx, verbose=False)
    assert np.array_equal(K1, gb().GAPIT_kinship_VanRaden(G.T, ctx=gpu_ctx, verbose=False))
    g.close()
    t[5, 7] = 3
    torch.cuda.synchronize()        # the context runs on its own stream: the producer must have finished
    with pytest.raises(gb().GapitCudaError):
        gb().Genotypes.from_device(t.data_ptr(), 131, 257, ctx=gpu_ctx)


def test_pack_rejects_bad_values_and_imputes_na(gpu_ctx):
    G, _, _ = case(40, 60, edge=False)
    bad = G.T.astype(np.float64)
    bad[3, 5] = 3.0
    with pytest.raises(gb().GapitCudaError):
        gb().Genotypes(bad, ctx=gpu_ctx)
    bad8 = G.copy()
    bad8[5, 3] = 7
    with pytest.raises(gb().GapitCudaError):
        gb().Genotypes(bad8, ctx=gpu_ctx, marker_major=True)
    na = G.T.astype(np.float64)
    na[2, 4] = np.nan
    with pytest.raises(gb().GapitCudaError):
        gb().Genotypes(na, ctx=gpu_ctx)
    nai = G.T.astype(np.int32)
    nai[2, 4] = np.iinfo(np.int32).min                                 # R's NA_integer_
    with pytest.raises(gb().GapitCudaError):
        gb().Genotypes(nai, ctx=gpu_ctx)
    for rule, val in [("Middle", 1), ("Major", 2), ("Minor", 0)]:      # GAPIT.EMMAxP3D.R:20-24
        for arr in (na, nai):
            g = gb().Genotypes(arr, ctx=gpu_ctx, impute=rule)
            ref = G.T.astype(np.int64).copy()
            ref[2, 4] = val
            assert np.array_equal(g.colsums()[0], ref.sum(0))
            g.close()
    neg0 = G.T.astype(np.float64)
    neg0[neg0 == 0] = -0.0                                             # -0 is a valid dosage 0
    g = gb().Genotypes(neg0, ctx=gpu_ctx)
    assert np.array_equal(g.colsums()[0], G.T.astype(np.int64).sum(0))
    g.close()


# ------------------------------------------------------------------ kinship
@pytest.mark.parametrize("n,m", [(4, 3), (37, 50), (281, 3093), (257, 1000), (600, 4097), (1030, 20000)])
def test_vanraden_gram_bit_exact_and_kinship(gpu_ctx, n, m):
    G, _, _ = case(n, m)
    geno = gb().Genotypes(G, ctx=gpu_ctx, marker_major=True)
    K, gram = gb().GAPIT_kinship_VanRaden(geno, ctx=gpu_ctx, return_gram=True, verbose=False)
    x = G.astype(np.int64)
    assert np.array_equal(gram, x.T @ x), "integer cross-product must be bit-exact"
    K_ref = go.kinship_vanraden(G.T)
    assert np.abs(K - K_ref).max() <= K_RTOL * np.abs(K_ref).max()
    assert np.array_equal(K, K.T)
    if n <= 300:
        Kex = cc.kinship_exact(G.T)
        assert np.abs(K - Kex).max() <= 4 * np.finfo(float).eps * np.abs(Kex).max()


def test_vanraden_hand_case_and_degenerate(gpu_ctx):
    snps = np.array([[0, 1, 2], [1, 1, 0], [2, 0, 0], [1, 2, 2]], dtype=float)
    K = gb().GAPIT_kinship_VanRaden(snps, ctx=gpu_ctx, verbose=False)
    p = snps.sum(0) / 8.0
    Z = snps - 2 * p
    assert np.allclose(K, Z @ Z.T / (2 * np.sum(p * (1 - p))), atol=1e-15)
    # every column monomorphic: the reference divides 0 by 0 -> NaN matrix
    mono = np.tile(np.array([[0.0, 2.0, 2.0]]), (6, 1))
    Km = gb().GAPIT_kinship_VanRaden(mono, ctx=gpu_ctx, verbose=False)
    assert np.all(np.isnan(Km)) and np.all(np.isnan(go.kinship_vanraden(mono)))


def test_vanraden_sharded_partials_sum_exactly(gpu_ctx):
    """Two marker shards on one GPU, partial integers added on the host = one-shot result (N-GPU path)."""
    import ctypes as C
    import torch
    from gapit_b200 import _lib
    G, _, _ = case(300, 5000)
    lib = gpu_ctx.lib
    n = 300
    ldg = lib.gapit_gram_ld(n)
    tot_G = torch.zeros((ldg, ldg), dtype=torch.int32, device="cuda")
    tot_s = torch.zeros(3, dtype=torch.int64, device="cuda")
    for lo, hi in [(0, 1777), (1777, 5000)]:
        g = gb().Genotypes(G[lo:hi], ctx=gpu_ctx, marker_major=True)
        Gp = torch.empty((ldg, ldg), dtype=torch.int32, device="cuda")
        sp = torch.empty(3, dtype=torch.int64, device="cuda")
        _lib.check(lib.gapit_vanraden_gram(gpu_ctx.h, g.h, Gp.data_ptr(), sp.data_ptr()))
        tot_G += Gp
        tot_s += sp
        g.close()
    K = np.empty((n, n), order="F")
    _lib.check(lib.gapit_vanraden_finish(gpu_ctx.h, n, tot_G.data_ptr(), tot_s.data_ptr(), None,
                                         K.ctypes.data_as(C.c_void_p), None))
    K1 = gb().GAPIT_kinship_VanRaden(gb().Genotypes(G, ctx=gpu_ctx, marker_major=True), ctx=gpu_ctx, verbose=False)
    assert np.array_equal(K, K1), "sharded and one-shot kinship must be bit-identical"


@pytest.mark.parametrize("n,m", [(60, 500), (281, 3093), (300, 1000)])
def test_zhang_kinship_and_pca(gpu_ctx, n, m):
    """GAPIT.kinship.ZHANG and prcomp (GAPIT.PCA.R:10) from the same exact integer cross-product."""
    G, _, _ = case(n, m)
    Kz = gb().GAPIT_kinship_ZHANG(G.T, ctx=gpu_ctx, verbose=False)
    ref = go.kinship_zhang(G.T)
    assert np.abs(Kz - ref).max() <= 1e-11 * np.abs(ref).max()
    assert np.array_equal(Kz, Kz.T)
    pc = gb().GAPIT_PCA(G.T, PC_number=5, ctx=gpu_ctx, verbose=False)
    x, ev = go.prcomp(G.T)
    assert pc["EV"].shape == (min(n, m),)
    assert np.abs(pc["EV"] - ev).max() <= 1e-10 * ev[0]
    for c in range(5):                     # leading components are well separated; the sign is LAPACK's to choose
        a, b = pc["PCs"][:, c], x[:, c]
        sgn = 1.0 if np.dot(a, b) > 0 else -1.0
        gap = min(ev[c - 1] - ev[c] if c else np.inf, ev[c] - ev[c + 1])
        assert np.abs(sgn * a - b).max() <= 1e-9 * np.abs(b).max() * ev[0] / gap


def test_gram_checksums_at_scale(gpu_ctx):
    """Size-independent identities at a size the oracle cannot reach quickly: sum_ij G_ij = sum_k c_k^2,
    trace G = sum_k sum_i x_ik^2, and K 1 = 0."""
    n, m = 3000, 200000
    G = synth.make_genotypes(n, m, seed=77, threads=8)
    g = gb().Genotypes(G, ctx=gpu_ctx, marker_major=True)
    K, gram = gb().GAPIT_kinship_VanRaden(g, ctx=gpu_ctx, return_gram=True, verbose=False)
    cs, cq = g.colsums()
    c = cs.astype(np.int64)
    assert int(gram.astype(np.int64).sum()) == int((c * c).sum())
    assert int(np.trace(gram.astype(np.int64))) == int(cq.astype(np.int64).sum())
    assert np.array_equal(gram, gram.T)
    assert np.abs(K.sum(1)).max() < 1e-9
    sub = G[:, :64].astype(np.int64)
    assert np.array_equal(gram[:64, :64], sub.T @ sub)
    g.close()


# ------------------------------------------------------------------ spectra + REML
def test_eigh_eigenR_reml_match_oracle(gpu_ctx):
    G, Y, X0 = case(220, 1500, npc=2)
    K = go.kinship_vanraden(G.T)
    eL = gb().emma_eigen_L_wo_Z(K, ctx=gpu_ctx)
    eLo = go.emma_eigen_L_wo_Z(K)
    assert np.abs(eL["values"] - eLo["values"]).max() < 1e-11
    assert np.all(np.diff(eL["values"]) <= 1e-12)                      # R's decreasing order
    rec = (eL["vectors"] * eL["values"]) @ eL["vectors"].T
    assert np.abs(rec - K).max() < 1e-11
    eR = gb().emma_eigen_R_wo_Z(K, X0, ctx=gpu_ctx)
    eRo = go.emma_eigen_R_wo_Z(K, X0)
    assert eR["values"].shape == (220 - 3,) and eR["vectors"].shape == (220, 217)
    assert np.abs(eR["values"] - eRo["values"]).max() < 1e-11
    assert np.abs(eR["vectors"].T @ X0).max() < 1e-9
    r1 = gb().GAPIT_emma_REMLE(Y, X0, K, ctx=gpu_ctx)                     # own eig.R
    r2 = go.emma_REMLE(Y, X0, K)
    for k in ("REML", "delta", "ve", "vg"):
        assert r1[k] == pytest.approx(r2[k], rel=1e-7), k


def test_device_resident_spectra_match_host_route(gpu_ctx):
    """K on the device -> gapit_eigh_dev in place -> gapit_eigsys_from_device -> projections -> gapit_reml_eigL:
    same numbers as the host-buffer route (gapit_eigh + gapit_eigsys_upload + eig.R REML)."""
    import torch
    G, Y, X0 = case(200, 900, npc=2)
    K = gb().GAPIT_kinship_VanRaden(G.T, ctx=gpu_ctx, verbose=False)
    eL = gb().emma_eigen_L_wo_Z(K, ctx=gpu_ctx)
    r_ref = gb().GAPIT_emma_REMLE(Y, X0, K, ctx=gpu_ctx)                  # eig.R route (GAPIT.emma.REMLE.R:21-54)
    Kd = torch.from_numpy(np.ascontiguousarray(K)).cuda()
    torch.cuda.synchronize()
    vals = gb().eigh_dev(Kd.data_ptr(), 200, ctx=gpu_ctx)
    assert np.abs(vals - eL["values"]).max() < 1e-12 * np.abs(vals).max()
    es = gb().EigenSystem.from_device(vals, Kd.data_ptr(), ctx=gpu_ctx)
    proj = es.project(np.column_stack([Y, X0]))
    Vd = Kd.cpu().numpy().T                                               # column k = eigenvector k
    assert np.abs(proj - Vd.T @ np.column_stack([Y, X0])).max() < 1e-11
    r = gb().reml_from_projections(vals, proj[:, 0], proj[:, 1:])
    for k in ("REML", "delta", "ve", "vg"):
        assert r[k] == pytest.approx(r_ref[k], rel=1e-9), k
    a1, _ = gb().p3d_scan(gb().Genotypes(G, ctx=gpu_ctx, marker_major=True), X0, Y, eigsys=es, delta=[r["delta"]],
                          vg=[r["vg"]], ctx=gpu_ctx)
    ora = go.emmax_p3d(Y, G.T, K, X0, eig_L=go.emma_eigen_L_wo_Z(K), reml=r_ref)
    rr, _ = rel_err(a1["effect"][0], ora.effect_est, 1e-3)
    assert rr < 1e-7                                                      # two independent eigen-solvers + REML searches
    es.close()


# ------------------------------------------------------------------ the scan
@pytest.mark.parametrize("n,m,npc", [(48, 160, 2), (281, 3093, 0), (200, 700, 3), (333, 900, 6), (150, 300, 12)])
def test_mlm_scan_parity_injected_spectrum(gpu_ctx, n, m, npc):
    G, Y, X0 = case(n, m, npc=npc)
    K = go.kinship_vanraden(G.T)
    eL = go.emma_eigen_L_wo_Z(K)
    reml = go.emma_REMLE(Y, X0, K)
    ora = go.emmax_p3d(Y, G.T, K, X0, ncol_CVI=X0.shape[1], eig_L=eL, reml=reml)
    res = gb().GAPIT_EMMAxP3D(Y, G.T.astype(np.float64), K, X0=X0, CVI=np.zeros((n, X0.shape[1])), ctx=gpu_ctx,
                              eig_L=eL, reml=reml)
    check_scan(res, ora)
    ok = ~np.isnan(ora.stats)
    assert np.max(np.abs(res["stderr"][ok, 0] - ora.stderr[ok]) / np.abs(ora.stderr[ok])) < 1e-6
    assert res["vgs"] == reml["vg"] and res["ves"] == reml["ve"] and res["REMLs"] == -2 * reml["REML"]
    assert np.allclose(res["effect.cv"], ora.effect_cv, rtol=1e-9)
    # duplicated adjacent marker: bit-identical row (the reference copies it, :512-527)
    for key in ("ps", "stats", "effect.est", "rsquare"):
        assert res[key][20, 0] == res[key][19, 0]
    # without CVI the reference's stderr index runs past beta -> NA (:972)
    res2 = gb().GAPIT_EMMAxP3D(Y, G.T, K, X0=X0, ctx=gpu_ctx, eig_L=eL, reml=reml)
    assert np.all(np.isnan(res2["stderr"]))


@pytest.mark.parametrize("slices", [5, 6, 7, 8])
def test_mlm_scan_slice_counts(gpu_ctx, slices):
    G, Y, X0 = case(260, 800, npc=1)
    K = go.kinship_vanraden(G.T)
    eL = go.emma_eigen_L_wo_Z(K)
    reml = go.emma_REMLE(Y, X0, K)
    ora = go.emmax_p3d(Y, G.T, K, X0, eig_L=eL, reml=reml)
    gpu_ctx.set_slices(slices)
    try:
        res = gb().GAPIT_EMMAxP3D(Y, G.T, K, X0=X0, ctx=gpu_ctx, eig_L=eL, reml=reml)
    finally:
        gpu_ctx.set_slices(0)
    r, a = rel_err(res["effect.est"][:, 0], ora.effect_est, 1e-3)
    assert r < BETA_RTOL   # 5 planes: ~1e-9 (tools/slice_accuracy.py); 6 and up: the oracle's own rounding
    with np.errstate(divide="ignore"):
        assert np.nanmax(np.abs(np.log10(res["ps"][:, 0]) - np.log10(ora.ps))) < LOGP_ATOL


@pytest.mark.parametrize("n,m,npc", [(48, 160, 2), (281, 3093, 0), (300, 1000, 5)])
def test_glm_scan_parity(gpu_ctx, n, m, npc):
    G, Y, X0 = case(n, m, npc=npc)
    ora = go.emmax_p3d(Y, G.T, None, X0, ncol_CVI=X0.shape[1])
    res = gb().GAPIT_EMMAxP3D(Y, G.T, None, X0=X0, CVI=np.zeros((n, X0.shape[1])), ctx=gpu_ctx)
    check_scan(res, ora, glm=True)
    assert res["ves"] == pytest.approx(ora.ves, rel=1e-11)
    assert res["REMLs"] == pytest.approx(ora.REMLs, rel=1e-11)
    assert res["vgs"] == 0.0


def test_end_to_end_own_spectrum_and_reml(gpu_ctx):
    """Whole function on the GPU side (own eigen + own REML) against the whole oracle.  delta comes out of a
    root finder with tolerance 1.2e-4 on both sides, so parity here is delta-agreement plus a looser scan bound."""
    G, Y, X0 = case(281, 3093, npc=0)
    K = gb().GAPIT_kinship_VanRaden(G.T, ctx=gpu_ctx, verbose=False)
    res = gb().GAPIT_EMMAxP3D(Y, G.T, K, X0=X0, ctx=gpu_ctx)
    ora = go.emmax_p3d(Y, G.T, go.kinship_vanraden(G.T), X0)
    assert res["delta"] == pytest.approx(ora.delta, rel=1e-6)
    assert res["vgs"] == pytest.approx(ora.vgs, rel=1e-6)
    ok = ~np.isnan(ora.stats)
    with np.errstate(divide="ignore"):
        assert np.nanmax(np.abs(np.log10(res["ps"][:, 0]) - np.log10(ora.ps))) < 1e-4
    r, a = rel_err(res["effect.est"][:, 0], ora.effect_est, 1e-3)
    assert r < 1e-5
    # output table order: sort by (P, marker index) as GAPIT.Perform.BH.FDR...R:46 does
    assert np.array_equal(np.argsort(res["ps"][ok, 0], kind="stable")[:50], np.argsort(ora.ps[ok], kind="stable")[:50])


def test_scan_golden_fixtures(gpu_ctx, tiny_golden, c1_golden):
    g = tiny_golden
    eL = {"values": g["eigL_values"], "vectors": g["eigL_vectors"]}
    reml = {"delta": float(g["delta"]), "vg": float(g["vg"]), "ve": float(g["ve"]), "REML": float(g["REML"])}
    res = gb().GAPIT_EMMAxP3D(g["Y"], g["G"].T, g["K"], X0=g["X0"], ctx=gpu_ctx, eig_L=eL, reml=reml)
    ok = ~np.isnan(g["mlm_stats"])
    assert np.allclose(res["effect.est"][ok, 0], g["mlm_effect"][ok], rtol=BETA_RTOL, atol=1e-12)
    assert np.nanmax(np.abs(np.log10(res["ps"][:, 0]) - np.log10(g["mlm_ps"]))) < LOGP_ATOL
    Kg, gram = gb().GAPIT_kinship_VanRaden(g["G"].T, ctx=gpu_ctx, return_gram=True, verbose=False)
    assert np.array_equal(gram, g["gram"]) and np.abs(Kg - g["K"]).max() <= K_RTOL * np.abs(g["K"]).max()
    # BASELINE configs[0] shape
    c = c1_golden
    G, Y, X0 = c1_inputs(c)
    Kc, gramc = gb().GAPIT_kinship_VanRaden(G.T, ctx=gpu_ctx, return_gram=True, verbose=False)
    assert np.array_equal(np.diag(gramc), c["gram_diag"]) and np.array_equal(gramc[0], c["gram_row0"])
    assert int(gramc.astype(np.int64).sum()) == int(c["gram_checksum"][0])
    assert np.allclose(np.diag(Kc), c["K_diag"], rtol=1e-12) and np.allclose(Kc[0], c["K_row0"], rtol=1e-9, atol=1e-13)
    glm = gb().GAPIT_EMMAxP3D(Y, G.T, None, X0=X0, ctx=gpu_ctx)
    with np.errstate(divide="ignore"):
        assert np.nanmax(np.abs(np.log10(glm["ps"][:, 0]) - np.log10(c["glm_ps"]))) < LOGP_ATOL
    mlm = gb().GAPIT_EMMAxP3D(Y, G.T, Kc, X0=X0, ctx=gpu_ctx)
    assert mlm["delta"] == pytest.approx(float(c["delta"]), rel=1e-6)
    with np.errstate(divide="ignore"):
        assert np.nanmax(np.abs(np.log10(mlm["ps"][:, 0]) - np.log10(c["mlm_ps"]))) < 1e-4


def test_scan_properties_at_scale(gpu_ctx):
    """n large enough for several eigen-index splits and marker tiles: duplicate columns give bit-identical rows
    wherever they sit, 0<->2 recoding flips the sign of beta only, a sub-range scan equals the same rows of the
    full scan (marker-shard independence), spot rows match the dense GLS solve."""
    n, m = 1500, 6000
    G = synth.make_genotypes(n, m, seed=5, threads=8)
    G[4999] = G[17]
    G[300] = G[17]
    Y = synth.make_phenotypes(G, seed=5)[:, 0]
    X0 = synth.make_covariates(G, npc=2)
    K = gb().GAPIT_kinship_VanRaden(G.T, ctx=gpu_ctx, verbose=False)
    eL = gb().emma_eigen_L_wo_Z(K, ctx=gpu_ctx)
    reml = gb().GAPIT_emma_REMLE(Y, X0, K, ctx=gpu_ctx)
    run = lambda snps: gb().GAPIT_EMMAxP3D(Y, snps, K, X0=X0, ctx=gpu_ctx, eig_L=eL, reml=reml)
    full = run(gb().Genotypes(G, ctx=gpu_ctx, marker_major=True))
    for key in ("ps", "stats", "effect.est", "rsquare", "logLM"):
        assert full[key][4999, 0] == full[key][17, 0] == full[key][300, 0]
    part = run(gb().Genotypes(G[2000:4500], ctx=gpu_ctx, marker_major=True))
    for key in ("ps", "stats", "effect.est"):
        assert np.array_equal(part[key][:, 0], full[key][2000:4500, 0], equal_nan=True)
    flip = run(gb().Genotypes((2 - G).astype(np.int8), ctx=gpu_ctx, marker_major=True))
    ok = ~np.isnan(full["stats"][:, 0])
    assert np.allclose(flip["effect.est"][ok, 0], -full["effect.est"][ok, 0], rtol=1e-7, atol=1e-10)
    assert np.allclose(flip["ps"][ok, 0], full["ps"][ok, 0], rtol=1e-6)
    for i in (0, 17, 1234, 5999):
        if np.isnan(full["stats"][i, 0]):
            continue
        d = cc.gls_dense(Y, X0, G[i].astype(float), K, reml["delta"], reml["vg"])
        assert full["effect.est"][i, 0] == pytest.approx(d["beta"][-1], rel=1e-7, abs=1e-10)
        assert full["stats"][i, 0] == pytest.approx(d["t"], rel=1e-7, abs=1e-10)


def test_multi_trait_matches_single_trait_calls(gpu_ctx):
    G, _, X0 = case(150, 400, npc=1)
    Yt = synth.make_phenotypes(G, ntraits=3, seed=8)
    K = go.kinship_vanraden(G.T)
    eL = go.emma_eigen_L_wo_Z(K)
    remls = [go.emma_REMLE(Yt[:, t], X0, K) for t in range(3)]
    geno = gb().Genotypes(G, ctx=gpu_ctx, marker_major=True)
    es = gb().EigenSystem(eL["values"], eL["vectors"], ctx=gpu_ctx)
    arrs, bases = gb().p3d_scan(geno, X0, Yt, eigsys=es, delta=[r["delta"] for r in remls], vg=[r["vg"] for r in remls],
                                ctx=gpu_ctx)
    for t in range(3):
        one, b1 = gb().p3d_scan(geno, X0, Yt[:, t], eigsys=es, delta=[remls[t]["delta"]], vg=[remls[t]["vg"]], ctx=gpu_ctx)
        # the single-trait launch reduces in registers, several traits go through the shared rotation (W^2 + back-rotated
        # columns): two evaluations of the same sums whose digit-plane rounding differs -- far below the oracle tolerance
        assert np.allclose(arrs["pvalue"][t], one["pvalue"][0], rtol=1e-8, atol=0, equal_nan=True)
        ora = go.emmax_p3d(Yt[:, t], G.T, K, X0, eig_L=eL, reml=remls[t])
        with np.errstate(divide="ignore"):
            assert np.nanmax(np.abs(np.log10(arrs["pvalue"][t]) - np.log10(ora.ps))) < LOGP_ATOL
        assert bases[t]["rsquare_base"] == pytest.approx(ora.rsquare_base[~np.isnan(ora.rsquare_base)][0], abs=1e-10)


def test_singular_covariates_fall_back_to_pseudo_inverse(gpu_ctx, capsys):
    G, Y, X0 = case(120, 200, npc=1)
    X0d = np.column_stack([X0, X0[:, 1] * 2.0])          # linearly dependent covariates (:708-712)
    K = go.kinship_vanraden(G.T)
    eL = go.emma_eigen_L_wo_Z(K)
    reml = go.emma_REMLE(Y, X0, K)
    res = gb().GAPIT_EMMAxP3D(Y, G.T, K, X0=X0d, ctx=gpu_ctx, eig_L=eL, reml=reml)
    assert "linearly dependent" in capsys.readouterr().out
    ref = gb().GAPIT_EMMAxP3D(Y, G.T, K, X0=X0, ctx=gpu_ctx, eig_L=eL, reml=reml)
    ok = ~np.isnan(ref["stats"][:, 0])
    # the marker effect is estimable either way
    assert np.allclose(res["effect.est"][ok, 0], ref["effect.est"][ok, 0], rtol=1e-6, atol=1e-9)


def test_bad_arguments_are_errors_not_crashes(gpu_ctx):
    G, Y, X0 = case(60, 80)
    geno = gb().Genotypes(G, ctx=gpu_ctx, marker_major=True)
    with pytest.raises(gb().GapitCudaError):
        gb().p3d_scan(geno, np.ones((60, 17)), Y, glm=True, ctx=gpu_ctx)        # q0 > GAPIT_MAX_Q0
    with pytest.raises(gb().GapitCudaError):
        gb().p3d_scan(geno, X0, Y, glm=False, ctx=gpu_ctx)                      # MLM without an eigen-system
    with pytest.raises(NotImplementedError):
        gb().GAPIT_EMMAxP3D(Y, G.T, None, X0=X0, ctx=gpu_ctx, SNP_P3D=False)


def test_streaming_scan_equals_resident_scan(gpu_ctx, monkeypatch):
    """gapit_p3d_scan_host (chunked upload on a copy stream) must reproduce the resident scan bit for bit,
    for every host dtype and with several chunks in flight."""
    monkeypatch.setenv("GAPIT_STREAM_CHUNK", "512")
    G, Y, X0 = case(200, 3000, npc=1)
    K = go.kinship_vanraden(G.T)
    eL = go.emma_eigen_L_wo_Z(K)
    reml = go.emma_REMLE(Y, X0, K)
    geno = gb().Genotypes(G, ctx=gpu_ctx, marker_major=True)
    es = gb().EigenSystem(eL["values"], eL["vectors"], ctx=gpu_ctx)
    kw = dict(eigsys=es, delta=[reml["delta"]], vg=[reml["vg"]], ctx=gpu_ctx)
    ref, bref = gb().p3d_scan(geno, X0, Y, **kw)
    for arr, mm in [(G, True), (G.T.astype(np.float64), False), (G.T.astype(np.int32), False)]:
        got, bgot = gb().p3d_scan_host(arr, X0, Y, marker_major=mm, **kw)
        for k in ref:
            assert np.array_equal(got[k], ref[k], equal_nan=True), k
        assert bgot[0]["rsquare_base"] == bref[0]["rsquare_base"]
    gref, _ = gb().p3d_scan(geno, X0, Y, glm=True, ctx=gpu_ctx)
    ggot, _ = gb().p3d_scan_host(G, X0, Y, glm=True, ctx=gpu_ctx, marker_major=True)
    for k in gref:
        assert np.array_equal(ggot[k], gref[k], equal_nan=True), k
    bad = G.T.astype(np.float64)
    bad[5, 2999] = 4.0
    with pytest.raises(gb().GapitCudaError):
        gb().p3d_scan_host(bad, X0, Y, **kw)
    na = G.T.astype(np.float64)
    na[5, 100] = np.nan
    with pytest.raises(gb().GapitCudaError):
        gb().p3d_scan_host(na, X0, Y, **kw)
    gb().p3d_scan_host(na, X0, Y, impute="Middle", **kw)


@pytest.mark.parametrize("npc,T", [(0, 18), (1, 14), (3, 9)])
def test_multi_trait_chunked_epilogue(gpu_ctx, npc, T):
    """BASELINE configs[4] shape in miniature: many traits share one eigendecomposition and one rotation pass per
    chunk of traits; every trait must match the oracle and the single-trait launch."""
    G, _, X0 = case(130, 520, npc=npc, seed=21)
    Yt = synth.make_phenotypes(G, ntraits=T, seed=21)
    K = go.kinship_vanraden(G.T)
    eL = go.emma_eigen_L_wo_Z(K)
    eR = go.emma_eigen_R_wo_Z(K, X0)
    remls = [go.emma_REMLE(Yt[:, t], X0, K, eig_R=eR) for t in range(T)]
    geno = gb().Genotypes(G, ctx=gpu_ctx, marker_major=True)
    es = gb().EigenSystem(eL["values"], eL["vectors"], ctx=gpu_ctx)
    arrs, bases = gb().p3d_scan(geno, X0, Yt, eigsys=es, delta=[r["delta"] for r in remls], vg=[r["vg"] for r in remls],
                                ctx=gpu_ctx)
    host, _ = gb().p3d_scan_host(G, X0, Yt, eigsys=es, delta=[r["delta"] for r in remls], vg=[r["vg"] for r in remls],
                                 ctx=gpu_ctx, marker_major=True)
    for t in (0, 1, T // 2, T - 1):
        ora = go.emmax_p3d(Yt[:, t], G.T, K, X0, eig_L=eL, reml=remls[t])
        ok = ~np.isnan(ora.stats)
        r, a = rel_err(arrs["effect"][t], ora.effect_est, 1e-4)
        assert r < BETA_RTOL, (t, r)
        with np.errstate(divide="ignore"):
            assert np.nanmax(np.abs(np.log10(arrs["pvalue"][t]) - np.log10(ora.ps))) < LOGP_ATOL
        assert np.max(np.abs(arrs["rsquare"][t][ok] - ora.rsquare[ok])) < 1e-9
        assert np.array_equal(host["pvalue"][t], arrs["pvalue"][t], equal_nan=True)


@pytest.mark.parametrize("npc,ridge", [(0, 0.0), (2, 0.0), (1, 0.05)])
def test_blup_pev_block(gpu_ctx, npc, ridge):
    """i = 0 block (GAPIT.EMMAxP3D.R:779-861): eigenbasis evaluation against the dense restatement with the
    reference's two n x n inversions; singular K (ridge = 0: pseudo-inverse branch) and regular K."""
    G, Y, X0 = case(160, 900, npc=npc, seed=31)
    K = go.kinship_vanraden(G.T) + ridge * np.eye(160)
    eL = go.emma_eigen_L_wo_Z(K)
    reml = go.emma_REMLE(Y, X0, K)
    ora = go.blup_pev(Y, X0, K, eL, reml, CVI=np.column_stack([np.zeros(160), X0[:, 1:]]))
    assert ora["used_ginv_K"] == (ridge == 0.0)
    res = gb().GAPIT_EMMAxP3D(Y, G.T, K, X0=X0, CVI=np.column_stack([np.zeros(160), X0[:, 1:]]), ctx=gpu_ctx,
                              eig_L=eL, reml=reml)
    assert np.max(np.abs(res["BLUP"][:, 0] - ora["BLUP"])) < 1e-9 * max(1.0, np.abs(ora["BLUP"]).max())
    assert np.max(np.abs(res["PEV"][:, 0] - ora["PEV"])) < 1e-8 * np.abs(ora["PEV"]).max()
    assert np.allclose(res["BLUP_Plus_Mean"][:, 0], ora["BLUP_Plus_Mean"], rtol=1e-9, atol=1e-9)
    assert np.allclose(res["BLUE"][:, 0], ora["BLUE"], rtol=1e-9, atol=1e-9)
    glm = gb().GAPIT_EMMAxP3D(Y, G.T, None, X0=X0, ctx=gpu_ctx)
    assert glm["BLUP"] == 0.0 and glm["PEV"] == glm["ves"] and np.isnan(glm["BLUP_Plus_Mean"])   # :872-875


def test_full_size_c2_properties(gpu_ctx):
    """BASELINE configs[1] at full size (5,000 taxa x 500,000 SNPs): the oracle cannot run this in seconds, so parity
    is carried by size-independent properties -- exact Gram checksums, K 1 = 0, bit-identical rows for duplicated
    columns in different tiles, shard independence of the scan, and spot rows against the dense Cholesky GLS."""
    n, m = 5000, 500000
    G = synth.make_genotypes(n, m, seed=20160302, threads=16)
    G[499999] = G[123]          # duplicates far apart (different tiles, different stream chunks)
    G[250000] = G[123]
    Y = synth.make_phenotypes(G[:4096], seed=3)[:, 0]
    X0 = np.ones((n, 1))
    geno = gb().Genotypes(G, ctx=gpu_ctx, marker_major=True)
    K, gram = gb().GAPIT_kinship_VanRaden(geno, ctx=gpu_ctx, return_gram=True, verbose=False)
    cs, cq = geno.colsums()
    c = cs.astype(np.int64)
    assert int(gram.astype(np.int64).sum()) == int((c * c).sum())                 # sum_ij G_ij = sum_k c_k^2
    assert int(np.trace(gram.astype(np.int64))) == int(cq.astype(np.int64).sum())  # tr G = sum x^2
    sub = G[:, 4900:5000].astype(np.int64)
    assert np.array_equal(gram[4900:, 4900:], sub.T @ sub)
    assert np.abs(K.sum(1)).max() < 1e-8 and np.array_equal(K, K.T)
    eL = gb().emma_eigen_L_wo_Z(K, ctx=gpu_ctx)
    reml = gb().GAPIT_emma_REMLE(Y, X0, K, ctx=gpu_ctx)
    es = gb().EigenSystem(eL["values"], eL["vectors"], ctx=gpu_ctx)
    kw = dict(eigsys=es, delta=[reml["delta"]], vg=[reml["vg"]], ctx=gpu_ctx)
    full, _ = gb().p3d_scan(geno, X0, Y, **kw)
    for key in ("pvalue", "tvalue", "effect", "rsquare", "loglm"):
        assert full[key][0, 499999] == full[key][0, 123] == full[key][0, 250000], key
    # a marker shard (what one rank of a multi-GPU run holds) reproduces its rows of the full scan bit for bit
    shard = gb().Genotypes(G[300000:337000], ctx=gpu_ctx, marker_major=True)
    part, _ = gb().p3d_scan(shard, X0, Y, **kw)
    for key in ("pvalue", "tvalue", "effect"):
        assert np.array_equal(part[key][0], full[key][0, 300000:337000], equal_nan=True), key
    # the streaming host path too
    host, _ = gb().p3d_scan_host(G, X0, Y, marker_major=True, **kw)
    assert np.array_equal(host["pvalue"], full["pvalue"], equal_nan=True)
    # three blocks of 1,024 markers (first, middle, last tiles) against the batched fp64 derivation (oracle/batched.py:
    # one dgemm V'X + the closed-form bordered solve -- independent of the per-marker loop and of the int8 split) at
    # north_star's tolerances, plus spot rows against the dense Cholesky GLS that never touches the eigen route
    from oracle import batched
    for k0 in (0, 249856, m - 1024):
        ref = batched.p3d_block(Y, G[k0:k0 + 1024].T, X0, eL["values"], eL["vectors"], reml["delta"], reml["vg"])
        c = batched.compare({k: full[k][0] for k in ("effect", "se", "tvalue", "pvalue")}, ref, rows=slice(k0, k0 + 1024))
        assert c["nan_rows_identical"] and c["n_tested"] > 1000, c
        assert c["beta_rel_max"] < BETA_RTOL and c["se_rel_max"] < BETA_RTOL and c["dlog10p_max"] < LOGP_ATOL, c
    L = np.linalg.cholesky(K + reml["delta"] * np.eye(n))
    yw = np.linalg.solve(L, Y)
    x0w = np.linalg.solve(L, X0[:, 0])
    for i in (0, 123, 77777, 499998):
        if np.isnan(full["tvalue"][0, i]):
            continue
        Xw = np.column_stack([x0w, np.linalg.solve(L, G[i].astype(np.float64))])
        iXX = np.linalg.inv(Xw.T @ Xw)
        b = iXX @ (Xw.T @ yw)
        se = np.sqrt(iXX[1, 1] * reml["vg"])
        assert full["effect"][0, i] == pytest.approx(b[1], rel=1e-7, abs=1e-10)
        assert full["se"][0, i] == pytest.approx(se, rel=1e-8)
    gref = batched.p3d_block(Y, G[:2048].T, X0, glm=True)
    # GLM rows against OLS with the reduced-model variance (:937)
    glm, gbase = gb().p3d_scan(geno, X0, Y, glm=True, ctx=gpu_ctx)
    ves = float(((Y - Y.mean()) ** 2).sum() / (n - 1))
    assert gbase[0]["ves"] == pytest.approx(ves, rel=1e-10)
    c = batched.compare({k: glm[k][0] for k in ("effect", "se", "tvalue", "pvalue")}, gref, rows=slice(0, 2048))
    assert c["nan_rows_identical"] and c["beta_rel_max"] < BETA_RTOL and c["se_rel_max"] < BETA_RTOL and c["dlog10p_max"] < LOGP_ATOL, c
    for i in (5, 400000):
        x = G[i].astype(np.float64)
        if x.min() == x.max():
            continue
        X = np.column_stack([X0[:, 0], x])
        iXX = np.linalg.inv(X.T @ X)
        b = iXX @ (X.T @ Y)
        assert glm["tvalue"][0, i] == pytest.approx(b[1] / np.sqrt(iXX[1, 1] * ves), rel=1e-9)


def test_large_taxa_shapes(gpu_ctx):
    """Taxa counts of BASELINE configs[2] (20,000) and configs[3] (50,000) on marker subsets one GPU reaches in
    seconds: exercises the eigen-index splits (gsplit = 16), the banded Gram walk and the 50,176-wide Gram buffer.
    Checked through exact checksums, symmetry, spot blocks, duplicate-row identity and shard independence."""
    import torch
    from gapit_b200 import _lib
    # ---- configs[3]: kinship only, 50,000 taxa
    n, m = 50000, 12000
    G = synth.make_genotypes(n, m, seed=50, threads=16)
    geno = gb().Genotypes(G, ctx=gpu_ctx, marker_major=True)
    lib = gpu_ctx.lib
    ldg = lib.gapit_gram_ld(n)
    G_t = torch.empty((ldg, ldg), dtype=torch.int32, device="cuda")
    sums = torch.empty(3, dtype=torch.int64, device="cuda")
    K_t = torch.empty((n, n), dtype=torch.float64, device="cuda")
    _lib.check(lib.gapit_vanraden_gram(gpu_ctx.h, geno.h, G_t.data_ptr(), sums.data_ptr()))
    _lib.check(lib.gapit_vanraden_finish(gpu_ctx.h, n, G_t.data_ptr(), sums.data_ptr(), K_t.data_ptr(), None, None))
    cs, cq = geno.colsums()
    c = cs.astype(np.int64)
    Gl = torch.tril(G_t[:n, :n]).to(torch.int64)
    dg = torch.diagonal(G_t[:n, :n]).to(torch.int64).sum().item()
    assert 2 * int(Gl.sum().item()) - int(dg) == int((c * c).sum())               # sum over the full symmetric matrix
    del Gl
    assert int(dg) == int(cq.astype(np.int64).sum())
    for lo in (0, 24900, 49700):                     # first, middle and last tile bands
        sub = G[:, lo:lo + 300].astype(np.int64)
        assert np.array_equal(np.tril(G_t[lo:lo + 300, lo:lo + 300].cpu().numpy()), np.tril(sub.T @ sub))
    a = G[:, 100:200].astype(np.int64)
    b = G[:, 49000:49100].astype(np.int64)
    assert np.array_equal(G_t[49000:49100, 100:200].cpu().numpy(), b.T @ a)      # lower triangle (the upper one is not formed)
    Ksub = K_t[49000:49100, 100:200].cpu().numpy()                                # K is written in both orientations
    assert np.array_equal(Ksub, K_t[100:200, 49000:49100].cpu().numpy().T)
    assert K_t.sum(dim=1).abs().max().item() < 1e-7
    del G_t, K_t, geno
    torch.cuda.empty_cache()
    # ---- configs[2]: one rank's view, 20,000 taxa
    n, m = 20000, 8192
    G = synth.make_genotypes(n, m, seed=51, threads=16)
    G[8000] = G[77]
    Y = synth.make_phenotypes(G[:2048], seed=4)[:, 0]
    X0 = np.ones((n, 1))
    geno = gb().Genotypes(G, ctx=gpu_ctx, marker_major=True)
    K = gb().GAPIT_kinship_VanRaden(geno, ctx=gpu_ctx, verbose=False)   # undoctored: rank <= 8,192, K 1 = 0
    eL = gb().emma_eigen_L_wo_Z(K, ctx=gpu_ctx)
    assert np.all(np.diff(eL["values"]) <= 1e-9)
    reml = gb().GAPIT_emma_REMLE(Y, X0, K, eig_L=eL, ctx=gpu_ctx)        # the REML delta of this very kinship
    assert reml["delta"] > 0 and np.isfinite(reml["vg"])
    es = gb().EigenSystem(eL["values"], eL["vectors"], ctx=gpu_ctx)
    kw = dict(eigsys=es, delta=[reml["delta"]], vg=[reml["vg"]], ctx=gpu_ctx)
    full, _ = gb().p3d_scan(geno, X0, Y, **kw)
    assert full["pvalue"][0, 8000] == full["pvalue"][0, 77] and full["effect"][0, 8000] == full["effect"][0, 77]
    part, _ = gb().p3d_scan(gb().Genotypes(G[3000:5000], ctx=gpu_ctx, marker_major=True), X0, Y, **kw)
    assert np.array_equal(part["pvalue"][0], full["pvalue"][0, 3000:5000], equal_nan=True)
    # 2 x 1,024 markers against the batched fp64 derivation at north_star's tolerances (beta / SE 1e-8, -log10 p 1e-6)
    from oracle import batched
    for k0 in (0, m - 1024):
        ref = batched.p3d_block(Y, G[k0:k0 + 1024].T, X0, eL["values"], eL["vectors"], reml["delta"], reml["vg"])
        c = batched.compare({k: full[k][0] for k in ("effect", "se", "tvalue", "pvalue")}, ref, rows=slice(k0, k0 + 1024))
        assert c["nan_rows_identical"] and c["n_tested"] > 1000, c
        assert c["beta_rel_max"] < BETA_RTOL and c["se_rel_max"] < BETA_RTOL and c["dlog10p_max"] < LOGP_ATOL, c
    # three traits through the shared rotation at this taxa count: every trait against its own single-trait launch
    Y3 = synth.make_phenotypes(G[:2048], ntraits=3, seed=4)
    r3 = [gb().GAPIT_emma_REMLE(Y3[:, t], X0, K, eig_L=eL, ctx=gpu_ctx) for t in range(3)]
    multi, _ = gb().p3d_scan(geno, X0, Y3, eigsys=es, delta=[r["delta"] for r in r3], vg=[r["vg"] for r in r3], ctx=gpu_ctx)
    for t in (0, 2):
        one, _ = gb().p3d_scan(geno, X0, Y3[:, t], eigsys=es, delta=[r3[t]["delta"]], vg=[r3[t]["vg"]], ctx=gpu_ctx)
        ok = ~np.isnan(one["tvalue"][0])
        assert np.array_equal(np.isnan(multi["tvalue"][t]), ~ok)
        assert np.max(np.abs(multi["se"][t][ok] - one["se"][0][ok]) / one["se"][0][ok]) < 1e-10
        assert np.max(np.abs(multi["effect"][t][ok] - one["effect"][0][ok]) / one["se"][0][ok]) < 1e-9   # in SE units
        with np.errstate(divide="ignore"):
            assert np.nanmax(np.abs(np.log10(multi["pvalue"][t][ok]) - np.log10(one["pvalue"][0][ok]))) < 1e-7


@pytest.mark.parametrize("T", [1, 4])
def test_unmerged_digit_epilogue_of_large_taxa_counts(gpu_ctx, monkeypatch, T):
    """n >= 32,640 cannot merge adjacent digit planes in int32 (257 * 256 * n overflows) and takes the unmerged
    epilogue; GAPIT_EPI_PAIR=0 forces that path at a size the oracle handles, for the fused single-trait epilogue
    (T = 1) and the store epilogue of the shared rotation (T = 4)."""
    G, _, X0 = case(333, 900, npc=2, seed=17)
    Yt = synth.make_phenotypes(G, ntraits=T, seed=17)
    K = go.kinship_vanraden(G.T)
    eL = go.emma_eigen_L_wo_Z(K)
    remls = [go.emma_REMLE(Yt[:, t], X0, K) for t in range(T)]
    geno = gb().Genotypes(G, ctx=gpu_ctx, marker_major=True)
    es = gb().EigenSystem(eL["values"], eL["vectors"], ctx=gpu_ctx)
    kw = dict(eigsys=es, delta=[r["delta"] for r in remls], vg=[r["vg"] for r in remls], ctx=gpu_ctx)
    merged, _ = gb().p3d_scan(geno, X0, Yt, **kw)
    monkeypatch.setenv("GAPIT_EPI_PAIR", "0")
    plain, _ = gb().p3d_scan(geno, X0, Yt, **kw)
    for t in range(T):
        ora = go.emmax_p3d(Yt[:, t], G.T, K, X0, eig_L=eL, reml=remls[t])
        ok = ~np.isnan(ora.stats)
        assert np.array_equal(np.isnan(plain["tvalue"][t]), ~ok)
        r, a = rel_err(plain["effect"][t], ora.effect_est, 1e-4)
        assert r < BETA_RTOL, (t, r)
        r, a = rel_err(plain["se"][t], ora.se_clean, 0.0)
        assert r < SE_RTOL, (t, r)
        with np.errstate(divide="ignore"):
            assert np.nanmax(np.abs(np.log10(plain["pvalue"][t]) - np.log10(ora.ps))) < LOGP_ATOL
        # the integer digit sums are exact either way: merging only changes how they are converted
        assert np.allclose(plain["effect"][t], merged["effect"][t], rtol=1e-12, atol=0, equal_nan=True)


def test_shared_rotation_blocks_do_not_change_results(gpu_ctx, monkeypatch):
    """The shared-rotation form of the multi-trait scan (StoreEpi + xx_contract) works through marker blocks sized from
    the free device memory; forcing several small blocks must give bit-identical rows (each marker's sums are formed
    in a fixed order of the eigen index wherever its tile sits)."""
    G, _, X0 = case(140, 1500, npc=2, seed=5)
    T = 5
    Yt = synth.make_phenotypes(G, ntraits=T, seed=5)
    K = go.kinship_vanraden(G.T)
    eL = go.emma_eigen_L_wo_Z(K)
    remls = [go.emma_REMLE(Yt[:, t], X0, K) for t in range(T)]
    geno = gb().Genotypes(G, ctx=gpu_ctx, marker_major=True)
    es = gb().EigenSystem(eL["values"], eL["vectors"], ctx=gpu_ctx)
    kw = dict(eigsys=es, delta=[r["delta"] for r in remls], vg=[r["vg"] for r in remls], ctx=gpu_ctx)
    whole, _ = gb().p3d_scan(geno, X0, Yt, **kw)
    monkeypatch.setenv("GAPIT_SHARED_BLOCK", "512")
    parts, _ = gb().p3d_scan(geno, X0, Yt, **kw)
    for key in ("effect", "tvalue", "se", "pvalue", "rsquare", "loglm"):
        assert np.array_equal(whole[key], parts[key], equal_nan=True), key
    assert np.array_equal(whole["maf"], parts["maf"])
    for t in range(T):
        ora = go.emmax_p3d(Yt[:, t], G.T, K, X0, eig_L=eL, reml=remls[t])
        r, a = rel_err(parts["effect"][t], ora.effect_est, 1e-4)
        assert r < BETA_RTOL, (t, r)
        with np.errstate(divide="ignore"):
            assert np.nanmax(np.abs(np.log10(parts["pvalue"][t]) - np.log10(ora.ps))) < LOGP_ATOL


# ------------------------------------------------------------------ PLINK .bed and numeric text
def _write_bed(G_na, count="A1"):
    """Independent encoder: (n, m) dosages of `count` (-1 = missing) -> bytes of a variant-major .bed."""
    n, m = G_na.shape
    code = {"A1": {2: 0, -1: 1, 1: 2, 0: 3}, "A2": {0: 0, -1: 1, 1: 2, 2: 3}}[count]
    nb = (n + 3) // 4
    out = bytearray([0x6C, 0x1B, 0x01])
    for k in range(m):
        row = bytearray(nb)
        for i in range(n):
            row[i // 4] |= code[int(G_na[i, k])] << (2 * (i % 4))
        out += row
    return bytes(out)


@pytest.mark.parametrize("n,m,count", [(5, 7, "A1"), (131, 97, "A2"), (600, 350, "A1")])
def test_plink_bed_rows_go_to_the_gpu_as_they_are(gpu_ctx, tmp_path, n, m, count):
    rng = np.random.default_rng(n + m)
    G = rng.integers(0, 3, size=(n, m)).astype(np.int8)
    G_na = G.copy()
    G_na[rng.random((n, m)) < 0.02] = -1
    G_na[n // 2, m // 2] = -1                                                    # at least one missing call
    bed = _write_bed(G_na, count)
    assert np.array_equal(go.read_bed(bed, n, count), G_na)                     # the oracle against the encoder
    rows = gb().Packed2.from_bed(bed, n, count=count)
    assert rows.m == m
    with pytest.raises(gb().GapitCudaError):                                     # missing calls, no impute rule
        gb().Genotypes(rows, ctx=gpu_ctx)
    for rule, val in (("Middle", 1), ("Major", 2), ("Minor", 0)):
        g = gb().Genotypes(rows, ctx=gpu_ctx, impute=rule)
        want = np.where(G_na < 0, val, G_na).astype(np.int64)
        cs, cq = g.colsums()
        assert np.array_equal(cs, want.sum(0)) and np.array_equal(cq, (want * want).sum(0))
        K, gram = gb().GAPIT_kinship_VanRaden(g, ctx=gpu_ctx, return_gram=True, verbose=False)
        assert np.array_equal(gram, want @ want.T)
    with pytest.raises(gb().GapitCudaError):
        gb().Packed2.from_bed(b"\x6c\x1b\x00" + bed[3:], n)                     # sample-major mode byte
    with pytest.raises(gb().GapitCudaError):
        gb().Packed2.from_bed(b"\x00\x1b\x01" + bed[3:], n)                     # wrong magic
    # files on disk: .fam / .bim parsed on the host, the .bed memory-mapped and streamed through the scan
    prefix = str(tmp_path / "geno")
    open(prefix + ".bed", "wb").write(bed)
    open(prefix + ".fam", "w").write("".join(f"F{i} T{i} 0 0 0 -9\n" for i in range(n)))
    open(prefix + ".bim", "w").write("".join(f"1\tS{k}\t0\t{100 + k}\tA\tG\n" for k in range(m)))
    rows2, GT, GI, alleles = gb().read_plink(prefix, count=count)
    assert GT[:2] == ["T0", "T1"] and GI["SNP"][-1] == f"S{m - 1}" and GI["Position"][0] == 100 and alleles[0] == ("A", "G")
    if n > 100:
        Y = np.random.default_rng(1).normal(size=n)
        X0 = np.ones((n, 1))
        want = np.where(G_na < 0, 1, G_na).astype(np.int8)
        ref, _ = gb().p3d_scan_host(want.T.copy(), X0, Y, glm=True, ctx=gpu_ctx, marker_major=True)
        got, _, _, _ = gb().p3d_scan_file(prefix, X0, Y, glm=True, ctx=gpu_ctx, count=count)
        for key in ("effect", "pvalue", "maf"):
            assert np.array_equal(got[key], ref[key], equal_nan=True), key


@pytest.mark.parametrize("n,m", [(3, 5), (70, 333), (300, 5000)])
def test_numeric_text_table_is_parsed_on_the_gpu(gpu_ctx, n, m):
    """genoFormat = "EMMA" (GAPIT.Fragment.R:94-136): file.GD with a heading row and one row per taxon."""
    rng = np.random.default_rng(7 * n + m)
    G = rng.integers(0, 3, size=(n, m))
    na = rng.random((n, m)) < 0.01
    seps = ["\t", " ", ",", "  "]
    lines = ["taxa\t" + "\t".join(f"m{k}" for k in range(m))]
    for i in range(n):
        sep = seps[i % 4]
        toks = []
        for k in range(m):
            if na[i, k]:
                toks.append(["NA", "NaN", "."][k % 3])
            else:
                toks.append(str(G[i, k]) if (i + k) % 5 else f"{G[i, k]}.0")
        lines.append(f"line_{i:04d}{sep}" + sep.join(toks))
    text = ("\r\n" if n % 2 else "\n").join(lines).encode() + (b"\n" if n % 3 else b"")
    GD_o, GT_o, names_o = go.read_numeric_text(text)
    assert np.array_equal(np.isnan(GD_o), na)
    gd, GT, names, _ = gb().GAPIT_numeric_text(text, SNP_impute="Middle", ctx=gpu_ctx)
    assert GT == GT_o and names == names_o
    assert np.array_equal(gd, np.where(na, 1, G))
    if na.any():
        with pytest.raises(gb().GapitCudaError):                                 # NA and no impute rule
            gb().GAPIT_numeric_text(text, SNP_impute=None, ctx=gpu_ctx)
    bad = text.replace(b"line_0001", b"line_0001 2", 1) if n > 1 else text + b"x 1\n"
    with pytest.raises(gb().GapitCudaError):                                     # ragged row
        gb().GAPIT_numeric_text(bad, SNP_impute="Middle", ctx=gpu_ctx)
    with pytest.raises(gb().GapitCudaError):                                     # a value that is no dosage
        gb().GAPIT_numeric_text(text.rstrip(b"\r\n") + b"\nextra" + b"\t7" * m + b"\n", SNP_impute="Middle", ctx=gpu_ctx)
    if n >= 70:   # the parsed table as a device handle: scan it, compare with the scan of the same dosages from int8
        Y = rng.normal(size=n)
        X0 = np.ones((n, 1))
        want = np.where(na, 1, G).astype(np.int8)
        ref, _ = gb().p3d_scan_host(want.T.copy(), X0, Y, glm=True, ctx=gpu_ctx, marker_major=True)
        got, _, GT2, names2 = gb().p3d_scan_file(text, X0, Y, glm=True, ctx=gpu_ctx, genoFormat="EMMA")
        assert GT2 == GT_o and names2 == names_o
        for key in ("effect", "pvalue", "maf"):
            assert np.array_equal(got[key], ref[key], equal_nan=True), key


def test_emmaxp3d_with_several_rows_of_ys_shares_one_rotation(gpu_ctx):
    """GAPIT.EMMAxP3D loops the rows of ys (`for (j in 1:g)`, :81) and fills m x g matrices; the drop-in runs all g traits
    through one eigendecomposition and one rotation.  Every column must equal the one-trait call of that row (and the
    oracle), the scalar entries are those of the last trait (what the reference's loop leaves behind)."""
    G, _, X0 = case(170, 700, npc=2, seed=41)
    g = 4
    Yt = synth.make_phenotypes(G, ntraits=g, seed=41)
    K = go.kinship_vanraden(G.T)
    eL = go.emma_eigen_L_wo_Z(K)
    CVI = np.column_stack([np.zeros(170), X0[:, 1:]])
    many = gb().GAPIT_EMMAxP3D(Yt.T, G.T, K, X0=X0, CVI=CVI, ctx=gpu_ctx, eig_L=eL)
    assert many["ps"].shape == (700, g) and many["maf"].shape == (700, g)
    for t in range(g):
        one = gb().GAPIT_EMMAxP3D(Yt[:, t], G.T, K, X0=X0, CVI=CVI, ctx=gpu_ctx, eig_L=eL)
        ok = ~np.isnan(one["stats"][:, 0])
        assert np.array_equal(np.isnan(many["stats"][:, t]), ~ok)
        assert np.allclose(many["effect.est"][ok, t], one["effect.est"][ok, 0], rtol=0, atol=1e-9 * np.abs(one["se"][ok, 0]).max())
        assert np.allclose(many["se"][ok, t], one["se"][ok, 0], rtol=1e-10, atol=0)
        with np.errstate(divide="ignore"):
            assert np.nanmax(np.abs(np.log10(many["ps"][:, t]) - np.log10(one["ps"][:, 0]))) < 1e-7
        assert np.allclose(many["rsquare_base"][ok, t], one["rsquare_base"][ok, 0], rtol=0, atol=1e-12)
        reml = go.emma_REMLE(Yt[:, t], X0, K)
        ora = go.emmax_p3d(Yt[:, t], G.T, K, X0, eig_L=eL, reml=reml)
        with np.errstate(divide="ignore"):
            assert np.nanmax(np.abs(np.log10(many["ps"][:, t]) - np.log10(ora.ps))) < 1e-4   # own REML vs the oracle's
    last = gb().GAPIT_EMMAxP3D(Yt[:, g - 1], G.T, K, X0=X0, CVI=CVI, ctx=gpu_ctx, eig_L=eL)
    assert many["vgs"] == pytest.approx(last["vgs"], rel=1e-12) and many["REMLs"] == pytest.approx(last["REMLs"], rel=1e-12)
    assert np.allclose(many["BLUP"], last["BLUP"], rtol=1e-9, atol=1e-12) and np.allclose(many["effect.cv"], last["effect.cv"])


def test_eigh_mg_entry_agrees_with_the_one_gpu_solver(gpu_ctx):
    """gapit_eigh_mg (eigen(K) of the R shim): on one GPU -- or in a process that cannot load the multi-GPU solver, like
    this one, which carries PyTorch's libcublas -- it must fall through to the one-GPU syevd and give the same answer."""
    import ctypes as C
    from gapit_b200 import _lib
    rng = np.random.default_rng(3)
    n = 2500                                     # above the 2,048 threshold of the multi-GPU route
    A = rng.normal(size=(n, n + 50))
    K = np.asfortranarray(A @ A.T / (n + 50))
    P = lambda a: a.ctypes.data_as(C.c_void_p)
    v1, V1 = np.empty(n), np.empty((n, n), order="F")
    v2, V2 = np.empty(n), np.empty((n, n), order="F")
    _lib.check(gpu_ctx.lib.gapit_eigh(gpu_ctx.h, P(K), n, P(v1), P(V1)))
    _lib.check(gpu_ctx.lib.gapit_eigh_mg(gpu_ctx.h, P(K), n, P(v2), P(V2), 0))
    assert np.all(np.diff(v2) <= 1e-12) and np.allclose(v1, v2, rtol=0, atol=1e-11 * v1[0])
    idx = [0, 1, n // 2, n - 1]
    assert np.abs(K @ V2[:, idx] - V2[:, idx] * v2[idx]).max() < 1e-10 * v1[0]


def test_by_file_scans_reorder_taxa_with_gtindex(gpu_ctx, tmp_path):
    """`xs = myFRG$GD[GTindex,]` (GAPIT.EMMAxP3D.R:315): the genotype file holds more taxa than were phenotyped, in its
    own order; GTindex brings them into phenotype order.  Applied on the device for HapMap text, numeric text and .bed."""
    rng = np.random.default_rng(77)
    G, _, _ = case(150, 500, npc=0, edge=False)                  # the file: 150 taxa
    idx = rng.permutation(150)[:110]                              # 110 phenotyped taxa, shuffled
    Y = rng.normal(size=110)
    X0 = np.column_stack([np.ones(110), rng.normal(size=110)])
    geno = gb().Genotypes(G, ctx=gpu_ctx, marker_major=True)
    sel = geno.select_taxa(idx)
    cs, cq = sel.colsums()
    want = G[:, idx].astype(np.int64)
    assert sel.n == 110 and np.array_equal(cs, want.sum(1)) and np.array_equal(cq, (want * want).sum(1))
    with pytest.raises(gb().GapitCudaError):
        geno.select_taxa([0, 150])
    ref, _ = gb().p3d_scan(sel, X0, Y, glm=True, ctx=gpu_ctx)
    # HapMap text without missing calls: the file's dosages are G up to the per-marker allele orientation, so compare
    # against the same text scanned through a host-side reordering of its own numericalization
    text = synth.make_hapmap_text(G, bit=2, missing=0.0)
    GD = go.hapmap(synth.hapmap_rows(text), SNP_impute="Middle")["GD"]           # 150 x 500
    href, _ = gb().p3d_scan_host(np.ascontiguousarray(GD[idx].T.astype(np.int8)), X0, Y, glm=True, ctx=gpu_ctx, marker_major=True)
    hgot, _, GT, _ = gb().p3d_scan_hapmap(text, X0, Y, glm=True, file_fragment=128, ctx=gpu_ctx, GTindex=idx)
    assert len(GT) == 150
    for key in ("effect", "pvalue", "maf"):
        assert np.array_equal(hgot[key], href[key], equal_nan=True), key
    # numeric text and .bed carry G itself
    lines = ["taxa " + " ".join(f"m{k}" for k in range(500))] + [f"t{i} " + " ".join(map(str, G[:, i])) for i in range(150)]
    ngot, _, GT2, _ = gb().p3d_scan_file("\n".join(lines).encode(), X0, Y, glm=True, ctx=gpu_ctx, genoFormat="EMMA", GTindex=idx)
    bed = _write_bed(G.T.astype(np.int8), "A1")
    prefix = str(tmp_path / "g")
    open(prefix + ".bed", "wb").write(bed)
    open(prefix + ".fam", "w").write("".join(f"F t{i} 0 0 0 -9\n" for i in range(150)))
    open(prefix + ".bim", "w").write("".join(f"1 m{k} 0 {k + 1} A C\n" for k in range(500)))
    bgot, _, GT3, _ = gb().p3d_scan_file(prefix, X0, Y, glm=True, ctx=gpu_ctx, GTindex=idx)
    assert GT2[:3] == ["t0", "t1", "t2"] and GT3 == GT2
    for key in ("effect", "pvalue", "maf"):
        assert np.array_equal(ngot[key], ref[key], equal_nan=True), key
        assert np.array_equal(bgot[key], ref[key], equal_nan=True), key


@pytest.mark.parametrize("n,m,npc,T", [(60, 200, 0, 3), (90, 256, 5, 2), (120, 700, 1, 130), (64, 300, 15, 9)])
def test_shared_rotation_edge_shapes(gpu_ctx, n, m, npc, T):
    """Corners of the multi-trait path: a single 256-marker tile (the one-CTA form of the main loop with the store
    epilogue), more than 128 traits (two trait tiles of the DMMA contraction), the maximum of 16 covariate columns
    (many back-rotated columns), n below one operand tile."""
    G, _, X0 = case(n, m, npc=npc, seed=100 + T)
    Yt = synth.make_phenotypes(G, ntraits=T, seed=100 + T)
    K = go.kinship_vanraden(G.T)
    eL = go.emma_eigen_L_wo_Z(K)
    eR = go.emma_eigen_R_wo_Z(K, X0)
    check_t = sorted({0, T // 2, T - 1})
    remls = [go.emma_REMLE(Yt[:, t], X0, K, eig_R=eR) for t in range(T)]
    geno = gb().Genotypes(G, ctx=gpu_ctx, marker_major=True)
    es = gb().EigenSystem(eL["values"], eL["vectors"], ctx=gpu_ctx)
    arrs, bases = gb().p3d_scan(geno, X0, Yt, eigsys=es, delta=[r["delta"] for r in remls], vg=[r["vg"] for r in remls],
                                ctx=gpu_ctx)
    for t in check_t:
        ora = go.emmax_p3d(Yt[:, t], G.T, K, X0, eig_L=eL, reml=remls[t])
        ok = ~np.isnan(ora.stats)
        assert np.array_equal(np.isnan(arrs["tvalue"][t]), ~ok)
        r, a = rel_err(arrs["effect"][t], ora.effect_est, 1e-4)
        assert r < BETA_RTOL, (t, r)
        r, a = rel_err(arrs["se"][t], ora.se_clean, 0.0)
        assert r < SE_RTOL, (t, r)
        with np.errstate(divide="ignore"):
            assert np.nanmax(np.abs(np.log10(arrs["pvalue"][t]) - np.log10(ora.ps))) < LOGP_ATOL
        assert np.max(np.abs(arrs["rsquare"][t][ok] - ora.rsquare[ok])) < 1e-9
        assert bases[t]["beta0"] == pytest.approx(ora.effect_cv, rel=1e-8, abs=1e-10)


def test_eigh_by_index_range_assembles_the_full_decomposition(gpu_ctx):
    """gapit_eigh_dev_range: three uneven slices of the spectrum, each solved from its own copy of K, put side by side
    are one eigendecomposition -- values equal to the full solve, V orthonormal ACROSS the slices, K V = V diag(values)."""
    import torch
    rng = np.random.default_rng(5)
    n = 700
    A = rng.normal(size=(n, n + 40))
    K = A @ A.T / (n + 40)
    ref = gb().emma_eigen_L_wo_Z(K, ctx=gpu_ctx)
    dev = torch.device("cuda", gpu_ctx.device)
    V_t = torch.empty((n, n), dtype=torch.float64, device=dev)          # row r = eigenvector r
    vals = np.empty(n)
    for lo, hi in ((0, 1), (1, 450), (450, n)):
        K_t = torch.from_numpy(K).to(dev)                               # symmetric: row- and column-major agree
        vals[lo:hi] = gb().eigh_dev_range(K_t.data_ptr(), n, lo, hi - lo, V_t[lo:hi].data_ptr(), ctx=gpu_ctx)
    V = V_t.cpu().numpy().T
    assert np.all(np.diff(vals) <= 1e-12) and np.allclose(vals, ref["values"], rtol=0, atol=1e-12 * vals[0])
    assert np.abs(V.T @ V - np.eye(n)).max() < 1e-11
    assert np.abs(K @ V - V * vals).max() < 1e-11 * vals[0]
    with pytest.raises(gb().GapitCudaError):
        gb().eigh_dev_range(V_t.data_ptr(), n, n - 3, 4, V_t.data_ptr(), ctx=gpu_ctx)      # past the end of the spectrum


def test_resident_eigen_system_from_kinship(gpu_ctx):
    """gapit_eigsys_from_kinship: eigen(K) decomposed and sliced on the device without the eigenvectors visiting the
    host -- the same eigen-system as gapit_eigh + gapit_eigsys_upload, so projections, BLUP/PEV and scan agree."""
    G, Y, X0 = case(260, 700, npc=2)
    K = gb().GAPIT_kinship_VanRaden(gb().Genotypes(G, ctx=gpu_ctx, marker_major=True), ctx=gpu_ctx, verbose=False)
    eL = gb().emma_eigen_L_wo_Z(K, ctx=gpu_ctx)
    a = gb().EigenSystem(eL["values"], eL["vectors"], ctx=gpu_ctx)
    b = gb().EigenSystem.from_kinship(K, ctx=gpu_ctx)
    assert np.allclose(a.values, b.values, rtol=0, atol=1e-13 * a.values[0])
    R = np.column_stack([X0, Y])
    pa, pb = a.project(R), b.project(R)
    reml_a = gb().reml_from_projections(a.values, pa[:, -1], pa[:, :-1])
    reml_b = gb().reml_from_projections(b.values, pb[:, -1], pb[:, :-1])
    assert reml_b["delta"] == pytest.approx(reml_a["delta"], rel=1e-10)
    geno = gb().Genotypes(G, ctx=gpu_ctx, marker_major=True)
    ra, _ = gb().p3d_scan(geno, X0, Y, eigsys=a, delta=[reml_a["delta"]], vg=[reml_a["vg"]], ctx=gpu_ctx)
    rb, _ = gb().p3d_scan(geno, X0, Y, eigsys=b, delta=[reml_a["delta"]], vg=[reml_a["vg"]], ctx=gpu_ctx)
    for k in ("effect", "se", "pvalue", "rsquare"):
        assert np.allclose(ra[k], rb[k], rtol=1e-9, atol=1e-300, equal_nan=True), k
    ba, bb = gb().blup_pev(a, X0, Y, reml_a["delta"], reml_a["vg"], ctx=gpu_ctx), gb().blup_pev(b, X0, Y, reml_a["delta"], reml_a["vg"], ctx=gpu_ctx)
    for x, y in zip(ba, bb):
        assert np.allclose(x, y, rtol=1e-8, atol=1e-10)
    # and through the reference-facing call: no eig_L injected -> the resident route; injected -> the host route
    r1 = gb().GAPIT_EMMAxP3D(Y, G.T, K, X0=X0, ctx=gpu_ctx)
    r2 = gb().GAPIT_EMMAxP3D(Y, G.T, K, X0=X0, ctx=gpu_ctx, eig_L=eL)
    assert np.allclose(r1["ps"], r2["ps"], rtol=1e-8, atol=1e-300, equal_nan=True) and r1["vgs"] == pytest.approx(r2["vgs"], rel=1e-9)
