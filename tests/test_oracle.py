"""CPU tests: the oracle against independent derivations and the committed golden vectors."""
import math

import numpy as np
import pytest

from gapit_b200 import synth
from oracle import crosscheck as cc
from oracle import gapit_oracle as go
from oracle import rmath
from conftest import c1_inputs


def small_case(n=70, m=150, npc=1, seed=3):
    G = synth.make_genotypes(n, m, seed=seed, threads=1)
    G[5, :] = 2
    G[6, :] = 0
    G[9, :] = G[8, :]
    Y = synth.make_phenotypes(G, seed=seed)[:, 0]
    X0 = synth.make_covariates(G, npc=npc)
    return G, Y, X0


def test_kinship_hand_case():
    # n = 4 taxa, m = 3 markers, computed by hand from GAPIT.kinship.VanRaden.R:20-32
    snps = np.array([[0, 1, 2], [1, 1, 0], [2, 0, 0], [1, 2, 2]], dtype=float)
    p = snps.sum(0) / 8.0                       # [0.5, 0.5, 0.5]
    Z = snps - 2 * p                            # x - 2p
    K = Z @ Z.T / (2 * np.sum(p * (1 - p)))     # / 1.5
    assert np.allclose(go.kinship_vanraden(snps), K, atol=1e-15)
    assert np.allclose(cc.kinship_exact(snps), K, atol=1e-15)
    assert K[0, 0] == pytest.approx(2.0 / 1.5)


def test_kinship_drops_monomorphic_and_matches_exact():
    G, _, _ = small_case()
    K = go.kinship_vanraden(G.T)
    Kex = cc.kinship_exact(G.T)
    assert np.abs(K - Kex).max() <= 1e-13 * np.abs(Kex).max()
    keep = np.ones(G.shape[0], bool)
    keep[[5, 6]] = False                        # all-2 and all-0 columns contribute nothing (:11-13)
    assert np.abs(go.kinship_vanraden(G[keep].T) - K).max() <= 1e-13
    # K 1 = 0: centring by 2p makes every row sum vanish
    assert np.abs(K.sum(1)).max() < 1e-10


def test_gram_identities():
    G, _, _ = small_case()
    gram, s, sum_c2, D, n = cc.kinship_int_parts(G.T)
    assert np.array_equal(gram.sum(1), s)       # s_i = sum_j G_ij
    assert int(s.sum()) == sum_c2               # sum_i s_i = sum_k c_k^2


def test_zeroin_matches_scipy_brentq_root():
    from scipy import optimize
    f = lambda x: math.cos(x) - x
    r = rmath.uniroot(f, 0.0, 1.0)
    assert abs(r - optimize.brentq(f, 0, 1, xtol=1e-14)) < rmath.DBL_EPSILON ** 0.25
    assert rmath.uniroot(lambda x: x, 0.0, 1.0) == 0.0      # root at an end point returns it


def test_reml_is_a_local_maximum_and_consistent():
    G, Y, X0 = small_case()
    K = go.kinship_vanraden(G.T)
    eR = go.emma_eigen_R_wo_Z(K, X0)
    r = go.emma_REMLE(Y, X0, K, eig_R=eR)
    etas = eR["vectors"].T @ Y
    ll = lambda ld: go.emma_delta_REML_LL_wo_Z(ld, eR["values"], etas)
    ld = math.log(r["delta"])
    assert ll(ld) >= ll(ld - 0.05) and ll(ld) >= ll(ld + 0.05)
    assert r["ve"] == pytest.approx(r["vg"] * r["delta"])
    assert eR["values"].shape[0] == Y.shape[0] - X0.shape[1]
    # eigenvectors of S(K+I)S span the complement of X0
    assert np.abs(eR["vectors"].T @ X0).max() < 1e-8


def test_p3d_matches_dense_gls_and_emma_reml_t():
    G, Y, X0 = small_case()
    snps = G.T
    K = go.kinship_vanraden(snps)
    eL = go.emma_eigen_L_wo_Z(K)
    reml = go.emma_REMLE(Y, X0, K)
    res = go.emmax_p3d(Y, snps, K, X0, ncol_CVI=X0.shape[1], eig_L=eL, reml=reml)
    for i in range(0, snps.shape[1], 7):
        if np.isnan(res.stats[i]):
            continue
        d = cc.gls_dense(Y, X0, snps[:, i], K, reml["delta"], reml["vg"])
        assert res.effect_est[i] == pytest.approx(d["beta"][-1], rel=1e-9, abs=1e-12)
        assert res.stats[i] == pytest.approx(d["t"], rel=1e-9, abs=1e-12)
        assert res.se_clean[i] == pytest.approx(d["se"], rel=1e-9)
        t2, _ = cc.emma_reml_t_row(Y, X0, snps[:, i], eL["values"], eL["vectors"], reml["delta"], reml["vg"])
        assert res.stats[i] == pytest.approx(t2, rel=1e-9, abs=1e-12)
        # logLM through the dense route: RSS and log-determinant of K + delta I
        n = Y.shape[0]
        ll = 0.5 * (-n * math.log(2 * math.pi / n * d["rss"]) - d["logdet"] - n)
        assert res.logLM[i] == pytest.approx(ll, rel=1e-10)
    # stderr quirk (:972): with ncol(CVI) == q0 it is beta_marker / t == clean SE
    ok = ~np.isnan(res.stats)
    assert np.allclose(res.stderr[ok], res.se_clean[ok], rtol=1e-12)
    res2 = go.emmax_p3d(Y, snps, K, X0, ncol_CVI=None, eig_L=eL, reml=reml)
    assert np.all(np.isnan(res2.stderr))


def test_p3d_edge_rows_and_symmetries():
    G, Y, X0 = small_case()
    snps = G.T
    K = go.kinship_vanraden(snps)
    eL = go.emma_eigen_L_wo_Z(K)
    reml = go.emma_REMLE(Y, X0, K)
    res = go.emmax_p3d(Y, snps, K, X0, eig_L=eL, reml=reml)
    # constant markers: p = 1, the rest NA (:499-511)
    for i in (5, 6):
        assert res.ps[i] == 1.0 and np.isnan(res.stats[i]) and np.isnan(res.effect_est[i])
    # duplicated marker: row copied (:512-527)
    assert res.ps[9] == res.ps[8] and res.stats[9] == res.stats[8]
    # recoding 0 <-> 2 flips the sign of beta and t, nothing else
    flip = go.emmax_p3d(Y, 2 - snps, K, X0, eig_L=eL, reml=reml)
    ok = ~np.isnan(res.stats)
    assert np.allclose(flip.effect_est[ok], -res.effect_est[ok], rtol=1e-8, atol=1e-12)
    assert np.allclose(flip.ps[ok], res.ps[ok], rtol=1e-8)
    # eigenvector sign does not matter
    eL2 = {"values": eL["values"], "vectors": eL["vectors"] * np.where(np.arange(len(Y)) % 2, -1.0, 1.0)[None, :]}
    sg = go.emmax_p3d(Y, snps, K, X0, eig_L=eL2, reml=reml)
    assert np.allclose(sg.stats[ok], res.stats[ok], rtol=1e-9, atol=1e-12)


def test_glm_matches_ols_formulas():
    G, Y, X0 = small_case()
    snps = G.T
    res = go.emmax_p3d(Y, snps, None, X0)
    n, q0 = X0.shape
    b0, *_ = np.linalg.lstsq(X0, Y, rcond=None)
    ves = float(((Y - X0 @ b0) ** 2).sum() / (n - q0))
    assert res.ves == pytest.approx(ves, rel=1e-10)
    i = 12
    X = np.column_stack([X0, snps[:, i]])
    iXX = np.linalg.inv(X.T @ X)
    b = iXX @ X.T @ Y
    # the reference divides by the reduced-model variance, not the full-model one (:937)
    assert res.stats[i] == pytest.approx(b[-1] / math.sqrt(iXX[-1, -1] * ves), rel=1e-9)


def test_pvalue_tail_against_mpmath():
    for t, df in [(0.3, 50.0), (2.5, 279.0), (9.0, 279.0), (25.0, 4998.0), (38.0, 4998.0), (37.0, 19998.0)]:
        p = float(rmath.pt_upper_two_sided(t, df))
        ref = rmath.log10_pt_upper_two_sided_mp(t, df)
        assert math.log10(p) == pytest.approx(ref, abs=1e-9)
    # beyond the double range the tail underflows to 0, as R's pt does (-log10 p = Inf)
    assert float(rmath.pt_upper_two_sided(60.0, 19998.0)) == 0.0


def test_golden_tiny_reproduces(tiny_golden):
    g = tiny_golden
    G, Y, X0 = g["G"], g["Y"], g["X0"]
    K = go.kinship_vanraden(G.T)
    assert np.abs(K - g["K"]).max() <= 1e-13
    x = G.astype(np.int64)
    assert np.array_equal(x.T @ x, g["gram"])
    eL = {"values": g["eigL_values"], "vectors": g["eigL_vectors"]}
    reml = {"delta": float(g["delta"]), "vg": float(g["vg"]), "ve": float(g["ve"]), "REML": float(g["REML"])}
    res = go.emmax_p3d(Y, G.T, K, X0, ncol_CVI=X0.shape[1], eig_L=eL, reml=reml)
    assert np.allclose(res.ps, g["mlm_ps"], rtol=1e-10, equal_nan=True)
    assert np.allclose(res.stats, g["mlm_stats"], rtol=1e-10, atol=1e-13, equal_nan=True)
    r2 = go.emma_REMLE(Y, X0, K)
    assert r2["delta"] == pytest.approx(float(g["delta"]), rel=1e-6)


def test_golden_c1_reproduces(c1_golden):
    g = c1_golden
    G, Y, X0 = c1_inputs(g)
    assert G.shape == (3093, 281)
    x = G.astype(np.int64)
    gram = x.T @ x
    assert np.array_equal(np.diag(gram), g["gram_diag"]) and int(gram.sum()) == int(g["gram_checksum"][0])
    K = go.kinship_vanraden(G.T)
    assert np.allclose(np.diag(K), g["K_diag"], rtol=1e-12)
    glm = go.emmax_p3d(Y, G.T, None, X0, ncol_CVI=1)
    assert np.allclose(glm.ps, g["glm_ps"], rtol=1e-10, equal_nan=True)


# ------------------------------------------------------------------ HapMap numericalization (GAPIT.Numericalization.R)
def test_numericalization_hand_cases():
    f = go.numericalization
    nan = np.nan
    x = ["AA", "AG", "GG", "NN", "AA", "GG", "GG"]
    assert np.array_equal(f(x), [0, 1, 2, nan, 0, 2, 2], equal_nan=True)              # :85, impute None keeps NA
    assert np.array_equal(f(x, impute="Middle"), [0, 1, 2, 1, 0, 2, 2])                # :92
    assert np.array_equal(f(x, impute="Minor"), [0, 1, 2, 1, 0, 2, 2])                 # :95 rarest class is the het
    assert np.array_equal(f(x, impute="Major"), [0, 1, 2, 2, 0, 2, 2])                 # :96
    # Major.allele.zero: the more frequent homozygote (GG) becomes 0 (:53-57)
    assert np.array_equal(f(x, impute="Major", Major_allele_zero=True), [2, 1, 0, 0, 2, 0, 0])
    # one-character codes: het letter sorts last, K -> Z (:13), X - + / -> N (:9-12)
    assert np.array_equal(f(["A", "G", "R", "N", "A", "G", "G"], bit=1, impute="Major"), [0, 2, 1, 2, 0, 2, 2])
    assert np.array_equal(f(["G", "T", "K", "X", "T"], bit=1), [0, 2, 1, nan, 2], equal_nan=True)
    # two levels: 0 / 2 (:79); Major/Minor impute on the 0/2 scale (:99-100)
    assert np.array_equal(f(["AA", "TT", "--", "TT"], impute="Major"), [0, 2, 2, 2])
    assert np.array_equal(f(["AA", "TT", "//", "TT"], impute="Minor"), [0, 2, 0, 2])
    # one level or more than three: all zero, NA included (:76)
    assert np.array_equal(f(["AA", "AA", "NN"]), [0, 0, 0])
    assert np.array_equal(f(["AA", "AC", "CA", "CC"]), [0, 0, 0, 0])
    # effects (:105-107)
    assert np.array_equal(f(x, impute="Middle", effect="Dom"), [0, 1, 0, 1, 0, 0, 0])
    assert np.array_equal(f(x, impute="Middle", effect="Left"), [0, 0, 2, 0, 0, 2, 2])
    assert np.array_equal(f(x, impute="Middle", effect="Right"), [0, 2, 2, 2, 0, 2, 2])
    # ties in the counts keep the sorted level order (R's stable order())
    assert np.array_equal(f(["AA", "GG", "NN"], impute="Major", Major_allele_zero=True), [0, 2, 2])


def test_hapmap_table_restatement():
    G = synth.make_genotypes(9, 40, seed=4, threads=1)
    text = synth.make_hapmap_text(G, bit=2, missing=0.0, odd_rows=False)
    hm = go.hapmap(synth.hapmap_rows(text), SNP_impute="Middle")
    assert hm["GT"][0] == "taxa00000" and len(hm["GT"]) == 9 and hm["GI"][3] == ("rs3", "4", "1111")
    # without missing calls a three-level row is the dosage itself or its mirror image (levels sort by allele letter)
    for k in range(40):
        col = hm["GD"][:, k]
        if len(set(G[k])) == 3:
            assert np.array_equal(col, G[k]) or np.array_equal(col, 2 - G[k])


def test_zhang_and_prcomp_restatements():
    G = synth.make_genotypes(40, 300, seed=8, threads=1)
    X = G.T.astype(np.float64)
    K = go.kinship_zhang(X)
    top = 1.0 + (1.0 - ((X == 1).sum(1) / (2.0 * ((X.sum(0) > 0) & (X.sum(0) < 80)).sum())).min())
    off = K[~np.eye(40, dtype=bool)]
    assert np.allclose(K, K.T) and off.max() <= top * (1 + 1e-12) and np.diag(K).min() >= 1.0 - 1e-12   # :47-58
    x, ev = go.prcomp(X)
    Xc = X - X.mean(0)
    assert np.allclose(x @ x.T, Xc @ Xc.T, atol=1e-8) and np.isclose(ev.sum(), Xc.var(0, ddof=1).sum())


def test_golden_ingest_reproduces(ingest_golden):
    """tests/golden/ingest_n37_m90.npz (oracle/gen_golden.py): HapMap text -> dosages for every impute rule and both
    genotype widths, ZHANG kinship and prcomp variances of the result."""
    g = ingest_golden
    for bit in (1, 2):
        rows = synth.hapmap_rows(g[f"text{bit}"].tobytes())
        for impute in ("None", "Middle", "Minor", "Major"):
            for maz in (False, True):
                gd = go.hapmap(rows, SNP_impute=impute, Major_allele_zero=maz)["GD"]
                assert np.array_equal(np.where(np.isnan(gd), -1, gd).astype(np.int8), g[f"GD{bit}_{impute}_{int(maz)}"])
    X = g["GD2_Major_0"].astype(np.float64)
    assert np.array_equal(go.kinship_zhang(X), g["zhang"])
    assert np.allclose(go.prcomp(X)[1], g["pca_ev"], rtol=1e-12, atol=1e-12)


def test_batched_derivation_agrees_with_the_per_marker_loop():
    """oracle/batched.py (one dgemm V'X + closed-form bordered solve) against the restated per-marker loop, MLM and GLM:
    two independent routes to the same table -- the batched one is what checks thousands of markers at n = 5,000 and
    20,000 in the GPU tests and in bench.py's in-run parity block."""
    from oracle import batched
    from gapit_b200 import synth
    n, m = 240, 400
    G = synth.make_genotypes(n, m, seed=3)
    G[7] = 2                                    # a constant marker
    Y = synth.make_phenotypes(G)[:, 0]
    X0 = synth.make_covariates(G, npc=2)
    K = go.kinship_vanraden(G.T)
    eL = go.emma_eigen_L_wo_Z(K)
    reml = go.emma_REMLE(Y, X0, K)
    ora = go.emmax_p3d(Y, G.T, K, X0, eig_L=eL, reml=reml)
    blk = batched.p3d_block(Y, G.T, X0, eL["values"], eL["vectors"], reml["delta"], reml["vg"])
    c = batched.compare({"effect": ora.effect_est, "se": ora.se_clean, "tvalue": ora.stats, "pvalue": ora.ps}, blk)
    assert c["nan_rows_identical"] and c["n_tested"] == m - 1
    assert c["beta_rel_max"] < 1e-10 and c["se_rel_max"] < 1e-13 and c["dlog10p_max"] < 1e-10
    ok = ~np.isnan(ora.stats)
    assert np.max(np.abs(ora.rsquare[ok] - blk["rsquare"][ok])) < 1e-13
    assert np.max(np.abs(ora.logLM[ok] - blk["loglm"][ok])) < 1e-9
    og = go.emmax_p3d(Y, G.T, None, X0)
    bg = batched.p3d_block(Y, G.T, X0, glm=True)
    c = batched.compare({"effect": og.effect_est, "se": og.se_clean, "tvalue": og.stats, "pvalue": og.ps}, bg)
    assert c["nan_rows_identical"] and c["beta_rel_max"] < 1e-10 and c["se_rel_max"] < 1e-13 and c["dlog10p_max"] < 1e-10
