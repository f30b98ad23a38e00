"""Generate the committed golden vectors under tests/golden/ (TEST INFRASTRUCTURE ONLY).

The reference ships no fixtures and R cannot run in this image, so
these vectors are produced by the oracle restatement and cross-checked, at
generation time, against the independent derivations in oracle/crosscheck.py.
Run from the repo root:  python -m oracle.gen_golden
"""
from __future__ import annotations

import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from gapit_b200 import synth  # noqa: E402  (data generator only; no CUDA involved)
from oracle import crosscheck as cc  # noqa: E402
from oracle import gapit_oracle as go  # noqa: E402


def make_case(name, n, m, npc, seed, store_inputs):
    G = synth.make_genotypes(n, m, seed=seed, threads=1)
    # guarantee the edge cases of GAPIT.EMMAxP3D.R:499,512 and GAPIT.kinship.VanRaden.R:11-13 are present
    G[3, :] = 0
    G[7, :] = 2
    G[11, :] = 1
    G[20, :] = G[19, :]
    Y = synth.make_phenotypes(G, seed=seed)[:, 0]
    X0 = synth.make_covariates(G, npc=npc)
    snps = G.T
    K = go.kinship_vanraden(snps)
    gram, s, sum_c2, D, _ = cc.kinship_int_parts(snps)
    x = G.astype(np.int64)
    gram_all = x.T @ x                      # over ALL markers (what gapit_vanraden's G_out reports)
    Kex = cc.kinship_exact(snps) if n <= 80 else None
    if Kex is not None:
        assert np.abs(K - Kex).max() <= 1e-13 * np.abs(Kex).max()
    eig_L = go.emma_eigen_L_wo_Z(K)
    reml = go.emma_REMLE(Y, X0, K)
    mlm = go.emmax_p3d(Y, snps, K, X0, ncol_CVI=X0.shape[1], eig_L=eig_L, reml=reml)
    glm = go.emmax_p3d(Y, snps, None, X0, ncol_CVI=X0.shape[1])
    # pin the oracle against the dense-GLS derivation on a few markers
    for i in range(0, m, max(1, m // 12)):
        if np.isnan(mlm.stats[i]):
            continue
        d = cc.gls_dense(Y, X0, snps[:, i], K, reml["delta"], reml["vg"])
        assert abs(d["beta"][-1] - mlm.effect_est[i]) <= 1e-9 * max(1e-3, abs(d["beta"][-1]))
        assert abs(d["t"] - mlm.stats[i]) <= 1e-9 * max(1e-3, abs(d["t"]))
    out = {
        "n": n, "m": m, "seed": seed, "npc": npc,
        "K_diag": np.diag(K).copy(), "K_row0": K[0].copy(), "K_checksum": np.array([K.sum(), (K * K).sum()]),
        "gram_diag": np.diag(gram_all).copy(), "gram_row0": gram_all[0].copy(), "gram_checksum": np.array([gram_all.sum()]),
        "delta": reml["delta"], "vg": reml["vg"], "ve": reml["ve"], "REML": reml["REML"],
        "mlm_ps": mlm.ps, "mlm_stats": mlm.stats, "mlm_effect": mlm.effect_est, "mlm_se": mlm.se_clean,
        "mlm_rsquare": mlm.rsquare, "mlm_rsquare_base": mlm.rsquare_base, "mlm_logLM": mlm.logLM, "maf": mlm.maf,
        "mlm_stderr": mlm.stderr, "mlm_REMLs": mlm.REMLs,
        "glm_ps": glm.ps, "glm_stats": glm.stats, "glm_effect": glm.effect_est, "glm_ves": glm.ves,
        "glm_REMLs": glm.REMLs, "glm_rsquare": glm.rsquare,
    }
    if store_inputs:
        out.update({"G": G, "Y": Y, "X0": X0, "K": K, "gram": gram_all,
                    "eigL_values": eig_L["values"], "eigL_vectors": eig_L["vectors"]})
    path = os.path.join(ROOT, "tests", "golden", name + ".npz")
    np.savez_compressed(path, **out)
    print(name, "->", path, os.path.getsize(path) // 1024, "KiB; delta", reml["delta"])


def make_ingest_case(name, n, m, seed):
    """HapMap text -> dosages (GAPIT.HapMap / GAPIT.Numericalization), ZHANG kinship and prcomp on the result."""
    G = synth.make_genotypes(n, m, seed=seed, threads=1)
    out = {"n": n, "m": m, "seed": seed}
    for bit in (1, 2):
        text = synth.make_hapmap_text(G, bit=bit, missing=0.04, seed=seed, crlf=(bit == 1))
        rows = synth.hapmap_rows(text)
        out[f"text{bit}"] = np.frombuffer(text, dtype=np.uint8)
        for impute in ("None", "Middle", "Minor", "Major"):
            for maz in (False, True):
                gd = go.hapmap(rows, SNP_impute=impute, Major_allele_zero=maz)["GD"]
                # independent check of the reference's rules: no value outside {0,1,2,NA}; NA only without a rule
                assert np.all(np.isnan(gd) | np.isin(gd, (0.0, 1.0, 2.0)))
                assert impute == "None" or not np.isnan(gd).any()
                out[f"GD{bit}_{impute}_{int(maz)}"] = np.where(np.isnan(gd), -1, gd).astype(np.int8)
        out[f"GD{bit}_Dom"] = go.hapmap(rows, SNP_effect="Dom", SNP_impute="Middle")["GD"].astype(np.int8)
    X = out["GD2_Major_0"].astype(np.float64)
    Kz = go.kinship_zhang(X)
    # independent derivation of the ZHANG raw matrix: centred cross-product from exact integers
    keep = (X.sum(0) > 0) & (X.sum(0) < 2 * n)
    xi = X[:, keep].astype(np.int64)
    c = xi.sum(0)
    num = n * n * (xi @ xi.T) - n * ((xi @ c)[:, None] + (xi @ c)[None, :]) + int((c * c).sum())
    raw = num / float(n * n)
    top = 1.0 + (1.0 - ((xi == 1).sum(1) / (2.0 * xi.shape[1])).min())
    chk = top * (raw - raw.min()) / (np.diag(raw).max() - raw.min())
    diag = np.eye(n, dtype=bool)
    Dmin = chk[diag].min()
    if Dmin < 1.0:
        chk[diag] = (chk[diag] - Dmin + 1.0) / ((top + 1.0 - Dmin) * .5)
        chk[~diag] = chk[~diag] * (1.0 / Dmin)
    if chk[~diag].max() > top:
        chk[~diag] = chk[~diag] * (top / chk[~diag].max())
    assert np.abs(chk - Kz).max() <= 1e-11 * np.abs(Kz).max()
    x, ev = go.prcomp(X)
    out.update({"zhang": Kz, "pca_ev": ev, "pca_x_abs": np.abs(x[:, :4])})
    path = os.path.join(ROOT, "tests", "golden", name + ".npz")
    np.savez_compressed(path, **out)
    print(name, "->", path, os.path.getsize(path) // 1024, "KiB")


def main():
    make_ingest_case("ingest_n37_m90", 37, 90, 77)
    # tiny case with every input stored (self-contained known-answer vector)
    make_case("tiny_n48_m160_q3", 48, 160, 2, 101, store_inputs=True)
    # BASELINE configs[0] shape; inputs are re-generated from the seed (synth is deterministic)
    make_case("c1_n281_m3093_q1", 281, 3093, 0, 20160301, store_inputs=False)


if __name__ == "__main__":
    main()
