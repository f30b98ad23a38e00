"""First-light diagnostics on a B200: each stage against the oracle, verbose (run by hand: python tests/diag_first_light.py).
Lives under tests/ because it uses the oracle, which only tests, smoke() and bench.py's CPU legs may touch."""
import os, sys, time, traceback
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import gapit_b200 as gb
from gapit_b200 import synth
from oracle import gapit_oracle as go, crosscheck as cc

def stage(name):
    print(f"\n==== {name}", flush=True)

def main():
    n, m = int(os.environ.get("DIAG_N", 300)), int(os.environ.get("DIAG_M", 1000))
    G = synth.make_genotypes(n, m, seed=11)
    snps = G.T
    Y = synth.make_phenotypes(G)
    X0 = synth.make_covariates(G, npc=2)
    ctx = gb.Context(0)
    stage("pack + colsums")
    geno = gb.Genotypes(G, ctx=ctx, marker_major=True)
    cs, cq = geno.colsums()
    x = G.astype(np.int64)
    print("colsum ok", np.array_equal(cs, x.sum(1)), "colsumsq ok", np.array_equal(cq, (x * x).sum(1)))
    geno_f = gb.Genotypes(snps.astype(np.float64), ctx=ctx)
    cs2, _ = geno_f.colsums()
    print("f64 pack colsum ok", np.array_equal(cs2, cs))

    stage("vanraden gram")
    try:
        K, Gm = gb.GAPIT_kinship_VanRaden(geno, ctx=ctx, return_gram=True, verbose=False)
        Gref = (x.T @ x)  # n x n
        bad = np.argwhere(Gm != Gref)
        print("G bit-exact:", bad.shape[0] == 0, "mismatches", bad.shape[0], "of", Gref.size)
        if bad.shape[0]:
            print("first mismatches", bad[:10].tolist())
            for (i, j) in bad[:10]:
                print(i, j, Gm[i, j], Gref[i, j])
            print("nonzero frac got", (Gm != 0).mean(), "ref", (Gref != 0).mean())
            print("G[0,:8]", Gm[0, :8], Gref[0, :8])
            print("G[128,:8]", Gm[128, :8] if n > 128 else None, Gref[128, :8] if n > 128 else None)
            print("G[:8,0]", Gm[:8, 0], Gref[:8, 0])
        Kref = go.kinship_vanraden(snps)
        Kex = cc.kinship_exact(snps)
        print("K vs oracle maxabs", np.abs(K - Kref).max(), "K vs exact maxabs", np.abs(K - Kex).max(), "max|K|", np.abs(Kex).max())
    except Exception:
        traceback.print_exc()
        K = go.kinship_vanraden(snps)

    Kref = go.kinship_vanraden(snps)
    stage("eigh / eigen_R / reml")
    try:
        eL = gb.emma_eigen_L_wo_Z(Kref, ctx=ctx)
        eLo = go.emma_eigen_L_wo_Z(Kref)
        print("eig.L values maxabs", np.abs(eL["values"] - eLo["values"]).max())
        rec = (eL["vectors"] * eL["values"]) @ eL["vectors"].T
        print("eig.L reconstruction", np.abs(rec - Kref).max())
        eR = gb.emma_eigen_R_wo_Z(Kref, X0, ctx=ctx)
        eRo = go.emma_eigen_R_wo_Z(Kref, X0)
        print("eig.R values maxabs", np.abs(eR["values"] - eRo["values"]).max())
        r1 = gb.GAPIT_emma_REMLE(Y[:, 0], X0, Kref, eig_R=eR, ctx=ctx)
        r2 = go.emma_REMLE(Y[:, 0], X0, Kref)
        print("reml gpu", r1)
        print("reml ora", r2)
        r3 = gb.GAPIT_emma_REMLE(Y[:, 0], X0, Kref, eig_R=eRo, ctx=ctx)
        print("reml C++ on oracle eig.R", r3)
    except Exception:
        traceback.print_exc()

    eLo = go.emma_eigen_L_wo_Z(Kref)
    r2 = go.emma_REMLE(Y[:, 0], X0, Kref)
    ora = go.emmax_p3d(Y[:, 0], snps, Kref, X0, ncol_CVI=X0.shape[1], eig_L=eLo, reml=r2)
    for S in (6, 5, 7, 8, 4):
        stage(f"MLM scan, slices={S}")
        try:
            ctx.set_slices(S)
            res = gb.GAPIT_EMMAxP3D(Y[:, 0], geno, Kref, X0=X0, CVI=np.zeros((n, X0.shape[1])), ctx=ctx, eig_L=eLo, reml=r2)
            ok = ~np.isnan(ora.stats)
            def rel(a, b):
                return np.nanmax(np.abs(a - b) / np.maximum(np.abs(b), 1e-300))
            print("nan pattern equal", np.array_equal(np.isnan(res["stats"][:, 0]), np.isnan(ora.stats)))
            print("beta rel", rel(res["effect.est"][ok, 0], ora.effect_est[ok]))
            print("se   rel", rel(res["se"][ok, 0], ora.se_clean[ok]))
            print("t    rel", rel(res["stats"][ok, 0], ora.stats[ok]))
            lp = -np.log10(res["ps"][:, 0]); lo = -np.log10(ora.ps)
            print("-log10p abs", np.nanmax(np.abs(lp - lo)), "max -log10p", np.nanmax(lo))
            print("rsq abs", np.nanmax(np.abs(res["rsquare"][ok, 0] - ora.rsquare[ok])))
            print("logLM abs", np.nanmax(np.abs(res["logLM"][ok, 0] - ora.logLM[ok])))
            print("maf abs", np.nanmax(np.abs(res["maf"][:, 0] - ora.maf)))
            print("stderr abs", np.nanmax(np.abs(res["stderr"][ok, 0] - ora.stderr[ok])))
            print("sample beta", res["effect.est"][:4, 0], ora.effect_est[:4])
        except Exception:
            traceback.print_exc()
    ctx.set_slices(6)
    stage("GLM scan")
    try:
        og = go.emmax_p3d(Y[:, 0], snps, None, X0, ncol_CVI=X0.shape[1])
        rg = gb.GAPIT_EMMAxP3D(Y[:, 0], geno, None, X0=X0, CVI=np.zeros((n, X0.shape[1])), ctx=ctx)
        ok = ~np.isnan(og.stats)
        print("beta rel", np.nanmax(np.abs(rg["effect.est"][ok, 0] - og.effect_est[ok]) / np.abs(og.effect_est[ok])))
        print("t rel", np.nanmax(np.abs(rg["stats"][ok, 0] - og.stats[ok]) / np.abs(og.stats[ok])))
        print("-log10p abs", np.nanmax(np.abs(np.log10(rg["ps"][:, 0]) - np.log10(og.ps))))
        print("ves", rg["ves"], og.ves, "REMLs", rg["REMLs"], og.REMLs)
        print("rsq abs", np.nanmax(np.abs(rg["rsquare"][ok, 0] - og.rsquare[ok])))
    except Exception:
        traceback.print_exc()
    print("\nDIAG DONE", flush=True)

if __name__ == "__main__":
    main()
