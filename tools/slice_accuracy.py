"""Accuracy and speed of the rotation against the number of int8 digit planes S (the Ozaki split of V).

For each taxa count the scan is run with S = 8 (62 bits of every eigenvector element, beyond fp64's 53) as the
reference and with S = 4..7; reported per S: max |d beta| / max(|beta|, 1e-3) (the parity tests' measure),
max |d beta| / SE (error in t units), max |d SE| / SE, max |d log10 p|, and kernel time.
"""
import os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import gapit_b200 as gb
from gapit_b200 import synth

shapes = [tuple(int(v) for v in s.split("x")) for s in os.environ.get("ACC_SHAPES", "260x3000,1000x20000,5000x100000").split(",")]
ctx = gb.Context(0)
for n, m in shapes:
    G = synth.make_genotypes(n, m, seed=11, threads=16)
    Y = synth.make_phenotypes(G[:4096])
    X0 = np.ones((n, 1))
    geno = gb.Genotypes(G, ctx=ctx, marker_major=True)
    K = gb.GAPIT_kinship_VanRaden(geno, ctx=ctx, verbose=False)
    eL = gb.emma_eigen_L_wo_Z(K, ctx=ctx)
    reml = gb.GAPIT_emma_REMLE(Y[:, 0], X0, K, eig_L=eL, ctx=ctx)
    res = {}
    for S in (8, 7, 6, 5, 4):
        ctx.set_slices(S)
        es = gb.EigenSystem(eL["values"], eL["vectors"], ctx=ctx)
        km = 1e9
        for _ in range(3):
            arrs, _b = gb.p3d_scan(geno, X0, Y[:, :1], eigsys=es, delta=[reml["delta"]], vg=[reml["vg"]], ctx=ctx)
            km = min(km, ctx.stats()["kernel_ms"])
        res[S] = ({k: v.copy() for k, v in arrs.items()}, km)
        es.close()
    ref = res[8][0]
    ok = ~np.isnan(ref["tvalue"][0])
    print(f"==== n={n} m={m} delta={reml['delta']:.6g}  median SE={np.median(ref['se'][0][ok]):.4g} "
          f"max -log10p={-np.log10(max(np.nanmin(ref['pvalue'][0]), 1e-300)):.1f}", flush=True)
    for S in (7, 6, 5, 4):
        a, km = res[S]
        db = np.abs(a["effect"][0][ok] - ref["effect"][0][ok])
        rel = np.max(db / np.maximum(np.abs(ref["effect"][0][ok]), 1e-3))
        tun = np.max(db / ref["se"][0][ok])
        dse = np.max(np.abs(a["se"][0][ok] - ref["se"][0][ok]) / ref["se"][0][ok])
        with np.errstate(divide="ignore", invalid="ignore"):
            dlp = np.nanmax(np.abs(np.log10(a["pvalue"][0][ok]) - np.log10(ref["pvalue"][0][ok])))
        print(f"S={S}: beta rel(floor 1e-3) {rel:.3e}  |dbeta|/SE {tun:.3e}  SE rel {dse:.3e}  dlog10p {dlp:.3e}  "
              f"kernel {km:.3f} ms ({m / km / 1e3:.2f} M SNPs/s)", flush=True)
    geno.close()
