"""TEST INFRASTRUCTURE (like everything under oracle/): a BATCHED fp64 derivation of the P3D scan, independent of both
the reference's per-marker loop (oracle/gapit_oracle.py:emmax_p3d) and of the product's int8 digit split.

For a block of markers X (n x mb, dosages) it forms the rotated markers with one dgemm, W = V'X, and then the closed-form
bordered solve of GAPIT.EMMAxP3D.R:736-748,772:  with w = 1/(lambda + delta), A = V'X0, b = V'y,

    xx = sum_k w W^2,  xa = A' (w W),  xy = (w b)' W,
    B22 = xx - xa' (A' w A)^-1 xa,   beta = (xy - xa' beta0) / B22,   se = sqrt(vg / B22),   t = beta / se   (:936),
    p = 2 pt(|t|, n - q0 - 1, lower = FALSE)  (:939),   RSS = RSS0 - (xy - xa' beta0) beta,   logLM (:958),  R^2 (:967).

It checks thousands of markers at the BASELINE shapes (n = 5,000 and 20,000) in seconds, where the per-marker loop
manages a few hundred.  GLM (K = NULL, :759-763,:865-870): V = I, w = 1, variance ves = RSS0 / (n - q0).
"""
from __future__ import annotations

import numpy as np
from scipy import special


def p3d_block(y, X, X0, values=None, vectors=None, delta=None, vg=None, glm=False):
    """y (n), X (n, mb) any numeric dtype, X0 (n, q0), eigen-system of K (values descending, vectors n x n).
    Returns dict of (mb,) arrays: effect, se, tvalue, pvalue, log10p, rsquare, loglm; constant markers -> NaN / p = 1."""
    y = np.asarray(y, dtype=np.float64)
    X0 = np.asarray(X0, dtype=np.float64)
    Xd = np.asarray(X, dtype=np.float64)
    n, q0 = X0.shape
    if glm:
        W, A, b, w, sumlog = Xd, X0, y, np.ones(n), 0.0
    else:
        V = np.asarray(vectors, dtype=np.float64)
        lam = np.asarray(values, dtype=np.float64)
        W = V.T @ Xd                                   # the rotation of :634 for the whole block, one dgemm
        A = V.T @ X0
        b = V.T @ y
        w = 1.0 / (lam + delta)
        sumlog = float(np.sum(np.log(lam + delta)))
    wW = w[:, None] * W
    xx = np.einsum("km,km->m", wW, W)
    xa = A.T @ wW                                      # (q0, mb)
    xy = (w * b) @ W
    XX = A.T @ (w[:, None] * A)
    iXX = np.linalg.inv(XX)
    beta0 = iXX @ (A.T @ (w * b))
    r0 = b - A @ beta0
    rss0 = float(np.sum(w * r0 * r0))
    quad = np.einsum("im,ij,jm->m", xa, iXX, xa)
    num = xy - beta0 @ xa
    with np.errstate(divide="ignore", invalid="ignore"):
        B22 = xx - quad
        beta = num / B22
        var = (rss0 / (n - q0)) if glm else vg
        se = np.sqrt(var / B22)
        t = beta / se
        df = float(n - q0 - 1)
        x = df / (df + t * t)
        p = special.betainc(0.5 * df, 0.5, x)          # = 2 pt(|t|, df, lower = FALSE)
        # log10 p without underflow: log I_x(a, 1/2) through the complementary form is not needed at these sizes;
        # where p underflows to 0 the comparison uses p itself
        log10p = np.log10(p)
        rss = rss0 - num * beta
        ll = 0.5 * (-n * np.log((2.0 * np.pi / n) * rss) - sumlog - n)
    ybar = y.mean()
    ss0 = float(np.sum((y - ybar) ** 2))
    logL0 = 0.5 * (-n * np.log((2.0 * np.pi / n) * ss0) - n)
    with np.errstate(over="ignore", invalid="ignore"):
        rsq = 1.0 - np.exp(-(2.0 / n) * (ll - logL0))
    const = Xd.min(axis=0) == Xd.max(axis=0)           # :499-511
    out = {"effect": beta, "se": se, "tvalue": t, "pvalue": p, "log10p": log10p, "rsquare": rsq, "loglm": ll}
    for k in out:
        out[k] = np.where(const, np.nan, out[k])
    out["pvalue"] = np.where(const, 1.0, out["pvalue"])
    out["log10p"] = np.where(const, 0.0, out["log10p"])
    return out


def compare(gpu, ref, rows=None, floor=1e-3):
    """Agreement of a dict of product arrays (trait 0) with p3d_block's output on the same markers.
    Relative errors of beta / se over markers whose |beta| exceeds `floor` x max|beta| (beta of an unassociated marker
    is a difference of nearly equal sums; its error is judged in SE units instead)."""
    sl = slice(None) if rows is None else rows
    g = {k: np.asarray(gpu[k])[sl] for k in ("effect", "se", "tvalue", "pvalue")}
    ok = ~np.isnan(ref["tvalue"])
    same_nan = bool(np.array_equal(np.isnan(g["tvalue"]), ~ok))
    big = ok & (np.abs(ref["effect"]) > floor * np.nanmax(np.abs(ref["effect"])))
    beta_rel = float(np.max(np.abs(g["effect"][big] - ref["effect"][big]) / np.abs(ref["effect"][big]))) if big.any() else 0.0
    se_rel = float(np.max(np.abs(g["se"][ok] - ref["se"][ok]) / ref["se"][ok])) if ok.any() else 0.0
    beta_se_units = float(np.max(np.abs(g["effect"][ok] - ref["effect"][ok]) / ref["se"][ok])) if ok.any() else 0.0
    with np.errstate(divide="ignore", invalid="ignore"):
        dl = np.abs(np.log10(g["pvalue"][ok]) - ref["log10p"][ok])
    dl = dl[np.isfinite(dl)]
    return {"n_markers": int(ok.size), "n_tested": int(ok.sum()), "nan_rows_identical": same_nan,
            "beta_rel_max": beta_rel, "se_rel_max": se_rel, "beta_err_in_se_units_max": beta_se_units,
            "dlog10p_max": float(dl.max()) if dl.size else 0.0}
