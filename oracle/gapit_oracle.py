"""CPU restatement (numpy, fp64) of GAPIT's data-parallel hot path.

TEST INFRASTRUCTURE ONLY -- the parity oracle.  Only ``tests/``,
``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline`` /
``--impl reference`` legs may import this module; the product path
(``gapit_b200`` -> ``libgapitcuda.so``) never does.

PARITY UNPINNED: the reference (/root/reference, GAPIT "2016.03.01") ships no
tests, golden vectors or example data and R is not
installed in this image, so the restatement below cannot be checked against
the reference's own output.  It is pinned instead by independent derivations
(``oracle/crosscheck.py``: exact-rational kinship, dense Cholesky GLS,
mpmath tails) and hand-computable cases in ``tests/``.

Every function cites the reference lines it follows.  The operation order of
the reference is kept where it matters numerically; R base calls are replaced
by their numpy/scipy stand-ins (``eigen`` -> ``numpy.linalg.eigh`` re-sorted
descending, ``solve`` -> ``numpy.linalg.solve``/``inv``, ``MASS::ginv`` ->
``numpy.linalg.pinv``, ``uniroot``/``pt`` -> ``oracle.rmath``).
"""
from __future__ import annotations

import math
from dataclasses import dataclass, field

import numpy as np

from . import rmath


# --------------------------------------------------------------------------
# GAPIT.kinship.VanRaden  (GAPIT.kinship.VanRaden.R:1-37)
# --------------------------------------------------------------------------
def kinship_vanraden(snps: np.ndarray) -> np.ndarray:
    """n x m dosage matrix (0/1/2) -> n x n VanRaden kinship.

    Follows GAPIT.kinship.VanRaden.R line by line:
      :11-13  drop columns whose allele frequency is >=1 or <=0
      :20-22  p = colSums/(2n); P = 2(p-.5); snps-1
      :25     Z = t(snps) - P           (m x n, Z[k,i] = x_ik - 2 p_k)
      :28     K = crossprod(Z, Z)       (dgemm)
      :31-32  K / (2 sum p(1-p))
    ``hasInbred`` is accepted and ignored by the reference (:2).
    """
    snps = np.asarray(snps, dtype=np.float64)
    fa = snps.sum(axis=0) / (2.0 * snps.shape[0])
    index_non = (fa >= 1.0) | (fa <= 0.0)
    snps = snps[:, ~index_non]
    n_ind = snps.shape[0]
    p = snps.sum(axis=0) / (2.0 * n_ind)
    P = 2.0 * (p - 0.5)
    snps = snps - 1.0
    Z = snps.T - P[:, None]
    K = Z.T @ Z
    adj = 2.0 * np.sum(p * (1.0 - p))
    return K / adj


# --------------------------------------------------------------------------
# EMMA pieces on the path  (emma.txt)
# --------------------------------------------------------------------------
def _eigh_desc(A):
    """R's eigen(A, symmetric=TRUE): eigenvalues in decreasing order."""
    w, v = np.linalg.eigh(A)
    return w[::-1].copy(), v[:, ::-1].copy()


def emma_eigen_L_wo_Z(K):
    """emma.txt:58-61."""
    values, vectors = _eigh_desc(np.asarray(K, dtype=np.float64))
    return {"values": values, "vectors": vectors}


def emma_eigen_R_wo_Z(K, X):
    """emma.txt:82-89: eigen(S (K+I) S), keep the first n-q, values minus 1."""
    K = np.asarray(K, dtype=np.float64)
    X = np.asarray(X, dtype=np.float64)
    n, q = X.shape
    S = np.eye(n) - X @ np.linalg.solve(X.T @ X, X.T)
    values, vectors = _eigh_desc(S @ (K + np.eye(n)) @ S)
    return {"values": values[: n - q] - 1.0, "vectors": vectors[:, : n - q]}


def emma_delta_REML_LL_wo_Z(logdelta, lam, etas):
    """emma.txt:145-149."""
    nq = etas.shape[0]
    delta = math.exp(logdelta)
    return 0.5 * (nq * (math.log(nq / (2 * math.pi)) - 1 - math.log(np.sum(etas * etas / (lam + delta))))
                  - np.sum(np.log(lam + delta)))


def emma_delta_REML_dLL_wo_Z(logdelta, lam, etas):
    """emma.txt:158-164 (note: no leading delta factor, unlike the grid dLL)."""
    nq = etas.shape[0]
    delta = math.exp(logdelta)
    etasq = etas * etas
    ldelta = lam + delta
    return 0.5 * (nq * np.sum(etasq / (ldelta * ldelta)) / np.sum(etasq / ldelta) - np.sum(1.0 / ldelta))


def replace_nan(LL):
    """GAPIT.replaceNaN.R:9-12: NaN log-likelihoods become the worst finite one."""
    LL = np.array(LL, dtype=np.float64)
    index = np.isnan(LL)
    if LL.size > 0 and (~index).any():
        LL[index] = LL[~index].min()
    return LL


def emma_REMLE(y, X, K, ngrids=100, llim=-10.0, ulim=10.0, esp=1e-10, eig_R=None):
    """GAPIT.emma.REMLE.R:1-111, Z = NULL branch (:21-54, :96-110)."""
    y = np.asarray(y, dtype=np.float64).reshape(-1)
    X = np.asarray(X, dtype=np.float64)
    n = y.shape[0]
    q = X.shape[1]
    if np.linalg.det(X.T @ X) == 0:                       # :15-18
        return {"REML": 0.0, "delta": 0.0, "ve": 0.0, "vg": 0.0}
    if eig_R is None:                                     # :22-24
        eig_R = emma_eigen_R_wo_Z(K, X)
    lam = eig_R["values"]
    etas = eig_R["vectors"].T @ y                         # :25
    logdelta = np.arange(ngrids + 1) / ngrids * (ulim - llim) + llim   # :27
    m = logdelta.shape[0]
    delta = np.exp(logdelta)
    Lambdas = lam[:, None] + delta[None, :]               # :30
    Etasq = (etas * etas)[:, None]                        # :31
    with np.errstate(divide="ignore", invalid="ignore"):
        dLL = 0.5 * delta * ((n - q) * np.sum(Etasq / (Lambdas * Lambdas), axis=0)
                             / np.sum(Etasq / Lambdas, axis=0) - np.sum(1.0 / Lambdas, axis=0))  # :33
    optlogdelta = []
    optLL = []
    if dLL[0] < esp:                                      # :37
        optlogdelta.append(llim)
        optLL.append(emma_delta_REML_LL_wo_Z(llim, lam, etas))
    if dLL[m - 2] > 0 - esp:                              # :41  (sic: dLL[m-1], 1-based)
        optlogdelta.append(ulim)
        optLL.append(emma_delta_REML_LL_wo_Z(ulim, lam, etas))
    for i in range(m - 1):                                # :46-54
        if dLL[i] * dLL[i + 1] < 0 and dLL[i] > 0 and dLL[i + 1] < 0:
            root = rmath.uniroot(lambda ld: emma_delta_REML_dLL_wo_Z(ld, lam, etas),
                                 float(logdelta[i]), float(logdelta[i + 1]))
            optlogdelta.append(root)
            optLL.append(emma_delta_REML_LL_wo_Z(root, lam, etas))
    optLL_arr = np.array(optLL, dtype=np.float64)
    # :96 which.max ignores NaN (R: which.max drops NA/NaN)
    maxdelta = math.exp(optlogdelta[int(np.nanargmax(optLL_arr))])
    optLL_arr = replace_nan(optLL_arr)                    # :99
    maxLL = float(optLL_arr.max())                        # :101
    maxva = float(np.sum(etas * etas / (lam + maxdelta)) / (n - q))    # :103
    maxve = maxva * maxdelta                              # :108
    return {"REML": maxLL, "delta": maxdelta, "ve": maxve, "vg": maxva}


# --------------------------------------------------------------------------
# GAPIT.EMMAxP3D  (GAPIT.EMMAxP3D.R), Z = NULL / SNP.P3D / fullGD / no indicator
# --------------------------------------------------------------------------
def _solve_or_ginv(A):
    """try(solve(A)) with the MASS::ginv fallback (GAPIT.EMMAxP3D.R:708-712)."""
    try:
        inv = np.linalg.inv(A)
        if not np.all(np.isfinite(inv)) or np.linalg.cond(A) > 1.0 / np.finfo(float).eps:
            raise np.linalg.LinAlgError
        return inv, False
    except np.linalg.LinAlgError:
        return np.linalg.pinv(A), True


@dataclass
class P3DResult:
    ps: np.ndarray
    stats: np.ndarray
    effect_est: np.ndarray
    rsquare_base: np.ndarray
    rsquare: np.ndarray
    dfs: np.ndarray
    df: np.ndarray
    tvalue: np.ndarray
    stderr: np.ndarray
    maf: np.ndarray
    nobs: np.ndarray
    REMLs: float
    vgs: float
    ves: float
    delta: float
    logLM: np.ndarray
    logLM_Base: float
    logL0: float
    effect_cv: np.ndarray
    se_clean: np.ndarray = field(default=None)
    singular_X0: bool = False

    def as_dict(self):
        return dict(self.__dict__)


def emmax_p3d(ys, xs, K=None, X0=None, ncol_CVI=None,
              ngrids=100, llim=-10.0, ulim=10.0, esp=1e-10,
              eig_L=None, reml=None, progress=None) -> P3DResult:
    """The per-marker scan of GAPIT.EMMAxP3D.R:397-979 with its prologue.

    ys : n phenotypes (one trait; the reference reshapes it to 1 x n at :29)
    xs : n x m dosages, no NA (the reference imputes before the loop, :20-24)
    K  : n x n kinship, or None / 1x1 for the GLM branch (:34)
    X0 : n x q0 covariates including the intercept (default matrix(1,n,1), :30)
    ncol_CVI : ncol(CVI) as seen by :972 (``stderr = beta[ncol(CVI)+1]/t``);
               None reproduces the no-covariate call where that index runs past
               ``beta`` and R yields NA.
    eig_L / reml : optionally inject (values, vectors) and {delta, vg, ve, REML}
               so the tight parity tests feed both sides identical spectra.

    The loop keeps the reference's per-marker structure (two n x n products
    per marker, :634 and :957) -- it doubles as the CPU baseline.
    """
    ys = np.asarray(ys, dtype=np.float64).reshape(-1)
    xs = np.asarray(xs)
    n = ys.shape[0]
    if X0 is None:                                        # :30
        X0 = np.ones((n, 1))
    X0 = np.asarray(X0, dtype=np.float64)
    if K is not None and np.size(K) <= 1:                 # :34
        K = None
    q0 = X0.shape[1]                                      # :40-41
    q1 = q0 + 1
    m = xs.shape[1]

    X = X0
    if K is not None:
        K = np.asarray(K, dtype=np.float64)
        if eig_L is None:
            eig_L = emma_eigen_L_wo_Z(K)                  # :48
        if reml is None:
            # :58 computes eig.R, :202 passes it into the eig.L formal so REMLE
            # recomputes it (GAPIT.emma.REMLE.R:2-3,22-24) -- same numbers.
            reml = emma_REMLE(ys, X, K, ngrids=ngrids, llim=llim, ulim=ulim, esp=esp)
        vgs, ves, REMLs, delta = reml["vg"], reml["ve"], reml["REML"], reml["delta"]
        lam = np.asarray(eig_L["values"], dtype=np.float64)
        U = np.asarray(eig_L["vectors"], dtype=np.float64) * np.sqrt(1.0 / (lam + delta))[None, :]   # :224
        eig_full_plus_delta = lam + delta                 # :227
    else:
        vgs = ves = REMLs = delta = None
        U = None

    ps = np.full(m, np.nan)
    stats = np.full(m, np.nan)
    effect_est = np.full(m, np.nan)
    rsquare_base = np.full(m, np.nan)
    rsquare = np.full(m, np.nan)
    dfs = np.full(m, np.nan)
    df = np.full(m, np.nan)
    tvalue = np.full(m, np.nan)
    stderr = np.full(m, np.nan)
    se_clean = np.full(m, np.nan)
    maf = np.full(m, np.nan)
    nobs = np.full(m, np.nan)
    logLM_all = np.full(m, np.nan)

    # ---- i = 0: the model without a marker (:404-408, :529-561, :617-620,
    #      :646, :678-679, :707-714, :772, :865-870, :912-928)
    yv = ys
    X_int = np.ones((n, 1))                               # :536
    beta_int = np.linalg.solve(X_int.T @ X_int, X_int.T @ yv)   # :537-539
    res_int = yv - X_int @ beta_int
    logL0 = 0.5 * ((-n) * math.log(((2 * math.pi) / n) * float(res_int @ res_int)) - n)   # :551-553

    singular = False
    if K is not None:
        yt = U.T @ yv                                     # :617
        X0t = U.T @ X0                                    # :619
        X0X0 = X0t.T @ X0t                                # :646
        X0Y = X0t.T @ yt                                  # :678
        iX0X0, singular = _solve_or_ginv(X0X0)            # :708-712
        iXX = iX0X0
        XY = X0Y
    else:
        yt = yv                                           # :640
        X0t = X0                                          # :641
        X0X0 = X0t.T @ X0t
        iXX, singular = _solve_or_ginv(X.T @ X)           # :760-761
        XY = X.T @ yv                                     # :762
        X0Y = XY
        iX0X0 = iXX
    beta = iXX.T @ XY                                     # :772
    if K is None:                                         # :865-870
        YY = float(yt @ yt)
        ves = (YY - float(beta @ XY)) / (n - q0)
        REMLs = -0.5 * (n - q0) * math.log(ves) - 0.5 * n - 0.5 * (n - q0) * math.log(2 * math.pi)
        vgs = 0.0
    beta_cv = beta.copy()                                 # :913
    X_beta = X @ beta                                     # :914
    if K is not None:                                     # :916-921
        r0 = U.T @ (yv - X_beta)
        logLM_Base = 0.5 * (-n * math.log(((2 * math.pi) / n) * float(r0 @ r0))
                            - float(np.sum(np.log(eig_full_plus_delta))) - n)
    else:                                                 # :922-925
        r0 = yv - X_beta
        logLM_Base = 0.5 * (-n * math.log(((2 * math.pi) / n) * float(r0 @ r0)) - n)
    rsquare_base_init = 1 - math.exp(-(2.0 / n) * (logLM_Base - logL0))   # :926

    # ---- i = 1..m
    x_prev = None
    for i in range(m):
        if progress is not None and (i + 1) % 1000 == 0:
            progress(i + 1)                               # :402
        xv = np.asarray(xs[:, i], dtype=np.float64)       # :477
        ns = n
        ss = float(xv.sum())                              # :484
        maf[i] = min(0.5 * ss / ns, 1 - 0.5 * ss / ns)    # :486
        nobs[i] = ns                                      # :487
        if xv.min() == xv.max():                          # :499-511
            ps[i] = 1.0
            x_prev = xv
            continue
        if x_prev is not None and i > 0 and np.array_equal(x_prev, xv):   # :512-527
            for arr in (dfs, stats, effect_est, ps, rsquare, rsquare_base, df, tvalue, stderr, se_clean, logLM_all):
                arr[i] = arr[i - 1]
            x_prev = xv
            continue
        dfs[i] = n - q1                                   # :586
        Xfull = np.column_stack([X0, xv])                 # :588
        if K is not None:
            xst = U.T @ xv                                # :634  (n x n GEMV #1)
            X0Xst = X0t.T @ xst                           # :653
            XstX0 = X0Xst[None, :]                        # :654
            xstxst = float(xst @ xst)                     # :655
            xsY = float(xst @ yt)                         # :682
            XY = np.concatenate([X0Y, [xsY]])             # :683
            B22 = xstxst - float((XstX0 @ iX0X0 @ X0Xst)[0])   # :736
            invB22 = 1.0 / B22                            # :737
            B21 = XstX0 @ iX0X0.T                         # :739 tcrossprod(XstX0, iX0X0)
            NeginvB22B21 = (-invB22) * B21                # :740
            B11 = iX0X0 + invB22 * (B21.T @ B21)          # :742
            iXX = np.empty((q1, q1))
            iXX[:q0, :q0] = B11                           # :745
            iXX[q0, q0] = 1.0 / B22                       # :746
            iXX[q0, :q0] = NeginvB22B21                   # :747
            iXX[:q0, q0] = NeginvB22B21                   # :748
        else:
            iXX, _ = _solve_or_ginv(Xfull.T @ Xfull)      # :760-761
            XY = Xfull.T @ yv                             # :762
        beta = iXX.T @ XY                                 # :772
        if K is not None:
            stats[i] = beta[q1 - 1] / math.sqrt(iXX[q1 - 1, q1 - 1] * vgs)   # :936
            se_clean[i] = math.sqrt(iXX[q1 - 1, q1 - 1] * vgs)
        else:
            stats[i] = beta[q1 - 1] / math.sqrt(iXX[q1 - 1, q1 - 1] * ves)   # :937
            se_clean[i] = math.sqrt(iXX[q1 - 1, q1 - 1] * ves)
        effect_est[i] = beta[q1 - 1]                      # :938
        p = float(rmath.pt_upper_two_sided(abs(stats[i]), dfs[i]))   # :939
        ps[i] = 1.0 if math.isnan(p) else p               # :940
        X_beta = Xfull @ beta                             # :955
        if K is not None:
            r = U.T @ (yv - X_beta)                       # :957  (n x n GEMV #2)
            logLM = 0.5 * (-n * math.log(((2 * math.pi) / n) * float(r @ r))
                           - float(np.sum(np.log(eig_full_plus_delta))) - n)   # :958-959
        else:
            r = yv - X_beta                               # :962
            logLM = 0.5 * (-n * math.log(((2 * math.pi) / n) * float(r @ r)) - n)   # :963
        logLM_all[i] = logLM
        rsquare_base[i] = rsquare_base_init               # :966
        rsquare[i] = 1 - math.exp(-(2.0 / n) * (logLM - logL0))   # :967
        df[i] = dfs[i]                                    # :970
        tvalue[i] = stats[i]                              # :971
        # :972  stderr = beta[ncol(CVI)+1] / t  (1-based index into beta)
        if ncol_CVI is not None and ncol_CVI + 1 <= beta.shape[0]:
            stderr[i] = beta[ncol_CVI] / stats[i]
        else:
            stderr[i] = np.nan
        x_prev = xv                                       # :977

    return P3DResult(ps=ps, stats=stats, effect_est=effect_est, rsquare_base=rsquare_base,
                     rsquare=rsquare, dfs=dfs, df=df, tvalue=tvalue, stderr=stderr, maf=maf,
                     nobs=nobs, REMLs=-2.0 * REMLs, vgs=vgs, ves=ves, delta=delta,
                     logLM=logLM_all, logLM_Base=logLM_Base, logL0=logL0, effect_cv=beta_cv,
                     se_clean=se_clean, singular_X0=singular)


# --------------------------------------------------------------------------
# i = 0 BLUP / PEV / BLUE block  (GAPIT.EMMAxP3D.R:779-861), Z = NULL
# --------------------------------------------------------------------------
def _r_solve(A):
    """R's solve(A): LAPACK dgesv, error when the reciprocal condition number (1-norm) is below
    .Machine$double.eps ("system is computationally singular")."""
    A = np.asarray(A, dtype=np.float64)
    with np.errstate(all="ignore"):
        rcond = 1.0 / np.linalg.cond(A, 1)
    if not np.isfinite(rcond) or rcond < np.finfo(float).eps:
        raise np.linalg.LinAlgError("system is computationally singular")
    return np.linalg.inv(A)


def _mass_ginv(A):
    """MASS::ginv(A): SVD pseudo-inverse with tol = sqrt(.Machine$double.eps)."""
    return np.linalg.pinv(np.asarray(A, dtype=np.float64), rcond=math.sqrt(np.finfo(float).eps))


def blup_pev(ys, X0, K, eig_L, reml, CVI=None, CV_Inheritance=None):
    """Dense restatement of GAPIT.EMMAxP3D.R:779-861 (the reference's two n x n inversions kept)."""
    ys = np.asarray(ys, dtype=np.float64).reshape(-1)
    X0 = np.asarray(X0, dtype=np.float64)
    K = np.asarray(K, dtype=np.float64)
    n = ys.shape[0]
    vgs, ves, delta = reml["vg"], reml["ve"], reml["delta"]
    lam = np.asarray(eig_L["values"], dtype=np.float64)
    U = np.asarray(eig_L["vectors"], dtype=np.float64) * np.sqrt(1.0 / (lam + delta))[None, :]   # :224
    yt = U.T @ ys
    Xt = U.T @ X0
    iX0X0, _ = _solve_or_ginv(Xt.T @ Xt)
    beta = iX0X0.T @ (Xt.T @ yt)                               # :772
    XtimesBetaHat = X0 @ beta                                  # :788
    YminusXtimesBetaHat = ys - XtimesBetaHat                   # :790
    Dt = U.T @ YminusXtimesBetaHat                             # :792
    Zt = U.T                                                   # :795
    BLUP = K @ (Zt.T @ Dt)                                     # :803
    BLUP_Plus_Mean = beta[0] + BLUP                            # :818-819
    try:                                                       # :824-825
        C11 = vgs * _r_solve(Xt.T @ Xt)
    except np.linalg.LinAlgError:
        C11 = vgs * _mass_ginv(Xt.T @ Xt)
    C21 = -K @ (Zt.T @ Xt) @ C11                               # :827
    # :828-829  Kinv = try(solve(K)); on error ginv(K).  R's solve() errors when LAPACK's 1-norm rcond estimate is
    # below .Machine$double.eps -- a coin toss for a kinship that is singular by construction (centring by 2p gives
    # K 1 = 0 exactly).  Both this oracle and the CUDA path decide from the spectrum instead:
    # |lambda|min / |lambda|max < eps  =>  pseudo-inverse.  (Documented deviation)
    alam = np.abs(lam)
    used_ginv = not (alam.min() / alam.max() >= np.finfo(float).eps)
    Kinv = _mass_ginv(K) if used_ginv else np.linalg.inv(K)
    term0 = np.diag(np.full(n, 1.0 / ves))                     # :832
    try:                                                       # :834-835
        term1 = _r_solve(term0 + Kinv / vgs)
    except np.linalg.LinAlgError:
        term1 = _mass_ginv(term0 + Kinv / vgs)
    term2 = C21 @ (Xt.T @ Zt) @ K                              # :837
    C22 = term1 - term2                                        # :838
    PEV = np.diag(C22).copy()                                  # :839
    BLUE = None
    if CVI is not None:                                        # :841-854
        CVI = np.asarray(CVI, dtype=np.float64)
        XCV = np.column_stack([np.ones(n), CVI[:, 1:]])
        beta_inh = beta
        if CV_Inheritance is not None:
            XCV = XCV[:, : 1 + CV_Inheritance]
            beta_inh = beta[: 1 + CV_Inheritance]
        if beta.shape[0] == 1:
            XCV = X0
        BLUE = XCV @ beta_inh if XCV.shape[1] == beta_inh.shape[0] else None
    return {"BLUP": BLUP, "BLUP_Plus_Mean": BLUP_Plus_Mean, "PEV": PEV, "BLUE": BLUE, "beta": beta,
            "used_ginv_K": used_ginv}


# --------------------------------------------------------------------------
# GAPIT.Numericalization  (GAPIT.Numericalization.R:1-112) and the HapMap loop
# around it (GAPIT.HapMap.R:19-37)
# --------------------------------------------------------------------------
def _r_order(values, decreasing=False):
    """R's order(): stable, ties keep their original order in both directions (radix method)."""
    idx = list(range(len(values)))
    return sorted(idx, key=lambda i: -values[i] if decreasing else values[i])


def numericalization(x, bit=2, effect="Add", impute="None", Major_allele_zero=False):
    """One marker: list of genotype strings -> float vector (NaN = NA).  Line references are to
    GAPIT.Numericalization.R.  Factor levels are sorted by byte value (R's sort in the C locale; identical to any
    locale for the upper-case nucleotide codes HapMap files hold)."""
    x = [str(v) for v in x]
    if bit == 1:                                                            # :8-14
        x = ["N" if v in ("X", "-", "+", "/") else v for v in x]
        x = ["Z" if v == "K" else v for v in x]
    if bit == 2:                                                            # :16-22
        x = ["N" if v in ("XX", "--", "++", "//", "NN") else v for v in x]
    n = len(x)
    lev = sorted(set(x) - {"N"}, key=lambda s: s.encode())                  # :25-26
    ln = len(lev)                                                           # :28
    count = [sum(1 for v in x if v == l) for l in lev]                      # :34-37
    if Major_allele_zero and 1 < ln <= 3:                                   # :41-62
        ct = [(count[i], i) for i in range(ln)]
        if bit == 1:
            if ln == 3:
                ct = ct[:2]
            o = _r_order([c for c, _ in ct], decreasing=True)
            order = [ct[i][1] for i in o] + ([2] if ln == 3 else [])
        else:
            if ln == 3:
                ct = [ct[0], ct[2]]
            o = _r_order([c for c, _ in ct], decreasing=True)
            order = [ct[o[0]][1], 1, ct[o[1]][1]] if ln == 3 else [ct[i][1] for i in o]
        count = [count[i] for i in order]
        lev = [lev[i] for i in order]
    if bit == 1 and ln == 3:                                                # :67-71
        count[1], count[2] = count[2], count[1]
    position = [p + 1 for p in _r_order(count)]                             # :72 (1-based)
    nan = float("nan")
    if ln <= 1 or ln > 3:                                                   # :76
        out = [0.0] * n
    elif ln == 2:                                                           # :79
        out = [nan if v == "N" else (0.0 if v == lev[0] else 2.0) for v in x]
    elif bit == 1:                                                          # :82-83
        out = [nan if v == "N" else (0.0 if v == lev[0] else (1.0 if v == lev[2] else 2.0)) for v in x]
    else:                                                                   # :85
        out = [nan if v == "N" else (0.0 if v == lev[0] else (2.0 if v == lev[2] else 1.0)) for v in x]
    out = np.array(out, dtype=np.float64)
    na = np.isnan(out)
    if impute == "Middle":                                                  # :92
        out[na] = 1.0
    if ln == 3:                                                             # :94-97
        if impute == "Minor":
            out[na] = position[0] - 1
        if impute == "Major":
            out[na] = position[ln - 1] - 1
    elif ln >= 1:                                                           # :99-101 (position is empty for ln == 0)
        if impute == "Minor":
            out[na] = 2 * (position[0] - 1)
        if impute == "Major":
            out[na] = 2 * (position[ln - 1] - 1)
    na = np.isnan(out)
    if effect == "Dom":                                                     # :105
        out = np.where(na, np.nan, np.where(out == 1.0, 1.0, 0.0))
    if effect == "Left":                                                    # :106
        out[out == 1.0] = 0.0
    if effect == "Right":                                                   # :107
        out[out == 1.0] = 2.0
    return out


def hapmap(rows, SNP_effect="Add", SNP_impute="Middle", heading=True, Major_allele_zero=False):
    """GAPIT.HapMap (GAPIT.HapMap.R:1-66) on a table given as a list of rows (each a list of strings, 11 leading
    HapMap columns).  Returns GT (taxa), GD (n x m float, NaN = NA), GI (SNP, Chromosome, Position)."""
    body = rows[1:] if heading else rows
    GT = list(rows[0][11:]) if heading else None                            # :14-16
    GI = [(r[0], r[2], r[3]) for r in body]                                 # :15,:18
    bit = len(str(rows[1][11]))                                             # :28  nchar(G[2,12])
    GD = np.column_stack([numericalization(r[11:], bit=bit, effect=SNP_effect, impute=SNP_impute,
                                           Major_allele_zero=Major_allele_zero) for r in body])   # :34-38
    return {"GT": GT, "GD": GD, "GI": GI}


# --------------------------------------------------------------------------
# GAPIT.kinship.ZHANG  (GAPIT.kinship.ZHANG.R:1-66)  and  prcomp of GAPIT.PCA.R:10
# --------------------------------------------------------------------------
def kinship_zhang(snps: np.ndarray) -> np.ndarray:
    """n x m dosages -> n x n ZHANG relationship, line by line."""
    snps = np.asarray(snps, dtype=np.float64)
    fa = snps.sum(axis=0) / (2.0 * snps.shape[0])                            # :9
    snps = snps[:, ~((fa >= 1.0) | (fa <= 0.0))]                             # :10-11
    het = 1.0 - np.abs(snps - 1.0)                                           # :13
    fi = het.sum(axis=1) / (2.0 * snps.shape[1])                             # :14-15
    inbreeding = 1.0 - fi.min()                                              # :16
    n = snps.shape[0]
    Z = snps.T - snps.mean(axis=0)[:, None]                                  # :21-23
    K = Z.T @ Z                                                              # :26
    d = np.diag(K).copy()                                                    # :31-35
    DL, DU, floor = d.min(), d.max(), K.min()                                # :36-38
    top = 1.0 + inbreeding                                                   # :42
    K = top * (K - floor) / (DU - floor)                                     # :43
    Dmin = top * (DL - floor) / (DU - floor)                                 # :44
    diag = np.eye(n, dtype=bool)
    if Dmin < 1.0:                                                           # :47-51
        K[diag] = (K[diag] - Dmin + 1.0) / ((top + 1.0 - Dmin) * .5)
        K[~diag] = K[~diag] * (1.0 / Dmin)
    Omax = K[~diag].max()                                                    # :54
    if Omax > top:                                                           # :55-58
        K[~diag] = K[~diag] * (top / Omax)
    return K


def prcomp(X: np.ndarray):
    """stats::prcomp(X) as GAPIT.PCA.R:10 calls it (centre, no scaling): svd of the centred matrix.
    Returns (x = scores n x min(n,m), sdev^2)."""
    X = np.asarray(X, dtype=np.float64)
    Xc = X - X.mean(axis=0, keepdims=True)
    u, s, _ = np.linalg.svd(Xc, full_matrices=False)
    return u * s, s * s / (X.shape[0] - 1.0)


# ---------------------------------------------------------------------------------------------------------
# Compression glue around the scan: GAPIT.Main.R:386-400 calls these three unconditionally
# before GAPIT.EMMAxP3D.  Tables are held as plain lists / arrays: KI = (names, n x n matrix); Z = (row taxa,
# column taxa, matrix) for the reference's data.frame whose first row / column carry the names; GA rows =
# (taxon, group); GAU rows = (taxon, group, block indicator, ID).  Character sorting follows R in the C locale.

def _first_appearance(labels):
    """cutree numbers the clusters in order of first appearance of their members (stats::cutree)."""
    seen, out = {}, []
    for v in labels:
        if v not in seen:
            seen[v] = len(seen) + 1
        out.append(seen[v])
    return np.asarray(out, dtype=np.int64)


def compress(line_names, KI, kinship_cluster="average", kinship_group="Mean", GN=None):
    """GAPIT.Compress (GAPIT.Compress.R:12-112): dist() between the ROWS of the kinship (:37), hclust (:38),
    cutree(k = GN) (:45), group kinship by Mean (:52-63) or Max / Min / Median (:72-96).
    Returns GA (list of (name, group)) and KG (GN x GN)."""
    from scipy.cluster.hierarchy import linkage, cut_tree
    from scipy.spatial.distance import pdist
    KI = np.asarray(KI, dtype=np.float64)
    n = KI.shape[0]
    GN = n if GN is None else int(GN)
    if n == 1:
        membership = np.ones(1, dtype=np.int64)
    else:
        method = {"average": "average", "complete": "complete", "single": "single", "ward": "ward",
                  "mcquitty": "weighted", "median": "median", "centroid": "centroid"}[kinship_cluster]
        tree = linkage(pdist(KI, metric="euclidean"), method=method)            # :37-38
        membership = _first_appearance(cut_tree(tree, n_clusters=GN)[:, 0])     # :45
    ng = int(membership.max())                                                   # :56
    b = np.zeros((n, ng))
    b[np.arange(n), membership - 1] = 1.0                                        # :57 diag(n)[x,]
    if kinship_group == "Mean":
        KG = (b.T @ KI @ b) / (b.T @ np.ones_like(KI) @ b)                       # :59-61
    else:
        f = {"Max": np.max, "Min": np.min, "Median": np.median}[kinship_group]   # :91-96 tapply
        KG = np.empty((ng, ng))
        for a in range(ng):
            for c in range(ng):
                KG[a, c] = f(KI[np.ix_(membership == a + 1, membership == c + 1)])
    GA = [(str(nm), int(g)) for nm, g in zip(line_names, membership)]            # :104
    return GA, KG


def block(Z_col_taxa, GA, KG):
    """GAPIT.Block (GAPIT.Block.R:14-64): split the group kinship into groups with (KW) and without (KO) a phenotyped
    member.  Z_col_taxa = Z[1,-1].  Returns GAU rows (taxon, group, (indiv block + group block)/2, ID) in merge()'s
    order (sorted by the group key AS CHARACTER, ties in x order), KW, KO, KWO."""
    zset = set(str(t) for t in Z_col_taxa)
    taxa = [t for t, _ in GA]
    grp = [g for _, g in GA]
    blk = [1 if t in zset else 2 for t in taxa]                                  # :19-22
    grp12 = list(dict.fromkeys(grp))                                             # :33 unique, first appearance
    grp1 = list(dict.fromkeys(g for g, b in zip(grp, blk) if b == 1))            # :34
    grp2 = [g for g in grp12 if g not in set(grp1)]                              # :35 setdiff keeps x order
    gb = {g: (1, i + 1) for i, g in enumerate(grp1)}                             # :39-42 (grp, block, ID)
    gb.update({g: (2, i + 1) for i, g in enumerate(grp2)})
    order_block = sorted(range(len(blk)), key=lambda i: blk[i])                  # :44 order() is stable
    rows = [(taxa[i], grp[i], blk[i]) for i in order_block]
    rows.sort(key=lambda r: str(r[1]))                                           # :51 merge sorts on the key column
    GAU = [(t, g, (b + gb[g][0]) / 2.0, gb[g][1]) for t, g, b in rows]           # :53-55
    i1 = np.asarray(grp1, dtype=np.int64) - 1
    i2 = np.asarray(grp2, dtype=np.int64) - 1
    KG = np.asarray(KG)
    return GAU, KG[np.ix_(i1, i1)], KG[np.ix_(i2, i2)], KG[np.ix_(i1, i2)]       # :56-58


def zmatrix_compress(Z_row_taxa, Z_col_taxa, Zmat, GAU):
    """GAPIT.ZmatrixCompress (GAPIT.ZmatrixCompress.R:13-45): individual -> group incidence Z %*% DS, columns = IDs
    of the phenotyped block, rows sorted by taxon.  Returns (row taxa sorted, matrix)."""
    zset = set(str(t) for t in Z_col_taxa)
    GAU0 = [r for r in GAU if r[0] in zset]                                      # :17
    GAU1 = sorted(GAU0, key=lambda r: r[0])                                      # :18-19
    n = max(r[3] for r in GAU1 if r[2] < 2)                                      # :21-22
    x = np.asarray([r[3] for r in GAU1], dtype=np.int64)                         # :23
    DS = np.eye(n)[x - 1, :]                                                     # :24
    order_Z = sorted(range(len(Z_col_taxa)), key=lambda i: str(Z_col_taxa[i]))   # :27
    Zs = np.asarray(Zmat, dtype=np.float64)[:, order_Z]                          # :28-33
    out = Zs @ DS                                                                # :34
    order_rows = sorted(range(len(Z_row_taxa)), key=lambda i: str(Z_row_taxa[i]))  # :41
    return [str(Z_row_taxa[i]) for i in order_rows], out[order_rows, :]


# ---------------------------------------------------------------------------------------------------------
# Genotype file formats beside HapMap.  The reference has no .bed reader; the restatement
# follows the PLINK 1 binary specification (variant-major: magic 6C 1B 01, then per variant ceil(n/4) bytes, sample i
# in bits 2(i%4).. of byte i/4; 00 = homozygous A1, 01 = missing, 10 = heterozygous, 11 = homozygous A2).
def read_bed(bed, n, count="A1"):
    """-> int8 (n, m) dosages (copies of `count`), -1 for missing."""
    b = np.frombuffer(bytes(bed), dtype=np.uint8)
    assert b[0] == 0x6C and b[1] == 0x1B and b[2] == 0x01, "not a variant-major PLINK .bed"
    nb = (n + 3) // 4
    m = (b.size - 3) // nb
    rows = b[3:3 + m * nb].reshape(m, nb)
    codes = np.stack([(rows >> (2 * j)) & 3 for j in range(4)], axis=2).reshape(m, nb * 4)[:, :n]
    a1 = np.array([2, -1, 1, 0], dtype=np.int8)
    a2 = np.array([0, -1, 1, 2], dtype=np.int8)
    return (a1 if count == "A1" else a2)[codes].T


def read_numeric_text(text, heading=True):
    """GAPIT's numeric genotype table (GAPIT.Fragment.R:100-128: read.table(head = TRUE), GT = first column, GD = the
    rest): -> (GD float (n, m) with NaN for NA, GT, marker names)."""
    lines = [ln for ln in bytes(text).decode().replace(",", " ").splitlines() if ln.strip()]
    names = lines[0].split()[1:] if heading else None
    rows = [ln.split() for ln in lines[1 if heading else 0:]]
    GT = [r[0] for r in rows]

    def val(tok):
        return float("nan") if tok in ("NA", "NaN", "nan", ".", "-") else float(tok)
    GD = np.array([[val(t) for t in r[1:]] for r in rows], dtype=np.float64)
    return GD, GT, names
