"""R-base arithmetic the GAPIT hot path leans on, restated for the oracle.

TEST INFRASTRUCTURE ONLY.  Nothing under ``oracle/`` is on the product path;
only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s cpu_baseline /
``--impl reference`` legs may import it.

None of this arithmetic lives in /root/reference: GAPIT calls R base
(``stats::uniroot``, ``stats::pt``, ``solve``, ``MASS::ginv``) whose version the
reference does not pin (``GAPIT.0000.R:27`` only carries "2016.03.01").  The
functions below restate the published algorithms R uses at the reference's
call sites:

* ``zeroin``  -- Brent's root finder as R's ``R_zeroin2`` (netlib ``zeroin.c``)
  implements it, called by ``uniroot`` at ``GAPIT.emma.REMLE.R:50``.
* ``uniroot`` -- the R wrapper's defaults: ``tol = .Machine$double.eps^0.25``,
  ``maxiter = 1000``, f(lower)/f(upper) pre-evaluated.
* ``pt_upper_two_sided`` -- ``2*pt(|t|, df, lower.tail=FALSE)`` as used at
  ``GAPIT.EMMAxP3D.R:939``; R's ``pt`` evaluates it through the regularised
  incomplete beta ``pbeta`` (TOMS 708).  scipy's ``betainc`` is the stand-in.
"""
from __future__ import annotations

import math

import numpy as np
from scipy import special

DBL_EPSILON = np.finfo(np.float64).eps


def zeroin(f, ax, bx, fa, fb, tol, maxit):
    """Brent's method exactly as R_zeroin2 steps it.

    Returns (root, iterations_used, estimated_precision).  ``maxit`` has the
    meaning of R's ``*Maxit`` on entry.
    """
    a, b = ax, bx
    c, fc = a, fa
    if fa == 0.0:
        return a, 0, 0.0
    if fb == 0.0:
        return b, 0, 0.0
    budget = maxit + 1
    it = 0
    while budget > 0:
        budget -= 1
        prev_step = b - a
        if abs(fc) < abs(fb):
            # swap so that b is the best approximation so far
            a, b, c = b, c, b
            fa, fb, fc = fb, fc, fb
        tol_act = 2.0 * DBL_EPSILON * abs(b) + tol / 2.0
        new_step = (c - b) / 2.0
        if abs(new_step) <= tol_act or fb == 0.0:
            return b, it, abs(c - b)
        if abs(prev_step) >= tol_act and abs(fa) > abs(fb):
            cb = c - b
            if a == c:
                # only two distinct points: linear interpolation
                t1 = fb / fa
                p = cb * t1
                q = 1.0 - t1
            else:
                # inverse quadratic interpolation
                q = fa / fc
                t1 = fb / fc
                t2 = fb / fa
                p = t2 * (cb * q * (q - t1) - (b - a) * (t1 - 1.0))
                q = (q - 1.0) * (t1 - 1.0) * (t2 - 1.0)
            if p > 0.0:
                q = -q
            else:
                p = -p
            if p < (0.75 * cb * q - abs(tol_act * q) / 2.0) and p < abs(prev_step * q / 2.0):
                new_step = p / q
        if abs(new_step) < tol_act:
            new_step = tol_act if new_step > 0.0 else -tol_act
        a, fa = b, fb
        b = b + new_step
        fb = f(b)
        it += 1
        if (fb > 0.0 and fc > 0.0) or (fb < 0.0 and fc < 0.0):
            c, fc = a, fa
    return b, -1, -1.0


def uniroot(f, lower, upper, tol=DBL_EPSILON ** 0.25, maxiter=1000):
    """``stats::uniroot(f, lower=, upper=)`` with R's defaults; returns the root."""
    f_lower = f(lower)
    f_upper = f(upper)
    if not (f_lower * f_upper <= 0.0):
        raise ValueError("f() values at end points not of opposite sign")
    root, it, _ = zeroin(f, lower, upper, f_lower, f_upper, tol, maxiter)
    if it < 0:
        raise RuntimeError("uniroot: _NOT_ converged")
    return root


def pt_upper_two_sided(t, df):
    """2 * pt(|t|, df, lower.tail = FALSE)  (GAPIT.EMMAxP3D.R:939).

    For |t| < Inf this is the regularised incomplete beta
    I_{df/(df+t^2)}(df/2, 1/2); R's pt switches between the two pbeta forms at
    n > t^2 only to protect precision, the value is the same function.
    NaN in -> NaN out (the reference then rewrites NA to 1 at :940).
    """
    t = np.asarray(t, dtype=np.float64)
    t2 = t * t
    out = np.empty_like(t)
    small = df > t2
    with np.errstate(invalid="ignore", divide="ignore", over="ignore"):
        # n > t^2 : upper tail of I_{t^2/(n+t^2)}(1/2, n/2)
        xs = t2 / (df + t2)
        out_small = special.betaincc(0.5, df / 2.0, xs)
        # else    : lower tail of I_{n/(n+t^2)}(n/2, 1/2)
        xl = 1.0 / (1.0 + (t / df) * t)
        out_large = special.betainc(df / 2.0, 0.5, xl)
    out = np.where(small, out_small, out_large)
    out = np.where(np.isinf(t), 0.0, out)
    out = np.where(np.isnan(t), np.nan, out)
    return out


def log10_pt_upper_two_sided_mp(t, df, dps=50):
    """High-precision log10 of the same tail, for the far-tail parity checks."""
    import mpmath as mp

    with mp.workdps(dps):
        t = mp.mpf(float(t))
        df = mp.mpf(float(df))
        x = df / (df + t * t)
        val = mp.betainc(df / 2, mp.mpf(1) / 2, 0, x, regularized=True)
        return float(mp.log10(val))


def lbeta(a, b):
    return math.lgamma(a) + math.lgamma(b) - math.lgamma(a + b)
