"""Independent derivations that pin the oracle (TEST INFRASTRUCTURE ONLY).

The reference has no golden vectors, so the oracle in
``gapit_oracle.py`` is pinned by results derived a different way:

* ``kinship_exact``  -- VanRaden kinship from exact integer sums and one
  rational division per element (python ints / fractions), no dgemm.
* ``gls_dense``      -- beta, SE, t, RSS of the marker model by dense GLS with
  V = K + delta*I via Cholesky, never touching the eigen-rotation the
  reference (and the GPU path) use.
* ``emma_reml_t_row`` -- the t statistic through emma.REML.t's own formulas
  (emma.txt:726-736), a second published derivation of the same quantity.
"""
from __future__ import annotations

from fractions import Fraction

import numpy as np


def kinship_int_parts(snps):
    """Exact integer ingredients of GAPIT.kinship.VanRaden.R:11-32.

    Returns (G, s, sum_c2, D, n) over the columns the reference keeps
    (:11-13 drops c==0 and c==2n):
      G[i,j] = sum_k x_ik x_jk         c_k = sum_i x_ik
      s[i]   = sum_k c_k x_ik          sum_c2 = sum_k c_k^2
      D      = sum_k c_k (2n - c_k)    so that adj = D / (2 n^2)
    and K_ij = 2 (n^2 G_ij - n (s_i + s_j) + sum_c2) / D.
    """
    x = np.asarray(snps).astype(np.int64)
    n = x.shape[0]
    c = x.sum(axis=0)
    keep = (c > 0) & (c < 2 * n)
    x = x[:, keep]
    c = c[keep]
    G = x @ x.T
    s = x @ c
    sum_c2 = int((c * c).sum())
    D = int((c * (2 * n - c)).sum())
    return G, s, sum_c2, D, n


def kinship_exact(snps, as_fraction=False):
    """VanRaden kinship with exact rational arithmetic, rounded once to fp64."""
    G, s, sum_c2, D, n = kinship_int_parts(snps)
    nn = G.shape[0]
    out = np.empty((nn, nn), dtype=object if as_fraction else np.float64)
    for i in range(nn):
        for j in range(nn):
            num = 2 * (n * n * int(G[i, j]) - n * (int(s[i]) + int(s[j])) + sum_c2)
            fr = Fraction(num, D)
            out[i, j] = fr if as_fraction else float(fr)
    return out


def gls_dense(y, X0, x, K, delta, vg):
    """Marker model y = [X0 x] b + e, Var(e) = vg (K + delta I), solved densely."""
    y = np.asarray(y, dtype=np.float64)
    X = np.column_stack([np.asarray(X0, dtype=np.float64), np.asarray(x, dtype=np.float64)])
    n = y.shape[0]
    H = np.asarray(K, dtype=np.float64) + delta * np.eye(n)
    L = np.linalg.cholesky(H)
    Xw = np.linalg.solve(L, X)
    yw = np.linalg.solve(L, y)
    XtX = Xw.T @ Xw
    iXX = np.linalg.inv(XtX)
    beta = iXX @ (Xw.T @ yw)
    r = yw - Xw @ beta
    rss = float(r @ r)
    se = float(np.sqrt(iXX[-1, -1] * vg))
    return {"beta": beta, "se": se, "t": float(beta[-1] / se), "rss": rss,
            "logdet": 2.0 * float(np.sum(np.log(np.diag(L))))}


def emma_reml_t_row(y, X0, x, eig_L_values, eig_L_vectors, delta, vg):
    """emma.REML.t's inner formulas (emma.txt:726-736) for one marker."""
    X = np.column_stack([X0, x])
    U = eig_L_vectors * np.sqrt(1.0 / (eig_L_values + delta))[None, :]
    yt = U.T @ y
    Xt = U.T @ X
    iXX = np.linalg.inv(Xt.T @ Xt)
    beta = iXX @ (Xt.T @ yt)
    return float(beta[-1] / np.sqrt(iXX[-1, -1] * vg)), beta
