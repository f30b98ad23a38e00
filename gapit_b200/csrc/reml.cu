// REML variance components, Z = NULL branch of GAPIT.emma.REMLE (GAPIT.emma.REMLE.R:21-54, 96-110)
// with emma.delta.REML.{LL,dLL}.wo.Z (emma.txt:145-164).  Host code: O((n-q) * 101) after the
// eigendecomposition.  The root finder restates the Brent iteration R's uniroot() runs (R_zeroin2,
// netlib zeroin.c: tol = DBL_EPSILON^0.25, maxiter = 1000, end-point values pre-evaluated).
#include <cfloat>
#include <cmath>
#include <atomic>
#include <thread>
#include <vector>

#include "common.cuh"

namespace gapit {

struct RemlFn {
  const double* lambda;
  const double* etasq;
  int64_t nq;
  // emma.delta.REML.LL.wo.Z (emma.txt:145-149)
  double LL(double logdelta) const {
    const double delta = std::exp(logdelta);
    double s1 = 0.0, s2 = 0.0;
    for (int64_t i = 0; i < nq; ++i) {
      const double ld = lambda[i] + delta;
      s1 += etasq[i] / ld;
      s2 += std::log(ld);
    }
    const double dnq = static_cast<double>(nq);
    return 0.5 * (dnq * (std::log(dnq / (2.0 * M_PI)) - 1.0 - std::log(s1)) - s2);
  }
  // emma.delta.REML.dLL.wo.Z (emma.txt:158-164); `scaled` adds the delta factor of the grid (:33)
  double dLL(double logdelta, bool scaled) const {
    const double delta = std::exp(logdelta);
    double a = 0.0, b = 0.0, c = 0.0;
    for (int64_t i = 0; i < nq; ++i) {
      const double ld = lambda[i] + delta;
      a += etasq[i] / (ld * ld);
      b += etasq[i] / ld;
      c += 1.0 / ld;
    }
    const double v = 0.5 * (static_cast<double>(nq) * a / b - c);
    return scaled ? delta * v : v;
  }
  double vg(double delta) const {
    double s = 0.0;
    for (int64_t i = 0; i < nq; ++i) s += etasq[i] / (lambda[i] + delta);
    return s / static_cast<double>(nq);
  }
};

template <class Fn>
static double zeroin(const Fn& f, double ax, double bx, double fa, double fb, double tol, int maxit, bool* ok) {
  double a = ax, b = bx, c = a, fc = fa;
  *ok = true;
  if (fa == 0.0) return a;
  if (fb == 0.0) return b;
  int budget = maxit + 1;
  while (budget--) {
    const double prev_step = b - a;
    if (std::fabs(fc) < std::fabs(fb)) {
      a = b; b = c; c = a;
      fa = fb; fb = fc; fc = fa;
    }
    const double tol_act = 2.0 * DBL_EPSILON * std::fabs(b) + tol / 2.0;
    double new_step = (c - b) / 2.0;
    if (std::fabs(new_step) <= tol_act || fb == 0.0) return b;
    if (std::fabs(prev_step) >= tol_act && std::fabs(fa) > std::fabs(fb)) {
      double p, q;
      const double cb = c - b;
      if (a == c) {
        const double t1 = fb / fa;
        p = cb * t1;
        q = 1.0 - t1;
      } else {
        q = fa / fc;
        const double t1 = fb / fc, t2 = fb / fa;
        p = t2 * (cb * q * (q - t1) - (b - a) * (t1 - 1.0));
        q = (q - 1.0) * (t1 - 1.0) * (t2 - 1.0);
      }
      if (p > 0.0) q = -q; else p = -p;
      if (p < (0.75 * cb * q - std::fabs(tol_act * q) / 2.0) && p < std::fabs(prev_step * q / 2.0)) new_step = p / q;
    }
    if (std::fabs(new_step) < tol_act) new_step = (new_step > 0.0) ? tol_act : -tol_act;
    a = b; fa = fb;
    b += new_step;
    fb = f.dLL(b, false);
    if ((fb > 0.0 && fc > 0.0) || (fb < 0.0 && fc < 0.0)) {
      c = a; fc = fa;
    }
  }
  *ok = false;
  return b;
}

// GAPIT.emma.REMLE.R:27-54, 96-110 over any function object with LL / dLL / vg: the grid of log(delta), the bracketed
// uniroot calls, the boundary candidates and the which.max rule.
template <class Fn>
static int reml_search(const Fn& f, int ngrids, double llim, double ulim, double esp, gapit_reml_out* out) {
  const int m = ngrids + 1;
  std::vector<double> logdelta(m), dLL(m);
  for (int i = 0; i < m; ++i) {
    logdelta[i] = static_cast<double>(i) / ngrids * (ulim - llim) + llim;  // :27
    dLL[i] = f.dLL(logdelta[i], true);                                       // :33
  }
  std::vector<double> optlogdelta, optLL;
  if (dLL[0] < esp) {  // :37
    optlogdelta.push_back(llim);
    optLL.push_back(f.LL(llim));
  }
  if (dLL[m - 2] > 0.0 - esp) {  // :41 (sic: dLL[m-1] in 1-based R)
    optlogdelta.push_back(ulim);
    optLL.push_back(f.LL(ulim));
  }
  for (int i = 0; i < m - 1; ++i) {  // :46-54
    if (dLL[i] * dLL[i + 1] < 0.0 && dLL[i] > 0.0 && dLL[i + 1] < 0.0) {
      bool ok = true;
      const double fl = f.dLL(logdelta[i], false), fu = f.dLL(logdelta[i + 1], false);
      const double root = zeroin(f, logdelta[i], logdelta[i + 1], fl, fu, std::pow(DBL_EPSILON, 0.25), 1000, &ok);
      if (!ok) return fail(GAPIT_ERR_NUMERIC, "gapit_reml: root finder did not converge");
      optlogdelta.push_back(root);
      optLL.push_back(f.LL(root));
    }
  }
  if (optLL.empty()) return fail(GAPIT_ERR_NUMERIC, "gapit_reml: no REML optimum found on the grid");
  // :96 which.max skips NaN; :99 replaces NaN by the worst finite value; :101 max
  int best = -1;
  for (size_t i = 0; i < optLL.size(); ++i)
    if (!std::isnan(optLL[i]) && (best < 0 || optLL[i] > optLL[best])) best = static_cast<int>(i);
  if (best < 0) return fail(GAPIT_ERR_NUMERIC, "gapit_reml: all candidate likelihoods are NaN");
  const double maxdelta = std::exp(optlogdelta[best]);
  out->REML = optLL[best];
  out->delta = maxdelta;
  out->vg = f.vg(maxdelta);       // :103
  out->ve = out->vg * maxdelta;   // :108
  return GAPIT_OK;
}

// The same restricted likelihood evaluated from the spectrum of K alone (eig.L), so that the second O(n^3)
// eigendecomposition of the reference (eig.R of S(K+I)S, emma.txt:82-89) is not needed.  With d = 1/(lambda+delta),
// yt = V'y, Xt = V'X, A = Xt'D Xt, r = yt - Xt A^-1 Xt'D yt and P = S(S H S)^+ S, H = K + delta I:
//   sum eta^2/(lambdaR+delta)   = y'P y  = sum d r^2
//   sum eta^2/(lambdaR+delta)^2 = y'PP y = sum d^2 r^2
//   sum 1/(lambdaR+delta)       = tr P   = sum d - tr(A^-1 Xt'D^2 Xt)
//   sum log(lambdaR+delta)      = sum log(lambda+delta) + log|A| - log|X'X|
// which are exactly the four sums emma.delta.REML.{LL,dLL}.wo.Z take over eig.R (emma.txt:145-164).
struct RemlFnL {
  const double* lambda;  // n eigenvalues of K
  const double* yt;      // n
  const double* Xt;      // n x q column-major
  int64_t n;
  int q;
  double logdetXX;       // log|X'X| (= log|Xt'Xt|, V orthogonal)
  struct Sums {
    double r1, r2, trP, logdet;
    bool ok;
  };
  // Cholesky of a q x q SPD matrix (row-major, lower); false if not positive definite
  static bool chol(int q, std::vector<double>& A) {
    for (int i = 0; i < q; ++i) {
      for (int j = 0; j <= i; ++j) {
        double s = A[i * q + j];
        for (int k = 0; k < j; ++k) s -= A[i * q + k] * A[j * q + k];
        if (i == j) {
          if (!(s > 0.0)) return false;
          A[i * q + i] = std::sqrt(s);
        } else {
          A[i * q + j] = s / A[j * q + j];
        }
      }
    }
    return true;
  }
  static void chol_solve(int q, const std::vector<double>& L, double* b) {
    for (int i = 0; i < q; ++i) {
      double s = b[i];
      for (int k = 0; k < i; ++k) s -= L[i * q + k] * b[k];
      b[i] = s / L[i * q + i];
    }
    for (int i = q - 1; i >= 0; --i) {
      double s = b[i];
      for (int k = i + 1; k < q; ++k) s -= L[k * q + i] * b[k];
      b[i] = s / L[i * q + i];
    }
  }
  Sums sums(double delta) const {
    Sums o{0.0, 0.0, 0.0, 0.0, true};
    std::vector<double> A(q * q, 0.0), A2(q * q, 0.0), b(q, 0.0);
    double sd = 0.0, sl = 0.0;
    for (int64_t i = 0; i < n; ++i) {
      const double ld = lambda[i] + delta;
      const double d = 1.0 / ld;
      sd += d;
      sl += std::log(ld);
      for (int a = 0; a < q; ++a) {
        const double xa = Xt[a * n + i] * d;
        b[a] += xa * yt[i];
        for (int c = 0; c <= a; ++c) {
          const double v = xa * Xt[c * n + i];
          A[a * q + c] += v;
          A2[a * q + c] += v * d;
        }
      }
    }
    for (int a = 0; a < q; ++a)
      for (int c = a + 1; c < q; ++c) {
        A[a * q + c] = A[c * q + a];
        A2[a * q + c] = A2[c * q + a];
      }
    std::vector<double> L(A);
    if (!chol(q, L)) {
      o.ok = false;
      return o;
    }
    std::vector<double> beta(b);
    chol_solve(q, L, beta.data());
    double r1 = 0.0, r2 = 0.0;
    for (int64_t i = 0; i < n; ++i) {
      double r = yt[i];
      for (int a = 0; a < q; ++a) r -= Xt[a * n + i] * beta[a];
      const double d = 1.0 / (lambda[i] + delta);
      const double dr2 = d * r * r;
      r1 += dr2;
      r2 += dr2 * d;
    }
    double tr = 0.0, ldA = 0.0;
    std::vector<double> col(q);
    for (int c = 0; c < q; ++c) {  // tr(A^-1 A2) column by column
      for (int a = 0; a < q; ++a) col[a] = A2[a * q + c];
      chol_solve(q, L, col.data());
      tr += col[c];
      ldA += 2.0 * std::log(L[c * q + c]);
    }
    o.r1 = r1;
    o.r2 = r2;
    o.trP = sd - tr;
    o.logdet = sl + ldA - logdetXX;
    return o;
  }
  double LL(double logdelta) const {
    const Sums s = sums(std::exp(logdelta));
    if (!s.ok) return NAN;
    const double dnq = static_cast<double>(n - q);
    return 0.5 * (dnq * (std::log(dnq / (2.0 * M_PI)) - 1.0 - std::log(s.r1)) - s.logdet);
  }
  double dLL(double logdelta, bool scaled) const {
    const double delta = std::exp(logdelta);
    const Sums s = sums(delta);
    if (!s.ok) return NAN;
    const double v = 0.5 * (static_cast<double>(n - q) * s.r2 / s.r1 - s.trP);
    return scaled ? delta * v : v;
  }
  double vg(double delta) const { return sums(delta).r1 / static_cast<double>(n - q); }
};

}  // namespace gapit

using namespace gapit;

extern "C" int gapit_reml(const double* lambda, const double* etas, int64_t nq, int ngrids, double llim, double ulim,
                          double esp, gapit_reml_out* out) {
  if (!lambda || !etas || !out || nq < 1 || ngrids < 2) return fail(GAPIT_ERR_ARG, "gapit_reml: bad argument");
  std::vector<double> etasq(nq);
  for (int64_t i = 0; i < nq; ++i) etasq[i] = etas[i] * etas[i];
  RemlFn f{lambda, etasq.data(), nq};
  return reml_search(f, ngrids, llim, ulim, esp, out);
}

extern "C" int gapit_reml_eigL(const double* lambda, const double* Vty, const double* VtX, int64_t n, int q, int ngrids,
                               double llim, double ulim, double esp, gapit_reml_out* out) {
  if (!lambda || !Vty || !VtX || !out || q < 1 || q > GAPIT_MAX_Q0 || n <= q || ngrids < 2)
    return fail(GAPIT_ERR_ARG, "gapit_reml_eigL: bad argument");
  std::vector<double> XX(q * q, 0.0);
  for (int a = 0; a < q; ++a)
    for (int c = 0; c <= a; ++c) {
      double s = 0.0;
      for (int64_t i = 0; i < n; ++i) s += VtX[a * n + i] * VtX[c * n + i];
      XX[a * q + c] = XX[c * q + a] = s;
    }
  if (!RemlFnL::chol(q, XX)) return fail(GAPIT_ERR_NUMERIC, "gapit_reml_eigL: X'X is singular");
  double ld = 0.0;
  for (int a = 0; a < q; ++a) ld += 2.0 * std::log(XX[a * q + a]);
  RemlFnL f{lambda, Vty, VtX, n, q, ld};
  return reml_search(f, ngrids, llim, ulim, esp, out);
}

// T traits that share X and the spectrum (BASELINE configs[4]): the searches are independent, so they run on the host's
// threads.  VtY: n x T column-major; out: T entries.  A failing trait sets its status in rc_out (may be NULL) and the
// call returns the first failure.
extern "C" int gapit_reml_eigL_multi(const double* lambda, const double* VtY, const double* VtX, int64_t n, int q, int T,
                                     int ngrids, double llim, double ulim, double esp, gapit_reml_out* out, int* rc_out) {
  if (!lambda || !VtY || !VtX || !out || T < 1 || q < 1 || q > GAPIT_MAX_Q0 || n <= q || ngrids < 2)
    return fail(GAPIT_ERR_ARG, "gapit_reml_eigL_multi: bad argument");
  std::vector<double> XX(q * q, 0.0);
  for (int a = 0; a < q; ++a)
    for (int c = 0; c <= a; ++c) {
      double s = 0.0;
      for (int64_t i = 0; i < n; ++i) s += VtX[a * n + i] * VtX[c * n + i];
      XX[a * q + c] = XX[c * q + a] = s;
    }
  if (!RemlFnL::chol(q, XX)) return fail(GAPIT_ERR_NUMERIC, "gapit_reml_eigL_multi: X'X is singular");
  double ld = 0.0;
  for (int a = 0; a < q; ++a) ld += 2.0 * std::log(XX[a * q + a]);
  std::vector<int> rcs(T, GAPIT_OK);
  const unsigned hw = std::max(1u, std::thread::hardware_concurrency());
  const int nthreads = static_cast<int>(std::min<unsigned>(hw, static_cast<unsigned>(T)));
  std::atomic<int> next{0};
  auto work = [&] {
    for (int t = next.fetch_add(1); t < T; t = next.fetch_add(1)) {
      RemlFnL f{lambda, VtY + size_t(t) * n, VtX, n, q, ld};
      rcs[t] = reml_search(f, ngrids, llim, ulim, esp, &out[t]);
    }
  };
  std::vector<std::thread> pool;
  for (int i = 1; i < nthreads; ++i) pool.emplace_back(work);
  work();
  for (auto& th : pool) th.join();
  int first = GAPIT_OK;
  for (int t = 0; t < T; ++t) {
    if (rc_out) rc_out[t] = rcs[t];
    if (first == GAPIT_OK && rcs[t] != GAPIT_OK) first = rcs[t];
  }
  return first;
}
