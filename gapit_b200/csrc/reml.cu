// REML variance components, Z = NULL branch of GAPIT.emma.REMLE (GAPIT.emma.REMLE.R:21-54, 96-110)
// with emma.delta.REML.{LL,dLL}.wo.Z (emma.txt:145-164).  Host code: O((n-q) * 101) after the
// eigendecomposition.  The root finder restates the Brent iteration R's uniroot() runs (R_zeroin2,
// netlib zeroin.c: tol = DBL_EPSILON^0.25, maxiter = 1000, end-point values pre-evaluated).
#include <cfloat>
#include <cmath>
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
};

static double zeroin(const RemlFn& f, double ax, double bx, double fa, double fb, double tol, int maxit, bool* ok) {
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

}  // namespace gapit

using namespace gapit;

extern "C" int gapit_reml(const double* lambda, const double* etas, int64_t nq, int ngrids, double llim, double ulim,
                          double esp, gapit_reml_out* out) {
  if (!lambda || !etas || !out || nq < 1 || ngrids < 2) return fail(GAPIT_ERR_ARG, "gapit_reml: bad argument");
  std::vector<double> etasq(nq);
  for (int64_t i = 0; i < nq; ++i) etasq[i] = etas[i] * etas[i];
  RemlFn f{lambda, etasq.data(), nq};
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
  double s = 0.0;
  for (int64_t i = 0; i < nq; ++i) s += etasq[i] / (lambda[i] + maxdelta);  // :103
  out->REML = optLL[best];
  out->delta = maxdelta;
  out->vg = s / static_cast<double>(nq);
  out->ve = out->vg * maxdelta;  // :108
  return GAPIT_OK;
}
