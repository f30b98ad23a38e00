/* .Call shim between R and libgapitcuda.so  (the only file that includes R headers).
 *
 * NOT BUILDABLE IN THE DEVELOPMENT IMAGE: R is not installed there (no Rinternals.h).  It is kept small
 * and mirrors, call for call, the ctypes binding in gapit_b200/_lib.py + gapit_b200/api.py, which is what
 * the test-suite drives.  Build on a machine with R:
 *     R CMD SHLIB gapitcuda_R.c -I../include -L../gapit_b200 -lgapitcuda -o gapitcuda_R.so
 *
 * Conventions: arguments are read-only SEXPs owned by R; results are freshly
 * allocated, PROTECTed R objects; Rf_error() is raised only after every CUDA resource of the call has
 * been released (the C ABI returns status codes, it never longjmps); everything runs on R's main thread.
 */
#include <R.h>
#include <Rinternals.h>
#include <string.h>

#include "gapitcuda.h"

/* Every visible GPU of the box (at most 8) behind one multi-context; options(gapit.gpus = k) is read by the R
 * wrappers and passed to gapit_R_init before the first call to use fewer.  ctx() is GPU 0's one-GPU context. */
static gapit_mctx* g_mctx = NULL;

static gapit_mctx* mctx(void) {
  if (!g_mctx && gapit_mctx_create(0, &g_mctx) != GAPIT_OK) Rf_error("gapitcuda: %s", gapit_last_error());
  return g_mctx;
}

static gapit_ctx* ctx(void) { return gapit_mctx_ctx(mctx(), 0); }

/* .Call("gapit_R_init", ngpus): (re)open the library on `ngpus` devices (0 = all); returns the count in use */
SEXP gapit_R_init(SEXP ngpus) {
  if (g_mctx) {
    gapit_mctx_destroy(g_mctx);
    g_mctx = NULL;
  }
  if (gapit_mctx_create(Rf_asInteger(ngpus), &g_mctx) != GAPIT_OK) Rf_error("gapitcuda: %s", gapit_last_error());
  return Rf_ScalarInteger(gapit_mctx_ngpus(g_mctx));
}

static gapit_dtype sexp_dtype(SEXP x) {
  if (TYPEOF(x) == REALSXP) return GAPIT_F64;
  if (TYPEOF(x) == INTSXP) return GAPIT_I32;
  Rf_error("gapitcuda: genotype matrix must be numeric or integer");
  return GAPIT_F64;
}

static const void* sexp_data(SEXP x) { return TYPEOF(x) == REALSXP ? (const void*)REAL(x) : (const void*)INTEGER(x); }

/* Genotype argument of the .Call targets: either an R matrix (n x m numeric / integer), or the raw bytes of a PLINK
 * .bed file (readBin(path, "raw", file.size(path))) carrying attributes "n" (rows of the .fam) and "count" (1 = copies
 * of A1, 2 = copies of A2) -- see r/GAPIT.Bed.R.
 * An R matrix holds 8 (numeric) or 4 (integer) bytes per genotype call: it is packed to 2 bits per call on the host's
 * threads (gapit_pack2_host) so that 1/32 (1/16) of the bytes cross PCIe; R_alloc memory is freed by R at the end of
 * the .Call, also on error.  A .bed is used in place.  NA / missing becomes code 3 and meets the impute rule on the
 * device, as in GAPIT.EMMAxP3D.R:20-24.  May Rf_error: call before any device handle exists. */
typedef struct {
  const uint8_t* rows;
  int64_t n, m, nb;
  gapit_dtype dtype;
} geno_src;

static geno_src geno_source(SEXP x) {
  geno_src g;
  if (TYPEOF(x) == RAWSXP) {
    SEXP an = Rf_getAttrib(x, Rf_install("n")), ac = Rf_getAttrib(x, Rf_install("count"));
    if (an == R_NilValue) Rf_error("gapitcuda: a raw .bed vector needs attribute \"n\" (samples in the .fam)");
    int64_t off = 0;
    g.n = (int64_t)Rf_asInteger(an);
    if (gapit_bed_header(RAW(x), XLENGTH(x), g.n, &g.m, &off) != GAPIT_OK) Rf_error("gapitcuda: %s", gapit_last_error());
    g.rows = (const uint8_t*)RAW(x) + off;
    g.nb = (g.n + 3) / 4;
    g.dtype = (ac != R_NilValue && Rf_asInteger(ac) == 2) ? GAPIT_BED_A2 : GAPIT_BED_A1;
    return g;
  }
  SEXP dim = Rf_getAttrib(x, R_DimSymbol);
  g.n = INTEGER(dim)[0];
  g.m = INTEGER(dim)[1];
  g.nb = (g.n + 3) / 4;
  g.dtype = GAPIT_B2;
  uint8_t* packed = (uint8_t*)R_alloc((size_t)g.m * (size_t)g.nb, 1);
  if (gapit_pack2_host(sexp_data(x), sexp_dtype(x), g.n, g.m, g.n, packed, 0) != GAPIT_OK) Rf_error("gapitcuda: %s", gapit_last_error());
  g.rows = packed;
  return g;
}

/* GAPIT.kinship.VanRaden(snps): n x m matrix -> n x n double matrix (GAPIT.kinship.VanRaden.R:1-37).
 * Markers are split across every GPU of the multi-context; the partial cross-products are exchanged inside the kernel. */
SEXP gapit_R_vanraden(SEXP snps) {
  /* everything that can longjmp (context creation, R allocations) happens before a device handle exists */
  gapit_mctx* mc = mctx();
  const geno_src src = geno_source(snps);
  const int64_t n = src.n;
  SEXP K = PROTECT(Rf_allocMatrix(REALSXP, (int)n, (int)n));
  gapit_mgeno* g = NULL;
  /* the reference does not impute here: NA propagates to a NaN matrix, so NA is an error for us */
  int rc = gapit_mgeno_pack(mc, src.rows, src.dtype, n, src.m, src.nb, GAPIT_IMPUTE_NONE, &g);
  if (rc == GAPIT_OK) rc = gapit_multi_vanraden(mc, g, REAL(K), NULL);
  gapit_mgeno_free(g);
  UNPROTECT(1);
  if (rc != GAPIT_OK) Rf_error("gapitcuda: %s", gapit_last_error());
  return K;
}

/* eigen(K, symmetric=TRUE) in R's order: list(values, vectors)   (emma.txt:58-61).  On a multi-GPU box every GPU of
 * the multi-context solves for its own slice of the spectrum (gapit_multi_eigh: cusolverDn syevdx by index range, ~10 %
 * faster than one GPU at n = 20,000; cusolverMgSyevd, gapit_eigh_mg, measured 2-3x SLOWER and is not used); one GPU
 * uses cusolverDn syevd. */
SEXP gapit_R_eigen_L(SEXP K) {
  const int64_t n = INTEGER(Rf_getAttrib(K, R_DimSymbol))[0];
  gapit_mctx* mc = mctx();
  SEXP values = PROTECT(Rf_allocVector(REALSXP, n));
  SEXP vectors = PROTECT(Rf_allocMatrix(REALSXP, (int)n, (int)n));
  int rc = gapit_multi_eigh(mc, REAL(K), n, REAL(values), REAL(vectors));
  if (rc != GAPIT_OK) {
    UNPROTECT(2);
    Rf_error("gapitcuda: %s", gapit_last_error());
  }
  SEXP out = PROTECT(Rf_allocVector(VECSXP, 2));
  SET_VECTOR_ELT(out, 0, values);
  SET_VECTOR_ELT(out, 1, vectors);
  SEXP nm = PROTECT(Rf_allocVector(STRSXP, 2));
  SET_STRING_ELT(nm, 0, Rf_mkChar("values"));
  SET_STRING_ELT(nm, 1, Rf_mkChar("vectors"));
  Rf_setAttrib(out, R_NamesSymbol, nm);
  UNPROTECT(4);
  return out;
}

/* emma.eigen.R.wo.Z(K, X): list(values (n-q), vectors n x (n-q))   (emma.txt:82-89) */
SEXP gapit_R_eigen_R(SEXP K, SEXP X) {
  SEXP dim = Rf_getAttrib(X, R_DimSymbol);
  const int64_t n = INTEGER(dim)[0], q = INTEGER(dim)[1];
  SEXP values = PROTECT(Rf_allocVector(REALSXP, n - q));
  SEXP vectors = PROTECT(Rf_allocMatrix(REALSXP, (int)n, (int)(n - q)));
  int rc = gapit_eigen_R(ctx(), REAL(K), REAL(X), n, q, REAL(values), REAL(vectors));
  if (rc != GAPIT_OK) {
    UNPROTECT(2);
    Rf_error("gapitcuda: %s", gapit_last_error());
  }
  SEXP out = PROTECT(Rf_allocVector(VECSXP, 2));
  SET_VECTOR_ELT(out, 0, values);
  SET_VECTOR_ELT(out, 1, vectors);
  SEXP nm = PROTECT(Rf_allocVector(STRSXP, 2));
  SET_STRING_ELT(nm, 0, Rf_mkChar("values"));
  SET_STRING_ELT(nm, 1, Rf_mkChar("vectors"));
  Rf_setAttrib(out, R_NamesSymbol, nm);
  UNPROTECT(4);
  return out;
}

/* REML on the spectrum: c(REML, delta, ve, vg)   (GAPIT.emma.REMLE.R:27-54, 96-108) */
SEXP gapit_R_reml(SEXP lambda, SEXP etas, SEXP ngrids, SEXP llim, SEXP ulim, SEXP esp) {
  gapit_reml_out r;
  int rc = gapit_reml(REAL(lambda), REAL(etas), XLENGTH(lambda), Rf_asInteger(ngrids), Rf_asReal(llim), Rf_asReal(ulim),
                      Rf_asReal(esp), &r);
  if (rc != GAPIT_OK) Rf_error("gapitcuda: %s", gapit_last_error());
  SEXP out = PROTECT(Rf_allocVector(REALSXP, 4));
  REAL(out)[0] = r.REML;
  REAL(out)[1] = r.delta;
  REAL(out)[2] = r.ve;
  REAL(out)[3] = r.vg;
  UNPROTECT(1);
  return out;
}

/* ---- the eigen-system of K as an R-visible handle -----------------------------------------------------------------
 * One GAPIT.EMMAxP3D call needs eig.L four times: for the REML projections, for BLUP/PEV, for the scan, and (in the
 * reference) as an R object in between.  At n = 20,000 the vectors are 3.2 GB: returning them to R and uploading them
 * again for every step costs more than the scan.  gapit_R_eigsys decomposes K on the GPUs (spectrum sliced over them)
 * and leaves the sliced eigen-system resident on every GPU behind an external pointer; R sees only the eigenvalues. */
static void eigsys_finalizer(SEXP p) {
  gapit_meigsys* me = (gapit_meigsys*)R_ExternalPtrAddr(p);
  if (me) {
    gapit_meigsys_free(me);
    R_ClearExternalPtr(p);
  }
}

static gapit_meigsys* eigsys_handle(SEXP h) {
  gapit_meigsys* me = TYPEOF(h) == EXTPTRSXP ? (gapit_meigsys*)R_ExternalPtrAddr(h) : NULL;
  if (!me) Rf_error("gapitcuda: not a live eigen-system handle (gapit_R_eigsys)");
  return me;
}

/* .Call("gapit_R_eigsys", K): list(values = eigen(K)$values, handle = <external pointer>) */
SEXP gapit_R_eigsys(SEXP K) {
  const int64_t n = INTEGER(Rf_getAttrib(K, R_DimSymbol))[0];
  gapit_mctx* mc = mctx();
  SEXP values = PROTECT(Rf_allocVector(REALSXP, n));
  SEXP out = PROTECT(Rf_allocVector(VECSXP, 2));
  SEXP nm = PROTECT(Rf_allocVector(STRSXP, 2));
  SET_STRING_ELT(nm, 0, Rf_mkChar("values"));
  SET_STRING_ELT(nm, 1, Rf_mkChar("handle"));
  Rf_setAttrib(out, R_NamesSymbol, nm);
  SEXP h = PROTECT(R_MakeExternalPtr(NULL, R_NilValue, R_NilValue));
  R_RegisterCFinalizerEx(h, eigsys_finalizer, TRUE);
  gapit_meigsys* me = NULL;   /* from here on nothing can longjmp until the handle is owned by the external pointer */
  if (gapit_meigsys_from_kinship(mc, REAL(K), n, REAL(values), &me) != GAPIT_OK) {
    UNPROTECT(4);
    Rf_error("gapitcuda: %s", gapit_last_error());
  }
  R_SetExternalPtrAddr(h, me);
  SET_VECTOR_ELT(out, 0, values);
  SET_VECTOR_ELT(out, 1, h);
  UNPROTECT(4);
  return out;
}

/* .Call("gapit_R_eigsys_free", handle): release the device copies now instead of at garbage collection */
SEXP gapit_R_eigsys_free(SEXP h) {
  if (TYPEOF(h) == EXTPTRSXP) eigsys_finalizer(h);
  return R_NilValue;
}

/* crossprod(eig.L$vectors, R) for an n x nc matrix R, on GPU 0 */
SEXP gapit_R_eigsys_project(SEXP h, SEXP R) {
  gapit_meigsys* me = eigsys_handle(h);
  SEXP dim = Rf_getAttrib(R, R_DimSymbol);
  const int n = INTEGER(dim)[0], nc = INTEGER(dim)[1];
  SEXP out = PROTECT(Rf_allocMatrix(REALSXP, n, nc));
  const int rc = gapit_eigsys_project(ctx(), gapit_meigsys_part(me, 0), REAL(R), nc, REAL(out));
  UNPROTECT(1);
  if (rc != GAPIT_OK) Rf_error("gapitcuda: %s", gapit_last_error());
  return out;
}

/* The marker loop of GAPIT.EMMAxP3D (GAPIT.EMMAxP3D.R:397-979) for T traits that share X0 and the eigen-system: the
 * reference loops the rows of ys (:81) and repeats the rotation for each; here one rotation serves them all.
 * xs n x m (numeric/integer) or the raw bytes of a .bed (see geno_source); the eigen-system is either a resident handle
 * (me) or eig.L on the host (vectors/values; uploaded for this call); neither for the GLM branch.  X0 n x q0, Y n x T
 * (a vector is one trait), delta / vg length T.  impute: -1 none, 0 Minor, 1 Middle, 2 Major.
 * Returns list(effect, tvalue, se, pvalue, rsquare, loglm) of m x T matrices, maf (length m) and base, a
 * (7 + q0) x T matrix with rows logL0, logLM_base, rsquare_base, ves, reml, df, singular_x0, beta0[1..q0]. */
static SEXP scan_core(SEXP xs, gapit_meigsys* me, SEXP vectors, SEXP values, SEXP X0, SEXP Y, SEXP delta, SEXP vg, int is_glm,
                      SEXP impute) {
  const int q0 = INTEGER(Rf_getAttrib(X0, R_DimSymbol))[1];
  /* everything that can longjmp happens before a device handle exists */
  gapit_mctx* mc = mctx();
  const geno_src src = geno_source(xs);
  const int64_t n = src.n, m = src.m, nb = src.nb;
  const uint8_t* packed = src.rows;
  const int T = (int)(XLENGTH(Y) / n);
  if ((int64_t)T * n != XLENGTH(Y) || T < 1) Rf_error("gapitcuda: Y must hold n values per trait");
  if (!is_glm && (XLENGTH(delta) != T || XLENGTH(vg) != T)) Rf_error("gapitcuda: delta and vg need one value per trait");
  SEXP res = PROTECT(Rf_allocVector(VECSXP, 8));
  for (int a = 0; a < 6; ++a) SET_VECTOR_ELT(res, a, Rf_allocMatrix(REALSXP, (int)m, T));
  SET_VECTOR_ELT(res, 6, Rf_allocVector(REALSXP, m));
  SET_VECTOR_ELT(res, 7, Rf_allocMatrix(REALSXP, 7 + q0, T));
  SEXP nm = PROTECT(Rf_allocVector(STRSXP, 8));
  const char* names[8] = {"effect", "tvalue", "se", "pvalue", "rsquare", "loglm", "maf", "base"};
  for (int a = 0; a < 8; ++a) SET_STRING_ELT(nm, a, Rf_mkChar(names[a]));
  Rf_setAttrib(res, R_NamesSymbol, nm);
  gapit_scan_base* base = (gapit_scan_base*)R_alloc((size_t)T, sizeof(gapit_scan_base));
  gapit_scan_out out = {REAL(VECTOR_ELT(res, 0)), REAL(VECTOR_ELT(res, 1)), REAL(VECTOR_ELT(res, 2)), REAL(VECTOR_ELT(res, 3)),
                        REAL(VECTOR_ELT(res, 4)), REAL(VECTOR_ELT(res, 5)), REAL(VECTOR_ELT(res, 6))};
  const double* d = is_glm ? NULL : REAL(delta);
  const double* v = is_glm ? NULL : REAL(vg);
  const gapit_impute imp = (gapit_impute)Rf_asInteger(impute);
  gapit_meigsys* own = NULL;   /* an eigen-system uploaded for this call only */
  int rc = GAPIT_OK;
  if (!is_glm && !me) {
    rc = gapit_meigsys_upload(mc, REAL(vectors), REAL(values), n, &own);
    me = own;
  }
  if (rc == GAPIT_OK && gapit_mctx_ngpus(mc) == 1) {
    /* one GPU: stream the packed rows (upload overlapped with the scan) */
    rc = gapit_p3d_scan_host(ctx(), packed, src.dtype, n, m, nb, imp, is_glm ? NULL : gapit_meigsys_part(me, 0), REAL(X0), q0,
                             REAL(Y), T, d, v, is_glm, &out, base);
  } else if (rc == GAPIT_OK) {
    /* several GPUs: one marker block per GPU, results in marker order */
    gapit_mgeno* g = NULL;
    rc = gapit_mgeno_pack(mc, packed, src.dtype, n, m, nb, imp, &g);
    if (rc == GAPIT_OK) rc = gapit_multi_p3d_scan_es(mc, g, is_glm ? NULL : me, REAL(X0), q0, REAL(Y), T, d, v, is_glm, &out, base);
    gapit_mgeno_free(g);
  }
  gapit_meigsys_free(own);
  if (rc != GAPIT_OK) {
    UNPROTECT(2);
    Rf_error("gapitcuda: %s", gapit_last_error());
  }
  double* b = REAL(VECTOR_ELT(res, 7));
  for (int t = 0; t < T; ++t, b += 7 + q0) {
    b[0] = base[t].logL0;
    b[1] = base[t].logLM_base;
    b[2] = base[t].rsquare_base;
    b[3] = base[t].ves;
    b[4] = base[t].reml;
    b[5] = base[t].df;
    b[6] = (double)base[t].singular_x0;
    memcpy(b + 7, base[t].beta0, sizeof(double) * q0);
  }
  UNPROTECT(2);
  return res;
}

/* eig.L as R objects (the reference's eig.R route of r/GAPIT.EMMAxP3D.R, and the GLM branch with glm = TRUE) */
SEXP gapit_R_p3d_scan(SEXP xs, SEXP vectors, SEXP values, SEXP X0, SEXP Y, SEXP delta, SEXP vg, SEXP glm, SEXP impute) {
  return scan_core(xs, NULL, vectors, values, X0, Y, delta, vg, Rf_asLogical(glm), impute);
}

/* the same against the resident handle of gapit_R_eigsys */
SEXP gapit_R_eigsys_scan(SEXP xs, SEXP h, SEXP X0, SEXP Y, SEXP delta, SEXP vg, SEXP impute) {
  return scan_core(xs, eigsys_handle(h), R_NilValue, R_NilValue, X0, Y, delta, vg, 0, impute);
}

/* i = 0 block of GAPIT.EMMAxP3D (GAPIT.EMMAxP3D.R:779-839) against the resident handle: list(BLUP, PEV, beta) */
SEXP gapit_R_eigsys_blup_pev(SEXP h, SEXP X0, SEXP y, SEXP delta, SEXP vg) {
  gapit_meigsys* me = eigsys_handle(h);
  SEXP dim = Rf_getAttrib(X0, R_DimSymbol);
  const int64_t n = INTEGER(dim)[0];
  const int q0 = INTEGER(dim)[1];
  gapit_ctx* c = ctx();
  SEXP blup = PROTECT(Rf_allocMatrix(REALSXP, (int)n, 1));
  SEXP pev = PROTECT(Rf_allocMatrix(REALSXP, (int)n, 1));
  SEXP beta = PROTECT(Rf_allocVector(REALSXP, q0));
  SEXP out = PROTECT(Rf_allocVector(VECSXP, 3));
  const int rc = gapit_blup_pev(c, gapit_meigsys_part(me, 0), REAL(X0), q0, REAL(y), Rf_asReal(delta), Rf_asReal(vg), REAL(blup),
                                REAL(pev), REAL(beta));
  if (rc != GAPIT_OK) {
    UNPROTECT(4);
    Rf_error("gapitcuda: %s", gapit_last_error());
  }
  SET_VECTOR_ELT(out, 0, blup);
  SET_VECTOR_ELT(out, 1, pev);
  SET_VECTOR_ELT(out, 2, beta);
  UNPROTECT(4);
  return out;
}

/* i = 0 block of GAPIT.EMMAxP3D (GAPIT.EMMAxP3D.R:779-839): list(BLUP, PEV, beta) */
SEXP gapit_R_blup_pev(SEXP vectors, SEXP values, SEXP X0, SEXP y, SEXP delta, SEXP vg) {
  SEXP dim = Rf_getAttrib(X0, R_DimSymbol);
  const int64_t n = INTEGER(dim)[0];
  const int q0 = INTEGER(dim)[1];
  gapit_ctx* c = ctx();
  SEXP blup = PROTECT(Rf_allocMatrix(REALSXP, (int)n, 1));
  SEXP pev = PROTECT(Rf_allocMatrix(REALSXP, (int)n, 1));
  SEXP beta = PROTECT(Rf_allocVector(REALSXP, q0));
  SEXP out = PROTECT(Rf_allocVector(VECSXP, 3));
  const double dl = Rf_asReal(delta), vgv = Rf_asReal(vg);
  gapit_eigsys* e = NULL;   /* from here on nothing can longjmp until the handle is released */
  int rc = gapit_eigsys_upload(c, REAL(vectors), REAL(values), n, &e);
  if (rc == GAPIT_OK) rc = gapit_blup_pev(c, e, REAL(X0), q0, REAL(y), dl, vgv, REAL(blup), REAL(pev), REAL(beta));
  gapit_eigsys_free(e);
  if (rc != GAPIT_OK) {
    UNPROTECT(4);
    Rf_error("gapitcuda: %s", gapit_last_error());
  }
  SET_VECTOR_ELT(out, 0, blup);
  SET_VECTOR_ELT(out, 1, pev);
  SET_VECTOR_ELT(out, 2, beta);
  UNPROTECT(4);
  return out;
}

/* REML from eig.L alone: c(REML, delta, ve, vg).  values = eig.L$values, Vty = crossprod(eig.L$vectors, y),
 * VtX = crossprod(eig.L$vectors, X0)   (same likelihood as gapit_R_reml on eig.R; GAPIT.emma.REMLE.R:27-54, 96-108) */
SEXP gapit_R_reml_eigL(SEXP values, SEXP Vty, SEXP VtX, SEXP ngrids, SEXP llim, SEXP ulim, SEXP esp) {
  gapit_reml_out r;
  const int64_t n = XLENGTH(values);
  const int q = (int)(XLENGTH(VtX) / n);
  int rc = gapit_reml_eigL(REAL(values), REAL(Vty), REAL(VtX), n, q, Rf_asInteger(ngrids), Rf_asReal(llim), Rf_asReal(ulim),
                           Rf_asReal(esp), &r);
  if (rc != GAPIT_OK) Rf_error("gapitcuda: %s", gapit_last_error());
  SEXP out = PROTECT(Rf_allocVector(REALSXP, 4));
  REAL(out)[0] = r.REML;
  REAL(out)[1] = r.delta;
  REAL(out)[2] = r.ve;
  REAL(out)[3] = r.vg;
  UNPROTECT(1);
  return out;
}

/* REML of T traits from eig.L alone (gapit_R_reml_eigL for every column of VtY, on the host's threads): 4 x T matrix with
 * rows REML, delta, ve, vg. */
SEXP gapit_R_reml_eigL_multi(SEXP values, SEXP VtY, SEXP VtX, SEXP ngrids, SEXP llim, SEXP ulim, SEXP esp) {
  const int64_t n = XLENGTH(values);
  const int q = (int)(XLENGTH(VtX) / n), T = (int)(XLENGTH(VtY) / n);
  gapit_reml_out* r = (gapit_reml_out*)R_alloc((size_t)T, sizeof(gapit_reml_out));
  int rc = gapit_reml_eigL_multi(REAL(values), REAL(VtY), REAL(VtX), n, q, T, Rf_asInteger(ngrids), Rf_asReal(llim),
                                 Rf_asReal(ulim), Rf_asReal(esp), r, NULL);
  if (rc != GAPIT_OK) Rf_error("gapitcuda: %s", gapit_last_error());
  SEXP out = PROTECT(Rf_allocMatrix(REALSXP, 4, T));
  for (int t = 0; t < T; ++t) {
    REAL(out)[4 * t + 0] = r[t].REML;
    REAL(out)[4 * t + 1] = r[t].delta;
    REAL(out)[4 * t + 2] = r[t].ve;
    REAL(out)[4 * t + 3] = r[t].vg;
  }
  UNPROTECT(1);
  return out;
}

/* GAPIT.kinship.ZHANG(snps): n x m matrix -> n x n double matrix (GAPIT.kinship.ZHANG.R:1-66) */
SEXP gapit_R_zhang(SEXP snps) {
  SEXP dim = Rf_getAttrib(snps, R_DimSymbol);
  const int64_t n = INTEGER(dim)[0], m = INTEGER(dim)[1];
  gapit_ctx* c = ctx();
  const gapit_dtype dt = sexp_dtype(snps);
  SEXP K = PROTECT(Rf_allocMatrix(REALSXP, (int)n, (int)n));
  gapit_geno* g = NULL;   /* nothing below can longjmp while the handle is held */
  int rc = gapit_geno_pack(c, sexp_data(snps), dt, n, m, n, GAPIT_IMPUTE_NONE, &g);   /* :27 "Missing data is not allowed ..." */
  if (rc == GAPIT_OK) rc = gapit_zhang(c, g, REAL(K));
  gapit_geno_free(g);
  UNPROTECT(1);
  if (rc != GAPIT_OK) Rf_error("gapitcuda: %s", gapit_last_error());
  return K;
}

/* prcomp(X) of GAPIT.PCA.R:10: list(x = n x ncomp scores, sdev2 = min(n, m) variances) */
SEXP gapit_R_pca(SEXP X, SEXP ncomp) {
  SEXP dim = Rf_getAttrib(X, R_DimSymbol);
  const int64_t n = INTEGER(dim)[0], m = INTEGER(dim)[1], k = Rf_asInteger(ncomp), rank = n < m ? n : m;
  gapit_ctx* c = ctx();
  const gapit_dtype dt = sexp_dtype(X);
  SEXP x = PROTECT(Rf_allocMatrix(REALSXP, (int)n, (int)k));
  SEXP ev = PROTECT(Rf_allocVector(REALSXP, rank));
  SEXP out = PROTECT(Rf_allocVector(VECSXP, 2));
  gapit_geno* g = NULL;   /* nothing below can longjmp while the handle is held */
  int rc = gapit_geno_pack(c, sexp_data(X), dt, n, m, n, GAPIT_IMPUTE_NONE, &g);
  if (rc == GAPIT_OK) rc = gapit_pca(c, g, k, REAL(x), REAL(ev));
  gapit_geno_free(g);
  if (rc != GAPIT_OK) {
    UNPROTECT(3);
    Rf_error("gapitcuda: %s", gapit_last_error());
  }
  SET_VECTOR_ELT(out, 0, x);
  SET_VECTOR_ELT(out, 1, ev);
  UNPROTECT(3);
  return out;
}

/* The numericalization loop of GAPIT.HapMap (GAPIT.HapMap.R:34-38 over GAPIT.Numericalization.R) on a raw vector:
 * `text` = the bytes of the HapMap file (readBin(path, "raw", file.size(path))), heading TRUE/FALSE.
 * effect 0 Add / 1 Dom / 2 Left / 3 Right, impute -1 None / 0 Minor / 1 Middle / 2 Major.
 * Returns the n x m integer matrix GD (NA_integer_ where a call stays missing). */
SEXP gapit_R_hapmap(SEXP text, SEXP heading, SEXP effect, SEXP impute, SEXP maz) {
  const char* t = (const char*)RAW(text);
  const int64_t bytes = XLENGTH(text);
  int64_t n = 0, m = 0, hoff = 0;
  int bit = 0;
  if (gapit_hapmap_index(t, bytes, Rf_asLogical(heading), &n, &m, &bit, NULL, 0, &hoff) != GAPIT_OK)
    Rf_error("gapitcuda: %s", gapit_last_error());
  int64_t* off = (int64_t*)R_alloc((size_t)m, sizeof(int64_t));
  gapit_hapmap_index(t, bytes, Rf_asLogical(heading), &n, &m, &bit, off, m, &hoff);
  int8_t* gd = (int8_t*)R_alloc((size_t)n * (size_t)m, 1);
  int rc = gapit_hapmap_numericalize(ctx(), t, bytes, off, bit, bit + 1, n, m, Rf_asInteger(effect),
                                     (gapit_impute)Rf_asInteger(impute), Rf_asLogical(maz), gd, NULL);
  if (rc != GAPIT_OK) Rf_error("gapitcuda: %s", gapit_last_error());
  SEXP GD = PROTECT(Rf_allocMatrix(INTSXP, (int)n, (int)m));   /* marker-major bytes == column-major n x m */
  int* o = INTEGER(GD);
  for (int64_t i = 0; i < n * m; ++i) o[i] = gd[i] < 0 ? NA_INTEGER : gd[i];
  UNPROTECT(1);
  return GD;
}
