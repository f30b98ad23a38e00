/* Minimal declarations of the R C API used by r/gapitcuda_R.c, so that the shim can be syntax- and type-checked
 * against include/gapitcuda.h on a machine without R (tests/test_host.py).  Not R's header; never linked. */
typedef struct SEXPREC* SEXP;
typedef ptrdiff_t R_xlen_t;
typedef unsigned char Rbyte;
extern SEXP R_DimSymbol, R_NamesSymbol, R_NilValue;
#define RAWSXP 24
#define REALSXP 14
#define INTSXP 13
#define VECSXP 19
#define STRSXP 16
#define EXTPTRSXP 22
typedef enum { FALSE = 0, TRUE } Rboolean;
typedef void (*R_CFinalizer_t)(SEXP);
#define NA_INTEGER (-2147483647-1)
int TYPEOF(SEXP); double* REAL(SEXP); int* INTEGER(SEXP); Rbyte* RAW(SEXP);
SEXP Rf_getAttrib(SEXP, SEXP); SEXP Rf_setAttrib(SEXP, SEXP, SEXP); SEXP Rf_allocMatrix(int, int, int); SEXP Rf_allocVector(int, R_xlen_t);
SEXP PROTECT(SEXP); void UNPROTECT(int); void Rf_error(const char*, ...); int Rf_asInteger(SEXP); double Rf_asReal(SEXP); int Rf_asLogical(SEXP);
R_xlen_t XLENGTH(SEXP); SEXP SET_VECTOR_ELT(SEXP, R_xlen_t, SEXP); SEXP VECTOR_ELT(SEXP, R_xlen_t); void SET_STRING_ELT(SEXP, R_xlen_t, SEXP); SEXP Rf_mkChar(const char*); SEXP Rf_install(const char*); SEXP Rf_ScalarInteger(int);
char* R_alloc(size_t, int);
SEXP R_MakeExternalPtr(void*, SEXP, SEXP); void* R_ExternalPtrAddr(SEXP); void R_ClearExternalPtr(SEXP); void R_SetExternalPtrAddr(SEXP, void*);
void R_RegisterCFinalizerEx(SEXP, R_CFinalizer_t, Rboolean);
