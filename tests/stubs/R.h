/* Minimal stand-in for R's C API, only so that tests can syntax-check rshim/xgc_r.c in an image without R.
 * Declarations follow R's public headers (Rinternals.h); nothing here is linked or executed. */
#ifndef XGC_TEST_R_STUB_H
#define XGC_TEST_R_STUB_H
#include <stddef.h>
typedef struct SEXPREC* SEXP;
typedef ptrdiff_t R_xlen_t;
typedef enum { FALSE = 0, TRUE } Rboolean;
#define INTSXP 13
#define REALSXP 14
#define STRSXP 16
#define VECSXP 19
#define RAWSXP 24
extern SEXP R_NilValue;
int Rf_asInteger(SEXP); double Rf_asReal(SEXP); Rboolean Rf_isNull(SEXP);
double* REAL(SEXP); int* INTEGER(SEXP); unsigned char* RAW(SEXP); int LENGTH(SEXP); R_xlen_t XLENGTH(SEXP);
SEXP STRING_ELT(SEXP, R_xlen_t); SEXP VECTOR_ELT(SEXP, R_xlen_t); const char* CHAR(SEXP);
SEXP Rf_allocVector(unsigned, R_xlen_t); SEXP Rf_allocMatrix(unsigned, int, int);
SEXP Rf_protect(SEXP); void Rf_unprotect(int);
#define PROTECT(s) Rf_protect(s)
#define UNPROTECT(n) Rf_unprotect(n)
void Rf_error(const char*, ...); char* R_alloc(size_t, int);
void* R_ExternalPtrAddr(SEXP); SEXP R_MakeExternalPtr(void*, SEXP, SEXP); void R_ClearExternalPtr(SEXP);
typedef void (*R_CFinalizer_t)(SEXP);
void R_RegisterCFinalizerEx(SEXP, R_CFinalizer_t, Rboolean);
SEXP Rf_mkString(const char*); SEXP Rf_install(const char*); SEXP Rf_setAttrib(SEXP, SEXP, SEXP); SEXP Rf_ScalarInteger(int);
SEXP Rf_coerceVector(SEXP, unsigned); int Rf_nrows(SEXP); int Rf_ncols(SEXP); Rboolean Rf_isInteger(SEXP); Rboolean Rf_isReal(SEXP);
SEXP Rf_mkChar(const char*); void SET_STRING_ELT(SEXP, R_xlen_t, SEXP); void SET_VECTOR_ELT(SEXP, R_xlen_t, SEXP);
extern SEXP R_DimSymbol; extern SEXP R_NamesSymbol;
#endif
