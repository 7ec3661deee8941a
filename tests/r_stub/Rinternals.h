/* Minimal stand-in for R's <Rinternals.h> (compile check only, see R.h in this directory). */
#ifndef R_STUB_RINTERNALS_H
#define R_STUB_RINTERNALS_H
#include <stddef.h>
typedef struct SEXPREC* SEXP;
typedef unsigned char Rbyte;
typedef int Rboolean;
#ifndef TRUE
#define TRUE 1
#define FALSE 0
#endif
#define INTSXP 13
#define REALSXP 14
#define STRSXP 16
#define VECSXP 19
#define RAWSXP 24
typedef ptrdiff_t R_xlen_t;
int R_IsNaN(double x);
int R_IsNA(double x);
#define ISNAN(x) ((x) != (x))
extern SEXP R_NilValue, R_NamesSymbol;
extern double R_NaReal;
#define NA_REAL R_NaReal
int TYPEOF(SEXP x);
int* INTEGER(SEXP x);
double* REAL(SEXP x);
Rbyte* RAW(SEXP x);
const char* CHAR(SEXP x);
SEXP STRING_ELT(SEXP x, ptrdiff_t i);
void SET_STRING_ELT(SEXP x, ptrdiff_t i, SEXP v);
SEXP VECTOR_ELT(SEXP x, ptrdiff_t i);
int Rf_length(SEXP x);
R_xlen_t Rf_xlength(SEXP x);
int Rf_nrows(SEXP x);
int Rf_ncols(SEXP x);
Rboolean Rf_isNull(SEXP x);
double Rf_asReal(SEXP x);
SEXP Rf_getAttrib(SEXP x, SEXP name);
SEXP Rf_setAttrib(SEXP x, SEXP name, SEXP val);
SEXP Rf_install(const char* name);
SEXP Rf_mkChar(const char* s);
SEXP Rf_allocVector(unsigned int type, ptrdiff_t n);
SEXP Rf_allocMatrix(unsigned int type, int nrow, int ncol);
SEXP Rf_alloc3DArray(unsigned int type, int nrow, int ncol, int nface);
SEXP Rf_ScalarInteger(int x);
SEXP Rf_protect(SEXP x);
void Rf_unprotect(int n);
#define PROTECT(x) Rf_protect(x)
#define UNPROTECT(n) Rf_unprotect(n)
void Rf_error(const char* fmt, ...) __attribute__((noreturn));
void R_CheckUserInterrupt(void);
Rboolean R_ToplevelExec(void (*fun)(void* data), void* data);
#endif
