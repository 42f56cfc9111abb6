/* Minimal stand-in for the part of R's C API that r/simples_rcall.c uses (TEST INFRASTRUCTURE).
 *
 * R and its headers are not in the build image, so the .Call shim of the drop-in boundary could not even be compiled
 * here.  These headers + rstub.c implement, on plain heap objects, exactly the entry points the shim touches
 * (allocVector / allocMatrix / REAL / INTEGER / VECTOR_ELT / mkNamed / asInteger ... / error / registerRoutines) with R's
 * documented semantics (Rf_ncols of a plain vector is 1, Rf_nrows its length; Rf_error does not return), so that
 * tests/test_r_shim.py can compile the shim UNMODIFIED, check its registration table against the R wrappers in r/, and
 * drive it end to end (.Call-style) on a GPU.  It is not R: no garbage collector (PROTECT is the identity), no attributes
 * beyond dim and names.
 */
#ifndef RSTUB_H
#define RSTUB_H
#include <stddef.h>

#ifdef __cplusplus
extern "C" {
#endif

#define NILSXP 0
#define LGLSXP 10
#define INTSXP 13
#define REALSXP 14
#define VECSXP 19
#ifndef TRUE
#define TRUE 1
#define FALSE 0
#endif

typedef ptrdiff_t R_xlen_t;
typedef struct rstub_sexp {
    int type;
    R_xlen_t len;
    int nrow, ncol;        /* ncol < 0: no dim attribute */
    void *data;            /* double[], int[] or SEXP[] */
    const char **names;    /* VECSXP made by Rf_mkNamed */
} *SEXP;

extern SEXP R_NilValue;

SEXP Rf_allocVector(int type, R_xlen_t n);
SEXP Rf_allocMatrix(int type, int nrow, int ncol);
SEXP Rf_mkNamed(int type, const char **names);
SEXP Rf_coerceVector(SEXP x, int type);
SEXP Rf_ScalarInteger(int v);
double *REAL(SEXP x);
int *INTEGER(SEXP x);
SEXP VECTOR_ELT(SEXP x, R_xlen_t i);
SEXP SET_VECTOR_ELT(SEXP x, R_xlen_t i, SEXP v);
int Rf_nrows(SEXP x);
int Rf_ncols(SEXP x);
int Rf_length(SEXP x);
int Rf_isNull(SEXP x);
int Rf_asInteger(SEXP x);
double Rf_asReal(SEXP x);
int Rf_asLogical(SEXP x);
void Rf_error(const char *fmt, ...) __attribute__((noreturn, format(printf, 1, 2)));
char *R_alloc(size_t n, int size);

#define PROTECT(x) (x)
#define UNPROTECT(n) ((void)(n))

typedef void *(*DL_FUNC)(void);
typedef struct {
    const char *name;
    DL_FUNC fun;
    int numArgs;
} R_CallMethodDef;
typedef struct rstub_dllinfo DllInfo;
int R_registerRoutines(DllInfo *dll, const void *c, const R_CallMethodDef *call, const void *f, const void *e);
int R_useDynamicSymbols(DllInfo *dll, int value);

/* ---- harness side (called from Python through ctypes) ---- */
SEXP rstub_nil(void);
SEXP rstub_real(const double *src, R_xlen_t n, int nrow, int ncol);   /* ncol < 0: plain vector */
SEXP rstub_int(const int *src, R_xlen_t n);
SEXP rstub_lgl(int v);
SEXP rstub_list(R_xlen_t n);
int rstub_type(SEXP x);
R_xlen_t rstub_len(SEXP x);
int rstub_nrow(SEXP x);
int rstub_ncol(SEXP x);
void *rstub_data(SEXP x);
const char *rstub_name(SEXP x, R_xlen_t i);
int rstub_registered(void);
const char *rstub_routine_name(int i);
int rstub_routine_nargs(int i);
int rstub_dynamic_symbols(void);
SEXP rstub_dot_call(const char *name, int nargs, SEXP *args);         /* NULL + rstub_last_error() when Rf_error fired */
const char *rstub_last_error(void);
void rstub_free_all(void);

#ifdef __cplusplus
}
#endif
#endif
