/* Heap-object implementation of the R API subset declared in rstub.h (TEST INFRASTRUCTURE, see rstub.h). */
#include "rstub.h"

#include <setjmp.h>
#include <stdarg.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

static struct rstub_sexp nil_obj = {NILSXP, 0, 0, -1, NULL, NULL};
SEXP R_NilValue = &nil_obj;

/* every allocation is remembered so that a test can release them in one go */
static void **g_allocs = NULL;
static size_t g_nalloc = 0, g_cap = 0;
static void *track(void *p) {
    if (g_nalloc == g_cap) {
        g_cap = g_cap ? 2 * g_cap : 256;
        g_allocs = (void **)realloc(g_allocs, g_cap * sizeof(void *));
    }
    g_allocs[g_nalloc++] = p;
    return p;
}
void rstub_free_all(void) {
    for (size_t i = 0; i < g_nalloc; ++i) free(g_allocs[i]);
    g_nalloc = 0;
}

static jmp_buf g_jmp;
static int g_jmp_armed = 0;
static char g_err[1024];

void Rf_error(const char *fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof g_err, fmt, ap);
    va_end(ap);
    if (g_jmp_armed) longjmp(g_jmp, 1);
    fprintf(stderr, "rstub: Rf_error outside .Call: %s\n", g_err);
    abort();
}
const char *rstub_last_error(void) { return g_err; }

static size_t elt_size(int type) {
    switch (type) {
        case REALSXP: return sizeof(double);
        case INTSXP:
        case LGLSXP: return sizeof(int);
        case VECSXP: return sizeof(SEXP);
        default: Rf_error("rstub: unsupported SEXP type %d", type);
    }
}

SEXP Rf_allocVector(int type, R_xlen_t n) {
    if (n < 0) Rf_error("negative length vectors are not allowed");
    const size_t es = elt_size(type);
    SEXP x = (SEXP)track(calloc(1, sizeof *x));
    x->type = type;
    x->len = n;
    x->nrow = (int)n;
    x->ncol = -1;
    x->data = track(calloc((size_t)(n > 0 ? n : 1), es));
    if (type == VECSXP)
        for (R_xlen_t i = 0; i < n; ++i) ((SEXP *)x->data)[i] = R_NilValue;
    return x;
}

SEXP Rf_allocMatrix(int type, int nrow, int ncol) {
    if (nrow < 0 || ncol < 0) Rf_error("negative extents to matrix");
    SEXP x = Rf_allocVector(type, (R_xlen_t)nrow * ncol);
    x->nrow = nrow;
    x->ncol = ncol;
    return x;
}

SEXP Rf_mkNamed(int type, const char **names) {
    R_xlen_t n = 0;
    while (names[n][0] != '\0') ++n;
    SEXP x = Rf_allocVector(type, n);
    x->names = (const char **)track(calloc((size_t)n + 1, sizeof(char *)));
    for (R_xlen_t i = 0; i < n; ++i) {
        char *c = (char *)track(malloc(strlen(names[i]) + 1));
        strcpy(c, names[i]);
        x->names[i] = c;
    }
    return x;
}

double *REAL(SEXP x) {
    if (x->type != REALSXP) Rf_error("REAL() can only be applied to a 'numeric', not a type-%d object", x->type);
    return (double *)x->data;
}
int *INTEGER(SEXP x) {
    if (x->type != INTSXP && x->type != LGLSXP)
        Rf_error("INTEGER() can only be applied to a 'integer', not a type-%d object", x->type);
    return (int *)x->data;
}
SEXP VECTOR_ELT(SEXP x, R_xlen_t i) {
    if (x->type != VECSXP) Rf_error("VECTOR_ELT() can only be applied to a 'list', not a type-%d object", x->type);
    if (i < 0 || i >= x->len) Rf_error("attempt to access index %ld/%ld in VECTOR_ELT", (long)i, (long)x->len);
    return ((SEXP *)x->data)[i];
}
SEXP SET_VECTOR_ELT(SEXP x, R_xlen_t i, SEXP v) {
    if (x->type != VECSXP) Rf_error("SET_VECTOR_ELT() can only be applied to a 'list', not a type-%d object", x->type);
    if (i < 0 || i >= x->len) Rf_error("attempt to set index %ld/%ld in SET_VECTOR_ELT", (long)i, (long)x->len);
    ((SEXP *)x->data)[i] = v;
    return v;
}
int Rf_isNull(SEXP x) { return x->type == NILSXP; }
int Rf_length(SEXP x) { return (int)x->len; }
int Rf_nrows(SEXP x) { return x->ncol < 0 ? (int)x->len : x->nrow; }      /* a vector counts as a one-column matrix */
int Rf_ncols(SEXP x) { return x->ncol < 0 ? 1 : x->ncol; }

static double first_as_double(SEXP x) {
    if (x->len < 1) Rf_error("argument of length 0 where a scalar is needed");
    if (x->type == REALSXP) return ((double *)x->data)[0];
    if (x->type == INTSXP || x->type == LGLSXP) return (double)((int *)x->data)[0];
    Rf_error("cannot coerce a type-%d object to a scalar", x->type);
}
int Rf_asInteger(SEXP x) { return (int)first_as_double(x); }
double Rf_asReal(SEXP x) { return first_as_double(x); }
int Rf_asLogical(SEXP x) { return first_as_double(x) != 0.0; }

SEXP Rf_coerceVector(SEXP x, int type) {
    if (x->type == type) return x;
    if (type != INTSXP || x->type != REALSXP) Rf_error("rstub: coercion %d -> %d not implemented", x->type, type);
    SEXP y = Rf_allocVector(INTSXP, x->len);
    for (R_xlen_t i = 0; i < x->len; ++i) ((int *)y->data)[i] = (int)((double *)x->data)[i];
    return y;
}
SEXP Rf_ScalarInteger(int v) {
    SEXP x = Rf_allocVector(INTSXP, 1);
    ((int *)x->data)[0] = v;
    return x;
}
char *R_alloc(size_t n, int size) { return (char *)track(calloc(n ? n : 1, (size_t)size)); }

/* ---- routine registration (R_init_<pkg>) ---- */
static const R_CallMethodDef *g_calls = NULL;
static int g_ncalls = 0, g_dynsym = -1;
int R_registerRoutines(DllInfo *dll, const void *c, const R_CallMethodDef *call, const void *f, const void *e) {
    (void)dll; (void)c; (void)f; (void)e;
    g_calls = call;
    g_ncalls = 0;
    while (call && call[g_ncalls].name) ++g_ncalls;
    return 1;
}
int R_useDynamicSymbols(DllInfo *dll, int value) {
    (void)dll;
    g_dynsym = value;
    return 1;
}
int rstub_registered(void) { return g_ncalls; }
const char *rstub_routine_name(int i) { return g_calls[i].name; }
int rstub_routine_nargs(int i) { return g_calls[i].numArgs; }
int rstub_dynamic_symbols(void) { return g_dynsym; }

/* ---- harness constructors / accessors ---- */
SEXP rstub_nil(void) { return R_NilValue; }
SEXP rstub_real(const double *src, R_xlen_t n, int nrow, int ncol) {
    SEXP x = ncol < 0 ? Rf_allocVector(REALSXP, n) : Rf_allocMatrix(REALSXP, nrow, ncol);
    if (src && n > 0) memcpy(x->data, src, sizeof(double) * (size_t)n);
    return x;
}
SEXP rstub_int(const int *src, R_xlen_t n) {
    SEXP x = Rf_allocVector(INTSXP, n);
    if (src && n > 0) memcpy(x->data, src, sizeof(int) * (size_t)n);
    return x;
}
SEXP rstub_lgl(int v) {
    SEXP x = Rf_allocVector(LGLSXP, 1);
    ((int *)x->data)[0] = v;
    return x;
}
SEXP rstub_list(R_xlen_t n) { return Rf_allocVector(VECSXP, n); }
int rstub_type(SEXP x) { return x->type; }
R_xlen_t rstub_len(SEXP x) { return x->len; }
int rstub_nrow(SEXP x) { return x->nrow; }
int rstub_ncol(SEXP x) { return x->ncol; }
void *rstub_data(SEXP x) { return x->data; }
const char *rstub_name(SEXP x, R_xlen_t i) { return x->names ? x->names[i] : NULL; }

#define A1 a[0]
#define A6 A1, a[1], a[2], a[3], a[4], a[5]
#define A7 A6, a[6]
#define A12 A7, a[7], a[8], a[9], a[10], a[11]
#define A16 A12, a[12], a[13], a[14], a[15]
#define A25 A16, a[16], a[17], a[18], a[19], a[20], a[21], a[22], a[23], a[24]
#define S1 SEXP
#define S6 S1, SEXP, SEXP, SEXP, SEXP, SEXP
#define S7 S6, SEXP
#define S12 S7, SEXP, SEXP, SEXP, SEXP, SEXP
#define S16 S12, SEXP, SEXP, SEXP, SEXP
#define S25 S16, SEXP, SEXP, SEXP, SEXP, SEXP, SEXP, SEXP, SEXP, SEXP

static SEXP dispatch(DL_FUNC f, int n, SEXP *a) {
    switch (n) {      /* the argument counts r/simples_rcall.c registers */
        case 1: return ((SEXP(*)(S1))f)(A1);
        case 6: return ((SEXP(*)(S6))f)(A6);
        case 7: return ((SEXP(*)(S7))f)(A7);
        case 12: return ((SEXP(*)(S12))f)(A12);
        case 16: return ((SEXP(*)(S16))f)(A16);
        case 25: return ((SEXP(*)(S25))f)(A25);
        default: Rf_error("rstub: no dispatcher for %d arguments", n);
    }
}

/* .Call(name, ...): looks the routine up in the registered table (as R does for a package with registered routines),
 * checks the argument count against numArgs, and turns Rf_error into a NULL return. */
SEXP rstub_dot_call(const char *name, int nargs, SEXP *args) {
    g_err[0] = '\0';
    g_jmp_armed = 1;
    if (setjmp(g_jmp)) {
        g_jmp_armed = 0;
        return NULL;
    }
    SEXP res = NULL;
    int found = 0;
    for (int i = 0; i < g_ncalls; ++i) {
        if (strcmp(g_calls[i].name, name) == 0) {
            found = 1;
            if (g_calls[i].numArgs != nargs)
                Rf_error("Incorrect number of arguments (%d), expecting %d for '%s'", nargs, g_calls[i].numArgs, name);
            res = dispatch(g_calls[i].fun, nargs, args);
        }
    }
    if (!found) Rf_error("C symbol name \"%s\" not in registered .Call table", name);
    g_jmp_armed = 0;
    return res;
}
