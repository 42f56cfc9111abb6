/*
 * .Call shim between R and libsimples_b200.so (include/simples_b200.h).
 *
 * SHIPPED AS SOURCE: neither R nor its headers exist in the build image.  tests/test_r_shim.py compiles this file
 * unmodified against a stand-in for the R API (tests/helpers/rstub) and drives it on the GPU; R itself never ran it.
 * A maintainer of JunLiuLab/SIMPLEs adds it as src/simples_rcall.c, links with -lsimples_b200, and replaces the bodies of R/EM_impute.R:75-339 and R/do_impute.R:29-160
 * with r/EM_impute_b200.R.  Argument order is the reference's own (R/EM_impute.R:75-78,
 * R/do_impute.R:29-30); every matrix is passed as the REAL() pointer of an R double matrix
 * (column-major), which is exactly the layout the C ABI takes, so no copies are made on the way in.
 *
 * Error convention: the library returns a negative status and never longjmps; Rf_error() is raised
 * here only after the library call has returned (no C++ frames on the stack).
 */
#include <R.h>
#include <Rinternals.h>
#include <R_ext/Rdynload.h>
#include <stdint.h>
#include <string.h>
#include "simples_b200.h"

static const double *opt_real(SEXP x) { return Rf_isNull(x) ? NULL : REAL(x); }

SEXP C_simples_em_impute(SEXP Y, SEXP Y0, SEXP pg, SEXP M0, SEXP K0, SEXP cutoff, SEXP iter, SEXP beta, SEXP sigma,
                         SEXP lambda, SEXP pi, SEXP z, SEXP mu, SEXP celltype, SEXP penl, SEXP est_z, SEXP max_lambda,
                         SEXP est_lam, SEXP impt_it, SEXP sigma0, SEXP pi_alpha, SEXP num_mc, SEXP lower, SEXP upper,
                         SEXP seed) {
    simples_em_args a;
    simples_em_out o;
    memset(&a, 0, sizeof a);
    memset(&o, 0, sizeof o);
    a.G = Rf_nrows(Y);
    a.n = Rf_ncols(Y);
    a.M0 = Rf_asInteger(M0);
    a.K0 = Rf_asInteger(K0);
    a.MB = Rf_ncols(pg);
    a.Y = REAL(Y);
    a.Y0 = REAL(Y0);
    a.pg = REAL(pg);
    a.cutoff = Rf_asReal(cutoff);
    a.iter = Rf_asInteger(iter);
    a.beta = REAL(beta);
    a.sigma = REAL(sigma);
    a.lambda = REAL(lambda);
    a.pi = REAL(pi);
    a.z = opt_real(z);
    a.mu = opt_real(mu);
    a.celltype = Rf_isNull(celltype) ? NULL : (const int32_t *)INTEGER(celltype);   /* 1-based, as in R */
    a.penl = Rf_asReal(penl);
    a.est_z = Rf_asInteger(est_z);
    a.max_lambda = Rf_asLogical(max_lambda);
    a.est_lam = Rf_asInteger(est_lam);
    a.impt_it = Rf_asInteger(impt_it);
    a.sigma0 = Rf_asReal(sigma0);
    a.pi_alpha = Rf_asReal(pi_alpha);
    a.num_mc = Rf_asInteger(num_mc);
    a.lower = Rf_asReal(lower);
    a.upper = Rf_asReal(upper);
    a.seed = (uint64_t)Rf_asReal(seed);
    a.ref_compat_flags = SIMPLES_COMPAT_DEFAULT;

    const int G = a.G, n = a.n, M = a.M0, K = a.K0;
    const char *names[] = {"loglik", "pi", "mu", "sigma", "beta", "lambda", "z", "Ef", "Varf", "Y", "geneM", "geneSd",
                           "priorB", ""};
    SEXP res = PROTECT(Rf_mkNamed(VECSXP, names));                  /* R/EM_impute.R:337-338 */
    SEXP s_ll = PROTECT(Rf_allocVector(REALSXP, a.iter));
    SEXP s_pi = PROTECT(Rf_allocVector(REALSXP, M));
    SEXP s_mu = PROTECT(Rf_allocMatrix(REALSXP, G, M));
    SEXP s_sg = PROTECT(Rf_allocMatrix(REALSXP, G, M));
    SEXP s_be = PROTECT(Rf_allocMatrix(REALSXP, G, K));
    SEXP s_la = PROTECT(Rf_allocMatrix(REALSXP, M, K));
    SEXP s_z = PROTECT(Rf_allocMatrix(REALSXP, n, M));
    SEXP s_Ef = PROTECT(Rf_allocVector(REALSXP, (R_xlen_t)n * K * M));   /* M0 n x K0 matrices back to back */
    SEXP s_Vf = PROTECT(Rf_allocVector(REALSXP, (R_xlen_t)K * K * M));
    SEXP s_Y = PROTECT(Rf_allocMatrix(REALSXP, G, n));
    SEXP s_gm = PROTECT(Rf_allocVector(REALSXP, G));
    SEXP s_gs = PROTECT(Rf_allocVector(REALSXP, G));
    SEXP s_pb = PROTECT(Rf_allocVector(REALSXP, 1));
    o.loglik = REAL(s_ll); o.pi = REAL(s_pi); o.mu = REAL(s_mu); o.sigma = REAL(s_sg); o.beta = REAL(s_be);
    o.lambda = REAL(s_la); o.z = REAL(s_z); o.Ef = REAL(s_Ef); o.Varf = REAL(s_Vf); o.Y = REAL(s_Y);
    o.geneM = REAL(s_gm); o.geneSd = REAL(s_gs); o.priorB = REAL(s_pb);

    const int rc = simples_em_impute(&a, &o);
    if (rc != SIMPLES_OK) {
        UNPROTECT(14);
        Rf_error("simples_b200: %s (status %d)", simples_last_error(), rc);
    }
    /* Ef / Varf become lists of M0 matrices, as in the reference (R/EM_impute.R:109-110,337) */
    SEXP l_Ef = PROTECT(Rf_allocVector(VECSXP, M)), l_Vf = PROTECT(Rf_allocVector(VECSXP, M));
    for (int m = 0; m < M; ++m) {
        SEXP e = PROTECT(Rf_allocMatrix(REALSXP, n, K)), v = PROTECT(Rf_allocMatrix(REALSXP, K, K));
        memcpy(REAL(e), REAL(s_Ef) + (size_t)m * n * K, sizeof(double) * (size_t)n * K);
        memcpy(REAL(v), REAL(s_Vf) + (size_t)m * K * K, sizeof(double) * (size_t)K * K);
        SET_VECTOR_ELT(l_Ef, m, e);
        SET_VECTOR_ELT(l_Vf, m, v);
        UNPROTECT(2);
    }
    SEXP parts[] = {s_ll, s_pi, s_mu, s_sg, s_be, s_la, s_z, l_Ef, l_Vf, s_Y, s_gm, s_gs, s_pb};
    for (int i = 0; i < 13; ++i) SET_VECTOR_ELT(res, i, parts[i]);
    UNPROTECT(16);
    return res;
}

SEXP C_simples_do_impute(SEXP dat, SEXP Y, SEXP beta, SEXP lambda, SEXP sigma, SEXP mu, SEXP pi, SEXP pos_mean,
                         SEXP pos_sd, SEXP celltype, SEXP mcmc, SEXP burnin, SEXP pg, SEXP cutoff, SEXP seed,
                         SEXP want_consensus) {
    simples_mi_args a;
    simples_mi_out o;
    memset(&a, 0, sizeof a);
    memset(&o, 0, sizeof o);
    a.G = Rf_nrows(Y);
    a.n = Rf_ncols(Y);
    a.M0 = Rf_ncols(mu);                                             /* R/do_impute.R:35 */
    a.K0 = Rf_ncols(beta);                                           /* R/do_impute.R:36 */
    a.MB = Rf_ncols(pg);                                             /* scalar pg is expanded in R, :47-48 */
    a.dat = REAL(dat); a.Y = REAL(Y); a.beta = REAL(beta); a.lambda = REAL(lambda); a.sigma = REAL(sigma);
    a.mu = REAL(mu); a.pi = REAL(pi); a.pos_mean = opt_real(pos_mean); a.pos_sd = opt_real(pos_sd);
    a.celltype = Rf_isNull(celltype) ? NULL : (const int32_t *)INTEGER(celltype);
    a.mcmc = Rf_asInteger(mcmc); a.burnin = Rf_asInteger(burnin); a.pg = REAL(pg); a.cutoff = Rf_asReal(cutoff);
    a.seed = (uint64_t)Rf_asReal(seed);
    a.ref_compat_flags = SIMPLES_COMPAT_DEFAULT;
    a.want_consensus = Rf_asLogical(want_consensus);
    const int G = a.G, n = a.n, K = a.K0;
    const char *names[] = {"loglik", "impt", "impt_var", "EF", "varF", "consensus_cluster", ""};
    SEXP res = PROTECT(Rf_mkNamed(VECSXP, names));                  /* R/do_impute.R:158-159 */
    SEXP s_ll = PROTECT(Rf_allocVector(REALSXP, 1));
    SEXP s_im = PROTECT(Rf_allocMatrix(REALSXP, G, n));
    SEXP s_iv = PROTECT(Rf_allocMatrix(REALSXP, G, n));
    SEXP s_ef = PROTECT(Rf_allocMatrix(REALSXP, n, K));
    SEXP s_vf = PROTECT(Rf_allocMatrix(REALSXP, n, K));
    SEXP s_cc = PROTECT(a.want_consensus ? Rf_allocMatrix(REALSXP, n, n) : R_NilValue);
    o.loglik = REAL(s_ll); o.impt = REAL(s_im); o.impt_var = REAL(s_iv); o.EF = REAL(s_ef); o.varF = REAL(s_vf);
    o.consensus_cluster = a.want_consensus ? REAL(s_cc) : NULL;
    const int rc = simples_do_impute(&a, &o);
    if (rc != SIMPLES_OK) {
        UNPROTECT(7);
        Rf_error("simples_b200: %s (status %d)", simples_last_error(), rc);
    }
    SEXP parts[] = {s_ll, s_im, s_iv, s_ef, s_vf, s_cc};
    for (int i = 0; i < 6; ++i) SET_VECTOR_ELT(res, i, parts[i]);
    UNPROTECT(7);
    return res;
}

/* init_impute (R/SIMPLE.R:15-73) when pg1 is NULL, init_impute_bulk (R/SIMPLE_B.R:23-114) otherwise.
 * Returns list(impute, pg). */
SEXP C_simples_init_impute(SEXP Y2, SEXP M0, SEXP clus, SEXP pg1, SEXP p_min, SEXP cutoff, SEXP seed) {
    simples_init_args a;
    simples_init_out o;
    memset(&a, 0, sizeof a);
    memset(&o, 0, sizeof o);
    a.G = Rf_nrows(Y2); a.n = Rf_ncols(Y2); a.M0 = Rf_asInteger(M0);
    a.Y2 = REAL(Y2); a.clus = (const int32_t *)INTEGER(clus); a.pg1 = opt_real(pg1);
    a.p_min = Rf_asReal(p_min); a.cutoff = Rf_asReal(cutoff); a.seed = (uint64_t)Rf_asReal(seed);
    const char *names[] = {"impute", "pg", ""};
    SEXP res = PROTECT(Rf_mkNamed(VECSXP, names));
    SEXP s_im = PROTECT(Rf_allocMatrix(REALSXP, a.G, a.n));
    SEXP s_pg = PROTECT(Rf_allocMatrix(REALSXP, a.G, a.M0));
    o.impute = REAL(s_im); o.pg = REAL(s_pg);
    const int rc = simples_init_impute(&a, &o);
    if (rc != SIMPLES_OK) {
        UNPROTECT(3);
        Rf_error("simples_b200: %s (status %d)", simples_last_error(), rc);
    }
    SET_VECTOR_ELT(res, 0, s_im); SET_VECTOR_ELT(res, 1, s_pg);
    UNPROTECT(3);
    return res;
}

/* kmeans(irlba(t(scale(t(X))), nv = K)$v %*% diag(d), M0, iter.max, nstart)$cluster (R/SIMPLE.R:258-266) */
SEXP C_simples_init_cluster(SEXP X, SEXP K, SEXP M0, SEXP nstart, SEXP iter_max, SEXP seed) {
    simples_clus_args a;
    simples_clus_out o;
    memset(&a, 0, sizeof a);
    memset(&o, 0, sizeof o);
    a.G = Rf_nrows(X); a.n = Rf_ncols(X); a.K = Rf_asInteger(K); a.M0 = Rf_asInteger(M0);
    a.X = REAL(X); a.nstart = Rf_asInteger(nstart); a.iter_max = Rf_asInteger(iter_max);
    a.seed = (uint64_t)Rf_asReal(seed);
    SEXP s_cl = PROTECT(Rf_allocVector(INTSXP, a.n));
    o.clus = (int32_t *)INTEGER(s_cl);
    const int rc = simples_init_cluster(&a, &o);
    if (rc != SIMPLES_OK) {
        UNPROTECT(1);
        Rf_error("simples_b200: %s (status %d)", simples_last_error(), rc);
    }
    UNPROTECT(1);
    return s_cl;
}

/* The foreach(g = lq_ind) regression and the for (i in 1:n) draw of R/SIMPLE.R:312-411 (= R/SIMPLE_B.R:335-419).
 * Ef / Varf are the R lists of impute_hq; they are packed back to back here.  Returns list(mu, beta, sigma, Y). */
SEXP C_simples_lq_fit(SEXP dat, SEXP resp, SEXP sigma, SEXP pg, SEXP z, SEXP Ef, SEXP Varf, SEXP celltype, SEXP penl,
                      SEXP cutoff, SEXP resid_mode, SEXP seed) {
    simples_lq_args a;
    simples_lq_out o;
    memset(&a, 0, sizeof a);
    memset(&o, 0, sizeof o);
    a.Gq = Rf_nrows(dat); a.n = Rf_ncols(dat); a.M0 = Rf_ncols(z); a.MB = Rf_ncols(pg);
    a.K0 = Rf_ncols(VECTOR_ELT(Ef, 0));
    const int Gq = a.Gq, n = a.n, M = a.M0, K = a.K0;
    double *ef = (double *)R_alloc((size_t)M * n * K, sizeof(double));
    double *vf = (double *)R_alloc((size_t)M * K * K, sizeof(double));
    for (int m = 0; m < M; ++m) {
        memcpy(ef + (size_t)m * n * K, REAL(VECTOR_ELT(Ef, m)), sizeof(double) * (size_t)n * K);
        memcpy(vf + (size_t)m * K * K, REAL(VECTOR_ELT(Varf, m)), sizeof(double) * (size_t)K * K);
    }
    a.dat = REAL(dat); a.resp = opt_real(resp); a.sigma = REAL(sigma); a.pg = REAL(pg); a.z = REAL(z);
    a.Ef = ef; a.Varf = vf;
    a.celltype = Rf_isNull(celltype) ? NULL : (const int32_t *)INTEGER(celltype);
    a.penl = Rf_asReal(penl); a.cutoff = Rf_asReal(cutoff); a.resid_mode = Rf_asInteger(resid_mode); a.impute = 1;
    a.seed = (uint64_t)Rf_asReal(seed);
    const char *names[] = {"mu", "beta", "sigma", "Y", ""};
    SEXP res = PROTECT(Rf_mkNamed(VECSXP, names));
    SEXP s_mu = PROTECT(Rf_allocMatrix(REALSXP, Gq, M));
    SEXP s_be = PROTECT(Rf_allocMatrix(REALSXP, Gq, K));
    SEXP s_sg = PROTECT(Rf_allocMatrix(REALSXP, Gq, M));
    SEXP s_y = PROTECT(Rf_allocMatrix(REALSXP, Gq, n));
    o.mu = REAL(s_mu); o.beta = REAL(s_be); o.sigma = REAL(s_sg); o.Y = REAL(s_y);
    const int rc = simples_lq_fit(&a, &o);
    if (rc != SIMPLES_OK) {
        UNPROTECT(5);
        Rf_error("simples_b200: %s (status %d)", simples_last_error(), rc);
    }
    SEXP parts[] = {s_mu, s_be, s_sg, s_y};
    for (int i = 0; i < 4; ++i) SET_VECTOR_ELT(res, i, parts[i]);
    UNPROTECT(5);
    return res;
}

/* simples_init(ndev, devs): devices this R session drives.  With more than one device the library shards the cells of
 * EM_impute / do_impute over them (host thread per GPU + NCCL inside the library); the R code above it is unchanged. */
SEXP C_simples_init(SEXP devices) {
    SEXP d = PROTECT(Rf_coerceVector(devices, INTSXP));
    const int rc = simples_init(Rf_length(d), INTEGER(d));
    UNPROTECT(1);
    if (rc != SIMPLES_OK) Rf_error("simples_b200: %s (status %d)", simples_last_error(), rc);
    return Rf_ScalarInteger(simples_device_count());
}

static const R_CallMethodDef call_methods[] = {{"C_simples_init", (DL_FUNC)&C_simples_init, 1},
                                               {"C_simples_em_impute", (DL_FUNC)&C_simples_em_impute, 25},
                                               {"C_simples_do_impute", (DL_FUNC)&C_simples_do_impute, 16},
                                               {"C_simples_init_impute", (DL_FUNC)&C_simples_init_impute, 7},
                                               {"C_simples_init_cluster", (DL_FUNC)&C_simples_init_cluster, 6},
                                               {"C_simples_lq_fit", (DL_FUNC)&C_simples_lq_fit, 12},
                                               {NULL, NULL, 0}};

void R_init_SIMPLEs(DllInfo *dll) {
    R_registerRoutines(dll, NULL, call_methods, NULL, NULL);
    R_useDynamicSymbols(dll, FALSE);
}
