/*
 * simples_b200 -- C ABI of the B200-native SIMPLEs EM imputation hot path.
 *
 * The reference (JunLiuLab/SIMPLEs) is a pure-R package with no FFI
 * (NAMESPACE:1-14).  The boundary this library replaces is therefore the R
 * function boundary of the two internal hot-path functions
 *
 *   EM_impute(Y, Y0, pg, M0, K0, cutoff, iter, beta, sigma, lambda, pi, z,
 *             mu, celltype, penl, est_z, max_lambda, est_lam, impt_it,
 *             sigma0, pi_alpha, verbose, num_mc, lower, upper)
 *                                                    R/EM_impute.R:75-78
 *   do_impute(dat, Y, beta, lambda, sigma, mu, pi, pos_mean, pos_sd,
 *             celltype, mcmc, burnin, verbose, pg, cutoff)
 *                                                    R/do_impute.R:29-30
 *
 * A `.Call` shim (r/simples_rcall.c, shipped as source) marshals the R
 * arguments 1:1 into the structs below; see INTEGRATION.md.
 *
 * Conventions
 *   - All matrices are FP64, COLUMN-MAJOR (R layout), owned by the caller.
 *     Inputs are never modified (R value semantics).
 *   - `celltype` is 1-based (R), NULL = all cells are type 1.
 *   - Every entry point returns 0 on success, a negative SIMPLES_E* code on
 *     failure; simples_last_error() gives the message (thread-local).
 *   - There is NO CPU fallback: without a usable sm_100 device every compute
 *     entry point fails with SIMPLES_ENODEV.
 *   - Multi-GPU, two ways (cells are sharded, parameters replicated, exactly
 *     one all-reduce of the M-step statistics per EM iteration plus two at
 *     call start: gene means, default mu; NCCL is loaded with dlopen):
 *     (1) ONE process drives several GPUs: simples_init(ndev > 1, devs).
 *         simples_em_impute / simples_do_impute then take the FULL host
 *         matrices, shard the cells over the devices with one host thread per
 *         device, all-reduce through NCCL inside the library and return the
 *         full results.  This is the route of an R session (.Call).
 *     (2) one process per GPU (torchrun): each process passes its contiguous
 *         shard of cells (columns of Y/Y0, rows of z, slice of celltype) with
 *         `n_total`/`cell_offset`.  The sum over ranks is either the library's
 *         own rank communicator (simples_comm_init_rank) or an `allreduce`
 *         callback that sums a device buffer of doubles in place.
 *   - Limits: K0 <= 32, M0 <= 32 (register-tiled kernels), G, n < 2^31 per
 *     device shard, n_total < 2^63.
 */
#ifndef SIMPLES_B200_H
#define SIMPLES_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define SIMPLES_ABI_VERSION 1
#define SIMPLES_MAX_DEVICES 16

#if defined(__GNUC__)
#define SIMPLES_API __attribute__((visibility("default")))
#else
#define SIMPLES_API
#endif

#define SIMPLES_OK        0
#define SIMPLES_EINVAL   -1   /* bad argument                          */
#define SIMPLES_ENODEV   -2   /* no CUDA device / not initialised      */
#define SIMPLES_ECUDA    -3   /* CUDA runtime error                    */
#define SIMPLES_ENOMEM   -4   /* device or host allocation failed      */
#define SIMPLES_ENUMERIC -5   /* non-finite log-likelihood             */
#define SIMPLES_ECOMM    -6   /* all-reduce callback failed            */

/* ref_compat_flags: reproduce reference quirks; default all on */
#define SIMPLES_COMPAT_STALE_DT     1u  /* Q1  R/EM_impute.R:206,212          */
#define SIMPLES_COMPAT_PI           2u  /* Q2  PI = 3.14159265359, :80        */
#define SIMPLES_COMPAT_MU_DIV_N     4u  /* Q3  default mu = Y z / n, :95      */
#define SIMPLES_COMPAT_PRIORB_PREV  8u  /* Q5  tot[it] -= previous priorB     */
#define SIMPLES_COMPAT_DEFAULT     15u
/* Not a quirk but a mode, carried in the same word: the Monte-Carlo sweeps become deterministic -- most probable
 * cluster, factor at its posterior mean, every zero entry set to its conditional expectation
 * ms - sds p phi(r) / (p Phi(r) + 1 - p) (R/SIMPLE.R:59 hints at it; the reference itself samples).  EM_impute only. */
#define SIMPLES_MODE_EXPECT        16u

/* Sum `count` doubles at device pointer `dev_buf` over all ranks, in place,
 * ordered after prior work on `stream` (a cudaStream_t) and before later
 * work on it.  Return 0 on success. */
typedef int (*simples_allreduce_fn)(void *user, double *dev_buf, size_t count, void *stream);

/* ---- EM_impute (R/EM_impute.R:75-339) ---------------------------------- */
typedef struct simples_em_args {
    int32_t G, n, M0, K0, MB;       /* genes, local cells, clusters, factors, ncol(pg) */
    const double *Y;                /* G x n   initial imputed matrix                  */
    const double *Y0;               /* G x n   observed matrix (mask = Y0 <= cutoff)   */
    const double *pg;               /* G x MB  P(normal component) per cell type       */
    double cutoff;
    int32_t iter;
    const double *beta;             /* G x K0 */
    const double *sigma;            /* G x M0 */
    const double *lambda;           /* M0 x K0 */
    const double *pi;               /* M0 */
    const double *z;                /* n x M0, or NULL (=> computed at it = 1)         */
    const double *mu;               /* G x M0, or NULL (=> Y z / n, R/EM_impute.R:93-99) */
    const int32_t *celltype;        /* n, 1-based, or NULL                             */
    double penl;
    int32_t est_z, max_lambda, est_lam, impt_it;
    double sigma0, pi_alpha;
    int32_t num_mc;
    double lower, upper;
    uint64_t seed;                  /* Philox key of the variate tape                  */
    uint32_t ref_compat_flags;
    int64_t n_total;                /* global cell count (0 => n)                      */
    int64_t cell_offset;            /* global index of local cell 0                    */
    simples_allreduce_fn allreduce; /* NULL => single rank                             */
    void *allreduce_user;
    void *stream;                   /* cudaStream_t, NULL => library-owned stream      */
} simples_em_args;

/* Return list of EM_impute (R/EM_impute.R:337-338); any pointer may be NULL
 * to skip that output. */
typedef struct simples_em_out {
    double *loglik;                 /* iter                                            */
    double *pi;                     /* M0                                              */
    double *mu;                     /* G x M0                                          */
    double *sigma;                  /* G x M0                                          */
    double *beta;                   /* G x K0                                          */
    double *lambda;                 /* M0 x K0                                         */
    double *z;                      /* n x M0                                          */
    double *Ef;                     /* M0 matrices n x K0, back to back                */
    double *Varf;                   /* M0 matrices K0 x K0, back to back               */
    double *Y;                      /* G x n   last imputed sample, un-centred         */
    double *geneM;                  /* G                                               */
    double *geneSd;                 /* G (all 1, R/EM_impute.R:89)                     */
    double *priorB;                 /* 1                                               */
} simples_em_out;

/* ---- do_impute (R/do_impute.R:29-160) ---------------------------------- */
typedef struct simples_mi_args {
    int32_t G, n, M0, K0, MB;
    const double *dat;              /* G x n observed                                  */
    const double *Y;                /* G x n initial imputed (un-centred)              */
    const double *beta, *lambda, *sigma, *mu, *pi;
    const double *pos_mean;         /* G or NULL (=> 0)                                */
    const double *pos_sd;           /* G or NULL (=> 1)                                */
    const int32_t *celltype;        /* n, 1-based, or NULL                             */
    int32_t mcmc, burnin;
    const double *pg;               /* G x MB (scalar pg is broadcast by the host layer, :47-48) */
    double cutoff;
    uint64_t seed;
    uint32_t ref_compat_flags;
    int32_t want_consensus;         /* materialise the n x n consensus matrix (single rank only) */
    int64_t n_total, cell_offset;
    simples_allreduce_fn allreduce;
    void *allreduce_user;
    void *stream;
} simples_mi_args;

typedef struct simples_mi_out {     /* R/do_impute.R:158-159 */
    double *loglik;                 /* 1: mean(tot[-1:-burnin]); burnin = 0 still drops
                                       the first sweep, as R's index -1:0 does          */
    double *impt;                   /* G x n                                           */
    double *impt_var;               /* G x n                                           */
    double *EF;                     /* n x K0                                          */
    double *varF;                   /* n x K0                                          */
    double *consensus_cluster;      /* n x n or NULL                                   */
} simples_mi_out;

/* ---- low-quality-gene fit between the two EM_impute calls ----------------
 * R/SIMPLE.R:305-411 (= R/SIMPLE_B.R:332-419): with the high-quality-gene fit
 * (z, Ef, Varf) fixed, every remaining gene gets M0 unpenalised intercepts and
 * K0 lasso loadings from one augmented regression (:312-366), a shared variance
 * (:365, clipped :371-373), and its zero entries are drawn once with the factor
 * held at its posterior mean (:376-411).  All arrays hold the lq genes only
 * (Gq rows); cells may be sharded exactly like EM_impute. */
typedef struct simples_lq_args {
    int32_t Gq, n, M0, K0, MB;
    const double *dat;              /* Gq x n observed rows dat[lq_ind, ]               */
    const double *resp;             /* Gq x n regression response: NULL => dat; init_imp[lq_ind, ] when SIMPLE() got one */
    const double *sigma;            /* Gq x M0 current sigma[lq_ind, ]                  */
    const double *pg;               /* Gq x MB                                          */
    const double *z;                /* n x M0   impute_hq$z                             */
    const double *Ef;               /* M0 matrices n x K0, back to back                 */
    const double *Varf;             /* M0 matrices K0 x K0, back to back                */
    const int32_t *celltype;        /* n, 1-based, or NULL                              */
    double penl, cutoff;
    int32_t resid_mode;             /* 0: sg without the residual sum (SIMPLE() with init_imp = NULL, the inverted
                                       branch of R/SIMPLE.R:354-359); 1: residual of dat (R/SIMPLE.R:360-364,
                                       R/SIMPLE_B.R:370-374)                            */
    int32_t impute;                 /* 0: regression only; 1: also the draw of :376-411 */
    uint64_t seed;
    uint32_t sweep;                 /* tape position of the draw (any value not used by the EM calls) */
    int64_t n_total, cell_offset;
    simples_allreduce_fn allreduce;
    void *allreduce_user;
    void *stream;
} simples_lq_args;

typedef struct simples_lq_out {
    double *mu;                     /* Gq x M0                                          */
    double *beta;                   /* Gq x K0                                          */
    double *sigma;                  /* Gq x M0 (columns equal)                          */
    double *Y;                      /* Gq x n  dat with the entries <= cutoff redrawn (NULL when impute = 0) */
} simples_lq_out;

/* ---- initial per-gene fit + imputation ------------------------------------
 * init_impute (R/SIMPLE.R:15-73) when pg1 = NULL, init_impute_bulk
 * (R/SIMPLE_B.R:23-114) when pg1 is given; both fit every gene within every
 * cluster by ztruncnorm (R/fit_ztrunc.R:55-134: Brent on the profile
 * likelihoods LL / LLwZ) and redraw the entries <= cutoff once. */
typedef struct simples_init_args {
    int32_t G, n, M0;
    const double *Y2;               /* G x n                                            */
    const int32_t *clus;            /* n, labels 1..M0                                  */
    const double *pg1;              /* NULL, or G x M0 upper bounds of pg (bulk variant) */
    double p_min, cutoff;           /* p_min is ignored (0.01) in the bulk variant      */
    uint64_t seed;
    uint32_t sweep;                 /* tape position of the draw                        */
    int64_t n_total, cell_offset;
    simples_allreduce_fn allreduce;
    void *allreduce_user;
    void *stream;
} simples_init_args;

typedef struct simples_init_out {
    double *impute;                 /* G x n                                            */
    double *pg;                     /* G x M0: init_impute's second return value; the per-cluster fitted pg in the bulk variant */
    double *mu, *sd;                /* G x M0 fitted per-cluster mean / sd (optional, may be NULL) */
} simples_init_out;

/* ---- initial clustering of the cells -------------------------------------
 * R/SIMPLE.R:258-266 (= R/SIMPLE_B.R:293-302): rows of X standardised
 * (scale), rank-K truncated SVD (irlba), cellmat = V diag(d), k-means with
 * `nstart` random starts, best total within-cluster sum of squares wins.
 * The SVD is the eigen-problem of the G x G Gram matrix (subspace iteration);
 * k-means is Lloyd's algorithm, all restarts advanced together.  The
 * reference's irlba start vector and Hartigan-Wong restarts are random, so
 * agreement with it is at the level of the partition (ARI), not of numbers. */
typedef struct simples_clus_args {
    int32_t G, n, K, M0;
    const double *X;                /* G x n (hq genes x cells)                         */
    int32_t nstart, iter_max;       /* kmeans(nstart = 300, iter.max = 80)              */
    uint64_t seed;
    int64_t n_total, cell_offset;
    simples_allreduce_fn allreduce;
    void *allreduce_user;
    void *stream;
} simples_clus_args;

typedef struct simples_clus_out {
    int32_t *clus;                  /* n labels in 1..M0                                */
    double *cellmat;                /* n x K  (V diag(d); column signs arbitrary) or NULL */
    double *d;                      /* K singular values or NULL                        */
    double *tot_withinss;           /* 1, of the winning start, or NULL                 */
} simples_clus_out;

/* ---- lifecycle / errors ------------------------------------------------ */
SIMPLES_API int simples_abi_version(void);
/* Select the device(s) this process drives.  ndev = 1: one process per GPU.
 * ndev in 2..SIMPLES_MAX_DEVICES: this process drives all of them (NCCL
 * communicators created here); the one-shot EM / MI entry points shard the
 * cells over them, the other entry points run on devs[0]. */
SIMPLES_API int simples_init(int ndev, const int *devs);
SIMPLES_API void simples_shutdown(void);
SIMPLES_API int simples_device_count(void);
/* One process per GPU: build this process's rank communicator inside the
 * library.  Rank 0 calls simples_comm_unique_id (128 bytes), the host layer
 * broadcasts the bytes, every rank calls simples_comm_init_rank after
 * simples_init.  Entry points called with n_total > n and allreduce = NULL
 * then reduce over it. */
SIMPLES_API int simples_comm_unique_id(void *id128, size_t cap);
SIMPLES_API int simples_comm_init_rank(int nranks, int rank, const void *id128);
SIMPLES_API void simples_comm_destroy(void);
SIMPLES_API int simples_comm_nranks(void);
SIMPLES_API int simples_nccl_version(void);          /* e.g. 22809, or -1 when libnccl cannot be loaded */
SIMPLES_API const char *simples_last_error(void);

/* ---- one-shot calls with HOST buffers (what the .Call shim binds) ------- */
SIMPLES_API int simples_em_impute(const simples_em_args *args, simples_em_out *out);
SIMPLES_API int simples_do_impute(const simples_mi_args *args, simples_mi_out *out);
SIMPLES_API int simples_lq_fit(const simples_lq_args *args, simples_lq_out *out);
SIMPLES_API int simples_init_impute(const simples_init_args *args, simples_init_out *out);
SIMPLES_API int simples_init_cluster(const simples_clus_args *args, simples_clus_out *out);

/* ---- resident-context API (keeps Y/mask/state in HBM between calls; used
 *      by SIMPLE()-level drivers and by bench.py's device-resident timing) -- */
typedef struct simples_ctx simples_ctx;

SIMPLES_API int simples_em_create(const simples_em_args *args, simples_ctx **ctx);   /* alloc + H2D + prologue */
SIMPLES_API int simples_em_run(simples_ctx *ctx, int iters);                         /* next `iters` EM iterations */
SIMPLES_API int simples_em_fetch(simples_ctx *ctx, simples_em_out *out);             /* D2H of the return list */
SIMPLES_API int simples_em_sync(simples_ctx *ctx);
/* Device time of the last simples_em_run, per iteration (ms, CUDA events on
 * the context's stream).  `cap` = capacity of ms_per_iter; returns count. */
SIMPLES_API int simples_em_iter_ms(simples_ctx *ctx, double *ms_per_iter, int cap);
/* Per-kernel profile of the last run when enabled: for each kernel class k,
 * launches[k] and total_ms[k].  Classes are listed by simples_kernel_name(). */
SIMPLES_API int simples_em_set_profile(simples_ctx *ctx, int enable);
SIMPLES_API int simples_em_profile(simples_ctx *ctx, int64_t *launches, double *total_ms, int cap);
SIMPLES_API const char *simples_kernel_name(int k);
SIMPLES_API int simples_kernel_count(void);
/* Driver-side consumers of the EM result (R/SIMPLE.R:421-430, R/SIMPLE_B.R:429-437):
 * Yimp0 = sum_m z_im (mu_m + B Ef_m[i,]) * geneSd + geneM  (G x n, this shard's cells) from the resident state
 * (post-M-step mu / beta with the z / Ef of the last E-pass, exactly as the reference mixes them), and the two
 * BIC variants from the last log-likelihood and the GLOBAL n. */
SIMPLES_API int simples_em_yimp0(simples_ctx *ctx, double *Yimp0);
SIMPLES_API int simples_bic(double loglik_last, int64_t n_total, int32_t G, int32_t K0, int32_t M0, double *bic, double *bic0);
/* Number of entries with Y0 <= cutoff in this context's shard. */
SIMPLES_API int64_t simples_em_nnz0(simples_ctx *ctx);
/* CUDA kernels enqueued by the last simples_em_run on this context. */
SIMPLES_API int64_t simples_em_kernel_launches(simples_ctx *ctx);
SIMPLES_API void simples_ctx_destroy(simples_ctx *ctx);
/* Host-side helper, no GPU involved: bit j of mask[w * n + i] <=> Y0[32 w + j, i] <= cutoff for a column-major G x n
 * HOST matrix (ceil(G/32) x n words).  This is the only thing the library needs from the observed matrix
 * (R/EM_impute.R:173): when Y0 / dat live in host memory the entry points pack it with host threads while Y is uploading
 * and only the mask crosses PCIe.  Returns the number of set bits, or a negative error code. */
SIMPLES_API int64_t simples_pack_mask(const double *Y0, int32_t G, int32_t n, double cutoff, uint32_t *mask);

#ifdef __cplusplus
}
#endif
#endif /* SIMPLES_B200_H */
