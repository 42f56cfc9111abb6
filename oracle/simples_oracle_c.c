/* Multi-threaded C restatement of the O(n G) loops of SIMPLEs' EM imputation hot path (TEST / BASELINE INFRASTRUCTURE).
 *
 * This is the CPU baseline of BASELINE.md section 3 ("a multi-threaded build of the oracle"): the same arithmetic as
 * oracle/simples_oracle.py -- which follows R/EM_impute.R:75-339 of JunLiuLab/SIMPLEs line by line -- for the three loops
 * that touch every (gene, cell) entry, parallelised over cells with OpenMP:
 *
 *   so_sweep_entries   R/EM_impute.R:173-198 (= R/do_impute.R:112-136): per zero entry the conditional mean / sd, the
 *                      dropout posterior, the Bernoulli draw, the normal / truncated-normal draw, the clamps and the
 *                      re-centring.  Variates come from the Philox4x32-10 tape of oracle.entry_uniforms.
 *   so_sq_over_sigma   the quadratic term sum_g Y_gi^2 / sigma_g of R/EM_impute.R:128 (expanded around mu, see
 *                      oracle/accel.py)
 *   so_rowsq_weighted  sum_i Y_gi^2 w_i, the second-moment statistic of the sigma update R/EM_impute.R:263-267
 *
 * Everything of size K0 x K0, M0 or G x K0 (cluster constants, lasso, lambda, mu, pi) and the BLAS-shaped products stay in
 * NumPy / OpenBLAS (oracle/accel.py).  Only tests/, __graft_entry__ and bench.py's CPU arms may load this library; the
 * product (simples_b200/) never does.  PARITY UNPINNED: it is checked against the NumPy oracle (tests/test_oracle_c.py),
 * which itself has never met an R output.
 *
 * Special functions: Phi = erfc of libm (R/EM_impute.R:179 pnorm), Phi^-1 = Wichura's AS241 PPND16 (the algorithm of R's
 * qnorm; published coefficients), deep lower tail (r < -30) in log space exactly as oracle.trunc_z.
 *
 * Build: gcc -O2 -fopenmp -shared -fPIC (oracle/build_c.py).  Matrices are column-major G x n (R's layout).
 */
#include <math.h>
#include <stdint.h>
#include <stddef.h>
#ifdef _OPENMP
#include <omp.h>
#endif

int so_abi_version(void) { return 1; }

int so_max_threads(void) {
#ifdef _OPENMP
    return omp_get_max_threads();
#else
    return 1;
#endif
}

/* ---- Philox4x32-10 (Salmon et al. 2011), same tape as oracle.philox4x32 / csrc/rng.cuh ---- */
static inline void philox4x32_10(uint32_t c[4], uint32_t k0, uint32_t k1) {
    for (int r = 0; r < 10; ++r) {
        const uint64_t p0 = (uint64_t)0xD2511F53u * c[0];
        const uint64_t p1 = (uint64_t)0xCD9E8D57u * c[2];
        const uint32_t n0 = (uint32_t)(p1 >> 32) ^ c[1] ^ k0;
        const uint32_t n1 = (uint32_t)p1;
        const uint32_t n2 = (uint32_t)(p0 >> 32) ^ c[3] ^ k1;
        const uint32_t n3 = (uint32_t)p0;
        c[0] = n0; c[1] = n1; c[2] = n2; c[3] = n3;
        k0 += 0x9E3779B9u;
        k1 += 0xBB67AE85u;
    }
}

/* two 32-bit words -> uniform in (0,1): ((x >> 12) + 0.5) 2^-52  (oracle._u52) */
static inline double u52(uint32_t hi, uint32_t lo) {
    const uint64_t x = ((uint64_t)hi << 32) | lo;
    return ((double)(x >> 12) + 0.5) * 2.220446049250313080847263336181640625e-16;
}

/* ---- Phi^-1: AS241 PPND16 ---- */
static const double QA[8] = {3.3871328727963666080, 1.3314166789178437745e+2, 1.9715909503065514427e+3,
                             1.3731693765509461125e+4, 4.5921953931549871457e+4, 6.7265770927008700853e+4,
                             3.3430575583588128105e+4, 2.5090809287301226727e+3};
static const double QB[8] = {1.0, 4.2313330701600911252e+1, 6.8718700749205790830e+2, 5.3941960214247511077e+3,
                             2.1213794301586595867e+4, 3.9307895800092710610e+4, 2.8729085735721942674e+4,
                             5.2264952788528545610e+3};
static const double QC[8] = {1.42343711074968357734, 4.63033784615654529590, 5.76949722146069140550,
                             3.64784832476320460504, 1.27045825245236838258, 2.41780725177450611770e-1,
                             2.27238449892691845833e-2, 7.74545014278341407640e-4};
static const double QD[8] = {1.0, 2.05319162663775882187, 1.67638483018380384940, 6.89767334985100004550e-1,
                             1.48103976427480074590e-1, 1.51986665636164571966e-2, 5.47593808499534494600e-4,
                             1.05075007164441684324e-9};
static const double QE[8] = {6.65790464350110377720, 5.46378491116411436990, 1.78482653991729133580,
                             2.96560571828504891230e-1, 2.65321895265761230930e-2, 1.24266094738807843860e-3,
                             2.71155556874348757815e-5, 2.01033439929228813265e-7};
static const double QF[8] = {1.0, 5.99832206555887937690e-1, 1.36929880922735805310e-1, 1.48753612908506148525e-2,
                             7.86869131145613259100e-4, 1.84631831751005468180e-5, 1.42151175831644588870e-7,
                             2.04426310338993978564e-15};

static inline double horner8(const double* c, double x) {
    double r = c[7];
    for (int i = 6; i >= 0; --i) r = r * x + c[i];
    return r;
}

double so_qnorm(double p) {
    const double q = p - 0.5;
    if (fabs(q) <= 0.425) {
        const double r = 0.180625 - q * q;
        return q * horner8(QA, r) / horner8(QB, r);
    }
    const double pp = q < 0.0 ? p : 1.0 - p;
    double r = sqrt(-log(pp));
    double v;
    if (r <= 5.0) {
        r -= 1.6;
        v = horner8(QC, r) / horner8(QD, r);
    } else {
        r -= 5.0;
        v = horner8(QE, r) / horner8(QF, r);
    }
    return q < 0.0 ? -v : v;
}

double so_pnorm(double x) { return 0.5 * erfc(-x * 0.70710678118654752440); }

/* erfcx(x) = exp(x^2) erfc(x) for x >= 15 (asymptotic series; the terms fall below 1e-18 after six) */
static inline double erfcx_large(double x) {
    const double h = 1.0 / (2.0 * x * x);
    double s = 1.0, t = 1.0;
    for (int k = 1; k <= 8; ++k) {
        t *= -(2.0 * k - 1.0) * h;
        s += t;
    }
    return s / (x * 1.77245385090551602730);
}

/* z with Phi(z) = u Phi(r), capped at r (oracle.trunc_z; R/EM_impute.R:191-192 msm::rtnorm(upper = cutoff)) */
double so_trunc_z(double u, double r) {
    double z;
    if (r < -30.0) {
        const double s2 = 1.41421356237309504880, s2pi = 2.50662827463100050242;
        const double t = log(u) + log(0.5 * erfcx_large(-r / s2)) - 0.5 * r * r;
        z = -sqrt(r * r - 2.0 * log(u));
        for (int i = 0; i < 4; ++i) {
            const double e = 0.5 * erfcx_large(-z / s2);
            z = z - (log(e) - 0.5 * z * z - t) * e * s2pi;
        }
    } else {
        z = so_qnorm(u * so_pnorm(r));
    }
    return z < r ? z : r;
}

/* One MC imputation sweep over the entries with mask != 0 (R/EM_impute.R:173-198).
 *   Yc     [G x n] column-major, centred / scaled matrix, updated in place
 *   mask   [G x n] column-major bytes (Y0 <= cutoff)
 *   im     [n]     sampled cluster of the cell (0-based)        f [n x K0] row-major sampled factors
 *   mu     [G x M0], beta [G x K0], sigma [G x M0], pg [G x MB] row-major;  celltype0 [n] 0-based
 *   seed / sweep / cell0: Philox key and counters of oracle.entry_uniforms (global cell index = cell0 + i)
 * Returns the number of entries drawn. */
long long so_sweep_entries(double* Yc, const unsigned char* mask, long long G, long long n, int K0, int M0, int MB,
                           const long long* im, const double* f, const double* mu, const double* beta,
                           const double* sigma, const double* pg, const long long* celltype0, const double* gene_mean,
                           const double* gene_sd, double cutoff, unsigned long long seed, unsigned sweep,
                           unsigned long long cell0, double lower, double upper, int clip) {
    const uint32_t k0 = (uint32_t)seed, k1 = (uint32_t)(seed >> 32);
    long long drawn = 0;
#pragma omp parallel for schedule(dynamic, 64) reduction(+ : drawn)
    for (long long i = 0; i < n; ++i) {
        const unsigned char* mk = mask + (size_t)i * G;
        double* y = Yc + (size_t)i * G;
        const int m = (int)im[i], ct = (int)celltype0[i];
        const double* fi = f + (size_t)i * K0;
        const unsigned long long cell = cell0 + (unsigned long long)i;
        for (long long g = 0; g < G; ++g) {
            if (!mk[g]) continue;
            const double* b = beta + (size_t)g * K0;
            double dot = 0.0;
            for (int k = 0; k < K0; ++k) dot += b[k] * fi[k];
            const double ms = (mu[(size_t)g * M0 + m] + dot) * gene_sd[g] + gene_mean[g];   /* :175 */
            const double sds = sqrt(sigma[(size_t)g * M0 + m]) * gene_sd[g];                /* :176 */
            const double p = pg[(size_t)g * MB + ct];                                       /* :177 */
            const double r = (cutoff - ms) / sds;
            const double prob = so_pnorm(r);                                                /* :179 */
            const double prob_drop = (1.0 - p) / (prob * p + (1.0 - p));                    /* :180 */
            uint32_t c[4] = {(uint32_t)g, (uint32_t)cell, sweep, (uint32_t)(cell >> 32) << 8};
            philox4x32_10(c, k0, k1);
            const double ua = u52(c[0], c[1]), ub = u52(c[2], c[3]);
            double x;
            if (ua < prob_drop) x = ms + sds * so_qnorm(ub);                                /* :181-186 */
            else x = ms + sds * so_trunc_z(ub, r);                                          /* :191-192 */
            if (clip) {
                x = x < upper ? x : upper;                                                  /* :195 */
                x = x > lower ? x : lower;                                                  /* :196 */
            }
            y[g] = (x - gene_mean[g]) / gene_sd[g];                                         /* :198 */
            ++drawn;
        }
    }
    return drawn;
}

/* out[s * n + i] = sum_g Y_gi^2 * w[s * G + g]   (w = 1 / sigma of the NS distinct sigma columns) */
void so_sq_over_sigma(const double* Yc, long long G, long long n, const double* w, int NS, double* out) {
#pragma omp parallel for schedule(static)
    for (long long i = 0; i < n; ++i) {
        const double* y = Yc + (size_t)i * G;
        for (int s = 0; s < NS; ++s) {
            const double* ws = w + (size_t)s * G;
            double a0 = 0.0, a1 = 0.0, a2 = 0.0, a3 = 0.0;
            long long g = 0;
            for (; g + 4 <= G; g += 4) {
                a0 += y[g] * y[g] * ws[g];
                a1 += y[g + 1] * y[g + 1] * ws[g + 1];
                a2 += y[g + 2] * y[g + 2] * ws[g + 2];
                a3 += y[g + 3] * y[g + 3] * ws[g + 3];
            }
            for (; g < G; ++g) a0 += y[g] * y[g] * ws[g];
            out[(size_t)s * n + i] = (a0 + a1) + (a2 + a3);
        }
    }
}

/* out[g] = sum_i Y_gi^2 * w[i]; `scratch` holds nthreads * G doubles (so_max_threads()) */
void so_rowsq_weighted(const double* Yc, long long G, long long n, const double* w, double* out, double* scratch) {
    int nt = 1;
#pragma omp parallel
    {
#ifdef _OPENMP
        const int t = omp_get_thread_num();
#pragma omp single
        nt = omp_get_num_threads();
#else
        const int t = 0;
#endif
        double* acc = scratch + (size_t)t * G;
        for (long long g = 0; g < G; ++g) acc[g] = 0.0;
#pragma omp for schedule(static)
        for (long long i = 0; i < n; ++i) {
            const double* y = Yc + (size_t)i * G;
            const double wi = w[i];
            for (long long g = 0; g < G; ++g) acc[g] += y[g] * y[g] * wi;
        }
    }
    for (long long g = 0; g < G; ++g) {
        double s = 0.0;
        for (int t = 0; t < nt; ++t) s += scratch[(size_t)t * G + g];
        out[g] = s;
    }
}

/* vectorised helpers for the tests: elementwise so_qnorm / so_pnorm / so_trunc_z */
void so_qnorm_v(const double* p, long long n, double* out) {
    for (long long i = 0; i < n; ++i) out[i] = so_qnorm(p[i]);
}
void so_pnorm_v(const double* x, long long n, double* out) {
    for (long long i = 0; i < n; ++i) out[i] = so_pnorm(x[i]);
}
void so_trunc_z_v(const double* u, const double* r, long long n, double* out) {
    for (long long i = 0; i < n; ++i) out[i] = so_trunc_z(u[i], r[i]);
}
