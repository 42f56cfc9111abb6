// Initial per-gene fit and imputation, R/fit_ztrunc.R:2-134 (ztruncnorm), R/SIMPLE.R:15-73 (init_impute),
// R/SIMPLE_B.R:23-114 (init_impute_bulk):
//   k_init_stats : per (gene, cluster) moments of the entries > cutoff and of all entries (cell-additive,
//                  deterministic split partials)
//   k_ztrunc     : one thread per (gene, cluster): Brent minimisation of the profile likelihoods LL / LLwZ
//                  (optim(method = "Brent") = Netlib fmin), the pg rules, the NA fallbacks
//   k_init_draw  : one thread per entry <= cutoff: the two-way (init_impute) or three-way (bulk) draw by inverse CDF
// Y2 is [n][G] (R's column-major G x n as is).
#include <cfloat>

#include "rng.cuh"
#include "state.h"

namespace sb {
namespace {

constexpr int GB = 64;     // genes per CTA in k_init_stats
constexpr int NST = 5;     // n1, sum(x - c), sum (x - c)^2 over x > c;  sum x, sum x^2 over all

__global__ void __launch_bounds__(GB) k_init_stats(const double* __restrict__ Y2, const int32_t* __restrict__ clus0, int n,
                                                   int G, int M0, double cutoff, double* __restrict__ part) {
    extern __shared__ __align__(16) double acc[];   // [M0][NST][GB]
    const int g = blockIdx.x * GB + threadIdx.x;
    for (int e = threadIdx.x; e < M0 * NST * GB; e += GB) acc[e] = 0.0;
    __syncthreads();
    const int per = (n + gridDim.y - 1) / gridDim.y;
    const int i0 = blockIdx.y * per, i1 = min(n, i0 + per);
    if (g < G) {
        for (int i = i0; i < i1; ++i) {
            const double x = Y2[size_t(i) * G + g];
            double* a = acc + (size_t(clus0[i]) * NST) * GB + threadIdx.x;
            if (x > cutoff) {
                const double d = x - cutoff;
                a[0] += 1.0;
                a[GB] += d;
                a[2 * GB] = fma(d, d, a[2 * GB]);
            }
            a[3 * GB] += x;
            a[4 * GB] = fma(x, x, a[4 * GB]);
        }
        double* out = part + size_t(blockIdx.y) * M0 * NST * G;
        for (int e = 0; e < M0 * NST; ++e) out[size_t(e) * G + g] = acc[e * GB + threadIdx.x];
    }
}

__device__ double log_pnorm(double x) {
    if (x > 0.0) return log1p(-pnorm_full(-x));
    if (x > -20.0) return log(pnorm_full(x));
    return -0.5 * x * x + log(0.5 * erfcx(-x * 0.70710678118654752440));
}

struct Prof {   // profile likelihood in sigma: LL (R/fit_ztrunc.R:2-10) when r01 < 0, LLwZ (:18-25) otherwise
    double xbar, Sx, delta, r01, p;
    bool wz;
    __device__ double operator()(double sigma) const {
        const double mu = xbar + (Sx - sigma * sigma) / (xbar - delta);
        const double dm = xbar - mu;
        double v = log(sigma) + (Sx + dm * dm) / 2.0 / (sigma * sigma);
        if (!wz) v += log_pnorm((mu - delta) / sigma);
        else v -= r01 * log(1.0 - p + p * pnorm_full((delta - mu) / sigma));
        return isfinite(v) ? v : DBL_MAX;      // optimize(): "NA/Inf replaced by maximum positive value"
    }
};

// Brent's fmin (Netlib; the algorithm behind stats::optimize / optim(method = "Brent")), tol = sqrt(DBL_EPSILON)
__device__ double brent_fmin(const Prof& f, double a, double b) {
    const double c = (3.0 - sqrt(5.0)) * 0.5;
    const double eps = sqrt(DBL_EPSILON), tol3 = sqrt(DBL_EPSILON) / 3.0;
    double v = a + c * (b - a), w = v, x = v, d = 0.0, e = 0.0, u;
    double fx = f(x), fv = fx, fw = fx;
    for (int it = 0; it < 1000; ++it) {
        const double xm = (a + b) * 0.5;
        const double tol1 = eps * fabs(x) + tol3, t2 = tol1 * 2.0;
        if (fabs(x - xm) <= t2 - (b - a) * 0.5) break;
        double p = 0.0, q = 0.0, r = 0.0;
        if (fabs(e) > tol1) {
            r = (x - w) * (fx - fv);
            q = (x - v) * (fx - fw);
            p = (x - v) * q - (x - w) * r;
            q = (q - r) * 2.0;
            if (q > 0.0) p = -p;
            else q = -q;
            r = e;
            e = d;
        }
        if (fabs(p) >= fabs(q * 0.5 * r) || p <= q * (a - x) || p >= q * (b - x)) {
            e = (x < xm) ? b - x : a - x;
            d = c * e;
        } else {
            d = p / q;
            u = x + d;
            if (u - a < t2 || b - u < t2) d = (x < xm) ? tol1 : -tol1;
        }
        if (fabs(d) >= tol1) u = x + d;
        else if (d > 0.0) u = x + tol1;
        else u = x - tol1;
        const double fu = f(u);
        if (fu <= fx) {
            if (u < x) b = x;
            else a = x;
            v = w; w = x; x = u;
            fv = fw; fw = fx; fx = fu;
        } else {
            if (u < x) a = u;
            else b = u;
            if (fu <= fw || w == x) {
                v = w; fv = fw;
                w = u; fw = fu;
            } else if (fu <= fv || v == x || v == w) {
                v = u; fv = fu;
            }
        }
    }
    return x;
}

__global__ void k_ztrunc(InitFitArgs a) {
    const int t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= a.G * a.M0) return;
    const int m = t / a.G, g = t - m * a.G;
    const double* st = a.stats + size_t(m) * NST * a.G + g;
    const double n1 = st[0], s1 = st[a.G], s2 = st[2 * size_t(a.G)], a1 = st[3 * size_t(a.G)], a2 = st[4 * size_t(a.G)];
    const double nm = a.stats[size_t(a.M0) * NST * a.G + m];          // cells of this cluster (all ranks)
    const double NA = nan("");
    const double c = a.cutoff;
    // moments of the positive part (R/fit_ztrunc.R:59-62)
    const double xbar = n1 > 0.0 ? c + s1 / n1 : NA;
    const double Sx = n1 >= 2.0 ? fmax(s2 / n1 - (s1 / n1) * (s1 / n1), 0.0) : NA;
    const double p_min = a.pg1 ? 0.01 : a.p_min;                      // R/SIMPLE_B.R:34
    const double p_max = a.pg1 ? a.pg1[size_t(m) * a.G + g] : 1.0;
    double mu = NA, sd = NA;
    if (n1 >= 4.0) {                                                  // :70-93
        Prof f{xbar, Sx, c, 0.0, 0.0, false};
        sd = brent_fmin(f, a.s_lower, a.s_upper);
        mu = xbar + (Sx - sd * sd) / (xbar - c);
    }
    const double r01 = (nm - n1) / n1, p1 = n1 / nm;                  // :96-98
    const double p11 = pnorm_full((mu - c) / sd);
    double pg = p1 / p11;                                             // :100
    bool refit = false;
    if (pg > p_max || isnan(pg)) {                                    // :102-103
        pg = p_max;
        refit = true;
    }
    if (pg < p_min) {                                                 // :104-105
        pg = p_min;
        refit = true;
    }
    if (refit) {                                                      // :106-131
        if (isnan(Sx)) {
            mu = -1.0;
            sd = 0.5;
        } else {
            Prof f{xbar, Sx, c, r01, pg, true};
            sd = brent_fmin(f, a.s_lower, a.s_upper);
            mu = xbar + (Sx - sd * sd) / (xbar - c);
        }
    }
    // callers' post-processing
    const bool na = isnan(mu);
    if (na) {                                                         // R/SIMPLE.R:30-36, R/SIMPLE_B.R:40-45
        mu = a1 / nm;
        sd = nm > 1.0 ? sqrt(fmax((a2 - a1 * a1 / nm) / (nm - 1.0), 0.0)) : NA;
    }
    double pg_used, pg_ret;
    if (!a.pg1) {
        pg_used = na ? 1.0 : fmin(pg, 0.99);                          // R/SIMPLE.R:28,36
        pg_ret = (m == a.M0 - 1) ? pg_used : fmin(pg_used, 0.99);     // :28 is re-applied to the whole matrix on later passes
    } else {
        pg_used = pg_ret = na ? p_max : pg;                           // R/SIMPLE_B.R:45
    }
    a.mu[t] = mu;
    a.sd[t] = sd;
    a.pg_used[t] = pg_used;
    a.pg_ret[t] = pg_ret;
}

__global__ void __launch_bounds__(256) k_init_draw(InitDrawArgs a) {
    const size_t e = blockIdx.x * size_t(blockDim.x) + threadIdx.x;
    if (e >= size_t(a.n) * a.G) return;
    const int i = int(e / a.G), g = int(e - size_t(i) * a.G);
    double x = a.Y2[e];
    if (x <= a.cutoff) {
        const int m = a.clus0[i];
        const size_t t = size_t(m) * a.G + g;
        const double ms = a.mu[t], sds = a.sd[t], p = a.pg_used[t];
        double ua, ub;
        entry_uniforms(a.seed, a.sweep, (unsigned long long)(a.cell_offset + i), unsigned(g), ua, ub);
        const double r = (a.cutoff - ms) / sds;
        const double prob = pnorm(r);
        const double den = prob * p + (1.0 - p);
        double z;
        int kind;                                   // 0 normal(ms, sds), 1 N(-1, 0.5), 2 truncated
        if (!a.pg1) {
            kind = (ua < (1.0 - p) / den) ? 0 : 2;                                     // R/SIMPLE.R:49-51
        } else {
            const double q1 = a.pg1[t];
            const double pd = (p / q1 - p) / den, pz = (1.0 - p / q1) / den;            // R/SIMPLE_B.R:75-78
            kind = (ua < pd) ? 0 : ((ua < pd + pz) ? 1 : 2);
        }
        if (kind == 2) {
            z = (r < -30.0) ? trunc_z_deep(ub, r) : fmin(qnorm(ub * prob), r);          // rtnorm(upper = cutoff)
            x = ms + sds * z;
        } else {
            z = qnorm(ub);
            x = kind == 0 ? ms + sds * z : -1.0 + 0.5 * z;
        }
        if (isnan(x)) x = 0.0;                                                          // R/SIMPLE.R:69, R/SIMPLE_B.R:108
    } else if (isnan(x)) {
        x = 0.0;
    }
    a.out[e] = x;
}

}  // namespace

int init_stats_splits(int n) { return n >= 4096 ? 32 : (n >= 256 ? 4 : 1); }

void launch_init_stats(const double* Y2, const int32_t* clus0, int n, int G, int M0, double cutoff, double* part, int nsplit,
                       cudaStream_t st) {
    const size_t smem = sizeof(double) * size_t(M0) * NST * GB;
    ensure_dyn_smem_k(k_init_stats, smem);
    dim3 grid((G + GB - 1) / GB, nsplit);
    k_init_stats<<<grid, GB, smem, st>>>(Y2, clus0, n, G, M0, cutoff, part);
}

void launch_ztrunc(const InitFitArgs& a, cudaStream_t st) { k_ztrunc<<<(a.G * a.M0 + 63) / 64, 64, 0, st>>>(a); }

void launch_init_draw(const InitDrawArgs& a, cudaStream_t st) {
    const size_t tot = size_t(a.n) * a.G;
    k_init_draw<<<unsigned((tot + 255) / 256), 256, 0, st>>>(a);
}

}  // namespace sb
