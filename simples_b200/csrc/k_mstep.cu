// M-step from (all-reduced) sufficient statistics -- R/EM_impute.R:234-318 without the augmented
// designs:
//   k_gene_solve : per gene lasso for the loadings (replaces glmnet, :256-261), sg (:263-267),
//                  priorB_g (:268), sigma clip (:273-274).  One thread per gene.
//   k_lmp        : priorB / loglik bookkeeping (Q5), lambda step (replaces solnp, :278-304), pi (:318).
//   k_mu         : Woodbury mu update without the G x G bigM (:307-315).  One CTA per cluster.
#include "linalg.cuh"
#include "ptx.cuh"
#include "state.h"

namespace sb {
namespace {

constexpr int GW = 8;   // warps (= genes) per CTA in k_gene_solve

struct StatView {   // pointers into the stats buffer
    const double *C, *R2, *U2, *S, *a, *nz, *c, *tot;
    int ldc;        // row stride of C
};

__device__ __host__ inline StatView stat_view(const double* st, const Dims& d, int general) {
    StatView v;
    const size_t nS = size_t(d.M0) * d.K0 * d.K0, nA = size_t(d.M0) * d.K0;
    v.C = st;
    if (!general) {
        v.ldc = d.NVP;
        v.R2 = st + size_t(d.ldG) * d.NVP;
        v.U2 = nullptr;
        v.S = v.R2 + d.ldG;
    } else {
        v.ldc = d.K0 * d.M0 + 2 * d.M0;
        v.R2 = nullptr;
        v.U2 = st + size_t(d.ldG) * v.ldc;
        v.S = v.U2 + size_t(d.ldG) * d.M0;
    }
    v.a = v.S + nS;
    v.nz = v.a + nA;
    v.c = v.nz + d.M0;
    v.tot = v.c + d.M0;
    return v;
}

// Exact active-set polish of a coordinate-descent iterate (one warp, lane k owns coordinate k):
// solve A_SS x = rhs_S - kappa sign_S and verify the KKT conditions.  On success b is overwritten.
// Returns 0 on success (b overwritten), 1 when the restricted system is not positive definite / nothing to adopt,
// 2 when a coordinate of the restricted solution changes sign (the support is too large), 3 when an inactive coordinate
// violates the KKT conditions (the support is too small).  For 2 and 3 `cand` receives the restricted solution of the
// lane's coordinate with sign-flipped coordinates set to 0 -- a much better restart point for coordinate descent than
// waiting for a coefficient to creep to zero.
__device__ int lasso_polish(const double* A, int K, double rhs, double kappa, double& b, double* L, double* xs, int* idx,
                            int lane, double& cand) {
    const bool kl = lane < K;
    cand = b;
    const unsigned act = __ballot_sync(0xffffffffu, kl && b != 0.0);
    const int ns = __popc(act);
    if (ns == 0) return __ballot_sync(0xffffffffu, kl && fabs(rhs) > kappa * (1.0 + 1e-12)) == 0u ? 0 : 1;
    const int pos = __popc(act & ((1u << lane) - 1u));
    const bool mine = (act >> lane) & 1u;
    if (mine) {
        idx[pos] = lane;
        xs[pos] = rhs - kappa * (b > 0.0 ? 1.0 : -1.0);
    }
    __syncwarp();
    for (int e = lane; e < ns * ns; e += 32) {
        const int i = e / ns, j = e - i * ns;
        L[e] = A[idx[i] * K + idx[j]];
    }
    __syncwarp();
    warp_cholesky(L, ns, lane);
    bool bad = false;
    for (int i = 0; i < ns; ++i) bad |= !(L[i * ns + i] > 0.0);
    if (bad) return 1;
    // (L L') x = xs: forward / backward substitution, each row a warp-wide dot product (ns <= 32)
    for (int i = 0; i < ns; ++i) {
        const double part = (lane < i) ? L[i * ns + lane] * xs[lane] : 0.0;
        const double sdot = warp_sum(part);
        __syncwarp();
        if (lane == 0) xs[i] = (xs[i] - sdot) / L[i * ns + i];
        __syncwarp();
    }
    for (int i = ns - 1; i >= 0; --i) {
        const double part = (lane > i && lane < ns) ? L[lane * ns + i] * xs[lane] : 0.0;
        const double sdot = warp_sum(part);
        __syncwarp();
        if (lane == 0) xs[i] = (xs[i] - sdot) / L[i * ns + i];
        __syncwarp();
    }
    const double xm = mine ? xs[pos] : 0.0;
    const bool flip = mine && ((xm > 0.0) != (b > 0.0) || xm == 0.0);
    cand = flip ? 0.0 : xm;
    if (__ballot_sync(0xffffffffu, flip)) return 2;
    double gk = rhs;
    if (kl && !mine)
        for (int i = 0; i < ns; ++i) gk -= A[lane * K + idx[i]] * xs[i];
    if (__ballot_sync(0xffffffffu, kl && !mine && fabs(gk) > kappa * (1.0 + 1e-10) + 1e-300)) return 3;
    if (mine) b = xm;
    return 0;
}

// min 0.5 b'Ab - rhs'b + kappa |b|_1 : cyclic coordinate descent from 0 (glmnet's start) + exact polish.
__device__ int g_lasso_dbg[4];   // SIMPLES_TRACE diagnostics: [0] solves, [1] total CD sweeps, [2] max sweeps, [3] capped

__device__ __forceinline__ void lasso_note(int sweeps, int lane, int debug) {
    if (debug && lane == 0) {     // 4 same-address atomics per gene serialise in L2: diagnostics only (SIMPLES_TRACE)
        atomicAdd(&g_lasso_dbg[0], 1);
        atomicAdd(&g_lasso_dbg[1], sweeps);
        atomicMax(&g_lasso_dbg[2], sweeps);
        if (sweeps >= 20000) atomicAdd(&g_lasso_dbg[3], 1);
    }
}

__device__ double lasso_solve(const double* A, int K, double rhs, double kappa, double* L, double* xs, int* idx, int lane,
                              int debug) {
    double b = 0.0, r = rhs;
    const bool kl = lane < K;
    // lane k owns coordinate k: its diagonal entry and reciprocal stay in registers, so a coordinate step is
    // [own candidate: fma, soft threshold, multiply] -> ONE shuffle of the step -> residual update; the round-1 loop
    // broadcast r_k and b_k (two shuffles) and divided by A_kk inside the chain (the kernel lasts as long as its slowest
    // gene, so the length of this chain is the kernel's duration)
    const double akk = kl ? A[lane * K + lane] : 1.0, ikk = 1.0 / akk;
    unsigned pat_p = 0u, pat_n = 0u, fail_p = ~0u, fail_n = ~0u;
    int stable = 0, last_fail = -1000, adopted = 0;
    for (int sweep = 0; sweep < 20000; ++sweep) {
        double maxd = 0.0;
        for (int k = 0; k < K; ++k) {
            const double t = fma(akk, b, r);
            const double at = fabs(t) - kappa;
            const double bn = at > 0.0 ? copysign(at, t) * ikk : 0.0;
            const double dlt = __shfl_sync(0xffffffffu, bn - b, k);
            if (dlt != 0.0) {
                if (kl) r = fma(-dlt, A[lane * K + k], r);
                if (lane == k) b = bn;
            }
            maxd = fmax(maxd, fabs(dlt));
        }
        const double maxb = fmax(1.0, warp_max(kl ? fabs(b) : 0.0));
        const bool done = maxd <= 1e-15 * maxb;
        // The exact polish costs several sweeps and can only succeed once the active set has settled: try it when the
        // sign pattern has been the same for three sweeps (a pattern that failed is retried 16 sweeps later, when the
        // inactive coordinates' gradients have moved), and when coordinate descent has converged by itself.
        const unsigned np = __ballot_sync(0xffffffffu, kl && b > 0.0), nn = __ballot_sync(0xffffffffu, kl && b < 0.0);
        stable = (np == pat_p && nn == pat_n) ? stable + 1 : 0;
        pat_p = np;
        pat_n = nn;
        const bool fresh = !(np == fail_p && nn == fail_n) || sweep - last_fail >= 16;
        if (done || (stable >= 2 && fresh)) {
            double cand;
            const int st = lasso_polish(A, K, rhs, kappa, b, L, xs, idx, lane, cand);
            if (st == 0 || done) {
                lasso_note(sweep + 1, lane, debug);
                return b;
            }
            fail_p = np;
            fail_n = nn;
            last_fail = sweep;
            if (st >= 2 && adopted < 12) {
                // Active-set step: continue from the exact solution on the current support with the sign-flipped
                // coordinates dropped (st = 2) / the violating coordinates about to enter by the next sweep (st = 3).
                // Acceptance is still only by the KKT check above, so this can shorten the path but not change the answer;
                // the objective is convex, coordinate descent converges from any restart point.
                ++adopted;
                b = kl ? cand : 0.0;
                double rr = rhs;
                for (int j = 0; j < K; ++j) {
                    const double bj = __shfl_sync(0xffffffffu, b, j);
                    if (kl) rr = fma(-bj, A[lane * K + j], rr);
                }
                r = rr;
                stable = 0;
                fail_p = fail_n = ~0u;
                last_fail = -1000;
            }
        }
    }
    lasso_note(20000, lane, debug);
    return b;
}

__global__ void __launch_bounds__(GW * 32) k_gene_solve(MstepArgs a) {
    extern __shared__ __align__(16) double sm[];
    const Dims& d = a.d;
    const int K0 = d.K0, M0 = d.M0;
    const StatView sv = stat_view(a.stats, d, a.general);
    const int KK = K0 * K0;
    double* sAsum = sm;                       // [K0][K0]  sum_m (S_m + nz_m M_m)
    double* sa = sAsum + KK;                  // [M0][K0]
    double* sred = sa + M0 * K0;              // [GW]
    double* sL = sred + GW;                   // [GW][K0*K0]
    double* sAl = sL + GW * KK;               // [GW][K0*K0]   (general path)
    double* sxs = sAl + (a.general ? GW * KK : 0);   // [GW][K0]
    int* sidx = reinterpret_cast<int*>(sxs + GW * K0);   // [GW][K0]
    for (int e = threadIdx.x; e < KK; e += blockDim.x) {
        double s = 0.0;
        for (int m = 0; m < M0; ++m) s += sv.S[size_t(m) * KK + e] + sv.nz[m] * a.Mm[size_t(m) * KK + e];
        sAsum[e] = s;
    }
    for (int e = threadIdx.x; e < M0 * K0; e += blockDim.x) sa[e] = sv.a[e];
    __syncthreads();

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const bool kl = lane < K0, ml = lane < M0;
    const int g = blockIdx.x * GW + warp;
    double prior = 0.0;
    if (g < d.G) {
        const double ntot = double(d.n_total);
        const double N = double(M0) * ntot + K0;
        const double* Cg = sv.C + size_t(g) * sv.ldc;
        const double mu = ml ? a.mu[size_t(g) * M0 + lane] : 0.0;       // lane m
        const double sgm = ml ? a.sigma[size_t(g) * M0 + lane] : 1.0;   // lane m
        const double nzm = ml ? sv.nz[lane] : 0.0, cm = ml ? sv.c[lane] : 0.0;
        double rhs = 0.0, bt = 0.0, d2, sumY, sumY2, var, kappa;
        const double* A;
        double* L = sL + warp * KK;
        if (!a.general) {
            const double sg0 = __shfl_sync(0xffffffffu, sgm, 0);
            double s = kl ? Cg[lane] : 0.0;
            for (int m = 0; m < M0; ++m) {
                const double mum = __shfl_sync(0xffffffffu, mu, m);
                if (kl) s -= mum * sa[m * K0 + lane];
            }
            bt = rhs = s;
            const double U = ml ? Cg[K0 + lane] : 0.0;
            d2 = sv.R2[g] + warp_sum(-2.0 * mu * U + mu * mu * nzm);
            sumY = (Cg[K0 + M0] - warp_sum(mu * cm)) / sqrt(sg0);      // sum_m (Uh[g,m] - mu_gm c_m) / sqrt(sigma_g)
            sumY2 = d2 / sg0;
            var = (sumY2 - sumY * sumY / N) / (N - 1.0);
            kappa = a.penl * var * sg0;     // objective scaled by sigma_g
            A = sAsum;
        } else {
            double* Al = sAl + warp * KK;
            for (int e = lane; e < KK; e += 32) {
                double s = 0.0;
                for (int m = 0; m < M0; ++m)
                    s += (sv.S[size_t(m) * KK + e] + sv.nz[m] * a.Mm[size_t(m) * KK + e]) / a.sigma[size_t(g) * M0 + m];
                Al[e] = s;
            }
            __syncwarp();
            double s = 0.0, u = 0.0;
            for (int m = 0; m < M0; ++m) {
                const double mum = __shfl_sync(0xffffffffu, mu, m), sg_m = __shfl_sync(0xffffffffu, sgm, m);
                if (kl) {
                    const double bm = Cg[m * K0 + lane] - mum * sa[m * K0 + lane];
                    s += bm / sg_m;
                    u += bm;
                }
            }
            rhs = s;
            bt = u;
            const double U = ml ? Cg[K0 * M0 + lane] : 0.0, Uh = ml ? Cg[K0 * M0 + M0 + lane] : 0.0;
            const double D2 = ml ? sv.U2[size_t(g) * M0 + lane] - 2.0 * mu * U + mu * mu * nzm : 0.0;
            d2 = warp_sum(D2);
            sumY2 = warp_sum(D2 / sgm);
            sumY = warp_sum((Uh - mu * cm) / sqrt(sgm));
            var = (sumY2 - sumY * sumY / N) / (N - 1.0);
            kappa = a.penl * var;
            A = Al;
        }
        const double b = lasso_solve(A, K0, rhs, kappa, L, sxs + warp * K0, sidx + warp * K0, lane, a.debug);
        // sg = (sum_m D2 - 2 nb.(sum_m bm) + nb' (sum_m C_m) nb + 1) / (n + 3)
        double ab = 0.0;
        for (int k = 0; k < K0; ++k) {
            const double bk = __shfl_sync(0xffffffffu, b, k);
            if (kl) ab = fma(sAsum[lane * K0 + k], bk, ab);
        }
        const double quad = warp_sum(kl ? b * ab : 0.0);
        const double lin = warp_sum(kl ? b * bt : 0.0);
        const double l1 = warp_sum(kl ? fabs(b) : 0.0);
        const double sg = (d2 - 2.0 * lin + quad + 1.0) / (ntot + 3.0);
        prior = a.penl * var / sg * l1;
        if (kl) a.beta[size_t(g) * K0 + lane] = b;
        const double sgc = fmin(fmax(sg, 1e-4), 9.0);
        if (ml) a.sigma[size_t(g) * M0 + lane] = sgc;
    }
    if (lane == 0) sred[warp] = prior;
    __syncthreads();
    if (threadIdx.x == 0) {
        double s = 0.0;
        for (int i = 0; i < GW; ++i) s += sred[i];
        a.part_gene[blockIdx.x] = s;
    }
}

// Low-quality-gene regression, R/SIMPLE.R:312-366 (= R/SIMPLE_B.R:335-378), from cell-additive statistics.  The
// intercept block of the augmented Gram matrix is diagonal (nz_m / sigma_gm), so the M0 unpenalised intercepts are
// profiled out exactly and a K0-dimensional lasso remains:
//   A  = sum_m (S_m + nz_m Varf_m - a_m a_m'/nz_m) / sigma_gm,   rhs = sum_m (T_gm - a_m U_gm / nz_m) / sigma_gm,
//   kappa = penl / var(Y_aug)   (lambda = penalty K0/(M0+K0) with glmnet's penalty.factor rescale, :341-349),
//   mu_gm = (U_gm - a_m'b) / nz_m.
// One warp per gene; lane k owns loading k, lane m owns cluster m.
__global__ void __launch_bounds__(GW * 32) k_lq_solve(LqSolveArgs a) {
    extern __shared__ __align__(16) double sm[];
    const Dims& d = a.d;
    const int K0 = d.K0, M0 = d.M0;
    const StatView sv = stat_view(a.stats, d, 1), svd = stat_view(a.stats_d, d, 1);
    const int KK = K0 * K0;
    double* sAsum = sm;                       // [K0][K0]  sum_m nz_m Varf_m (+ S_m when the residual is kept)
    double* sa = sAsum + KK;                  // [M0][K0]
    double* sL = sa + M0 * K0;                // [GW][K0*K0]
    double* sAl = sL + GW * KK;               // [GW][K0*K0]
    double* sxs = sAl + GW * KK;              // [GW][K0]
    int* sidx = reinterpret_cast<int*>(sxs + GW * K0);   // [GW][K0]
    for (int e = threadIdx.x; e < KK; e += blockDim.x) {
        double s = 0.0;
        for (int m = 0; m < M0; ++m)
            s += sv.nz[m] * a.Varf[size_t(m) * KK + e] + (a.resid_mode ? sv.S[size_t(m) * KK + e] : 0.0);
        sAsum[e] = s;
    }
    for (int e = threadIdx.x; e < M0 * K0; e += blockDim.x) sa[e] = sv.a[e];
    __syncthreads();

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const bool kl = lane < K0, ml = lane < M0;
    const int g = blockIdx.x * GW + warp;
    if (g >= d.G) return;
    const double ntot = double(d.n_total);
    const double N = double(M0) * ntot + K0;
    const double* Cg = sv.C + size_t(g) * sv.ldc;
    const double sgm = ml ? a.sigma[size_t(g) * M0 + lane] : 1.0;   // lane m
    const double nzm = ml ? sv.nz[lane] : 0.0;
    const double inz = nzm > 0.0 ? 1.0 / nzm : 0.0;
    const double U = ml ? Cg[K0 * M0 + lane] : 0.0, Uh = ml ? Cg[K0 * M0 + M0 + lane] : 0.0;
    const double U2 = ml ? sv.U2[size_t(g) * M0 + lane] : 0.0;
    double* Al = sAl + warp * KK;
    for (int e = lane; e < KK; e += 32) {
        const int i = e / K0, j = e - i * K0;
        double s = 0.0;
        for (int m = 0; m < M0; ++m) {
            const double nz = sv.nz[m];
            const double t = sv.S[size_t(m) * KK + e] + nz * a.Varf[size_t(m) * KK + e] -
                             (nz > 0.0 ? sa[m * K0 + i] * sa[m * K0 + j] / nz : 0.0);
            s += t / a.sigma[size_t(g) * M0 + m];
        }
        Al[e] = s;
    }
    __syncwarp();
    double rhs = 0.0;
    for (int m = 0; m < M0; ++m) {
        const double Um = __shfl_sync(0xffffffffu, U, m), sg_m = __shfl_sync(0xffffffffu, sgm, m),
                     iz = __shfl_sync(0xffffffffu, inz, m);
        if (kl) rhs += (Cg[m * K0 + lane] - sa[m * K0 + lane] * Um * iz) / sg_m;
    }
    const double sumY = warp_sum(ml ? Uh / sqrt(sgm) : 0.0);
    const double sumY2 = warp_sum(ml ? U2 / sgm : 0.0);
    const double var = (sumY2 - sumY * sumY / N) / (N - 1.0);
    const double kappa = a.penl / var;
    const double b = lasso_solve(Al, K0, rhs, kappa, sL + warp * KK, sxs + warp * K0, sidx + warp * K0, lane, 0);
    double ab_m = 0.0;                                 // lane m: a_m'b
    for (int k = 0; k < K0; ++k) {
        const double bk = __shfl_sync(0xffffffffu, b, k);
        if (ml) ab_m = fma(sa[lane * K0 + k], bk, ab_m);
    }
    const double mu = ml ? (U - ab_m) * inz : 0.0;
    double ab = 0.0;
    for (int k = 0; k < K0; ++k) {
        const double bk = __shfl_sync(0xffffffffu, b, k);
        if (kl) ab = fma(sAsum[lane * K0 + k], bk, ab);
    }
    double sg = warp_sum(kl ? b * ab : 0.0);           // b'(sum_m Vm [+ S_m]) b
    if (a.resid_mode) {
        const double* Dg = svd.C + size_t(g) * svd.ldc;
        const double Ud = ml ? Dg[K0 * M0 + lane] : 0.0, U2d = ml ? svd.U2[size_t(g) * M0 + lane] : 0.0;
        sg += warp_sum(ml ? U2d - 2.0 * mu * Ud + mu * mu * nzm : 0.0);
        double bt = 0.0;
        for (int m = 0; m < M0; ++m) {
            const double mum = __shfl_sync(0xffffffffu, mu, m);
            if (kl) bt += Dg[m * K0 + lane] - mum * sa[m * K0 + lane];
        }
        sg -= 2.0 * warp_sum(kl ? b * bt : 0.0);
    }
    sg = (sg + 1.0) / (ntot + 3.0);
    if (kl) a.beta[size_t(g) * K0 + lane] = b;
    if (ml) {
        a.mu[size_t(g) * M0 + lane] = mu;
        a.sigma[size_t(g) * M0 + lane] = fmin(fmax(sg, 1e-4), 9.0);
    }
}

// sum of v over lanes 0..mt-1 in ascending lane order (mirrors the oracle's sequential Python sums)
__device__ __forceinline__ double ordered_sum(double v, int mt) {
    double s = 0.0;
    for (int m = 0; m < mt; ++m) s += __shfl_sync(0xffffffffu, v, m);
    return s;
}

// Damped Newton for  min sum(E/x + c log x)  s.t. sum x = T  (mirrors oracle lambda_solve_k operation for
// operation).  One warp per factor, lane m owns component m (mt <= 32).
__device__ void lambda_solve_warp(double E, double c, double& x, int mt, double T, int lane) {
    const bool act = lane < mt;
    if (!act) { E = 1.0; c = 1.0; x = 1.0; }
    double f = ordered_sum(act ? E / x + c * log(x) : 0.0, mt);
    for (int it = 0; it < 500; ++it) {
        const double g = -E / (x * x) + c / x;
        double h = fabs(2.0 * E / (x * x * x) - c / (x * x));
        if (h < 1e-300) h = 1e-300;
        const double sg = ordered_sum(act ? g / h : 0.0, mt);
        const double sh = ordered_sum(act ? 1.0 / h : 0.0, mt);
        const double nu = -sg / sh;
        const double dx = act ? -(g + nu) / h : 0.0;
        const double gd = ordered_sum(g * dx, mt);
        double al = (act && dx < 0.0) ? -0.9 * x / dx : 1.0;
        double dm = fabs(dx);
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            al = fmin(al, __shfl_xor_sync(0xffffffffu, al, o));
            dm = fmax(dm, __shfl_xor_sync(0xffffffffu, dm, o));
        }
        if (dm <= 1e-15 * T) break;
        double fn = f, xn = x;
        bool ok = false;
        for (int bt = 0; bt < 60; ++bt) {
            xn = x + al * dx;
            fn = ordered_sum(act ? E / xn + c * log(xn) : 0.0, mt);
            if (fn <= f + 1e-4 * al * gd + 1e-14 * fabs(f)) {
                ok = true;
                break;
            }
            al *= 0.5;
        }
        if (!ok) break;
        x = xn;
        f = fn;
    }
}

__global__ void __launch_bounds__(1024) k_lmp(MstepArgs a) {
    const Dims& d = a.d;
    const int K0 = d.K0, M0 = d.M0;
    const StatView sv = stat_view(a.stats, d, a.general);
    __shared__ int colnz[MAXK];
    __shared__ int ind[MAXM];
    __shared__ int mt_s;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (threadIdx.x < MAXK) colnz[threadIdx.x] = 0;
    __syncthreads();
    // any(beta[,k] != 0) on the NEW beta (R/EM_impute.R:293)
    for (int e = threadIdx.x; e < d.G * K0; e += blockDim.x)
        if (a.beta[e] != 0.0) colnz[e % K0] = 1;
    if (warp == 0) {
        // loglik[it] = tot - priorB(previous iteration)  (Q5), then priorB <- sum of this M-step
        double pb = 0.0;
        for (int i = lane; i < a.ncta_gene; i += 32) pb += a.part_gene[i];
        pb = warp_sum(pb);
        if (lane == 0) {
            const double prev = a.priorB[0];
            *a.loglik_slot = sv.tot[0] - (a.compat_priorb_prev ? prev : 0.0);
            a.priorB[1] = prev;
            a.priorB[0] = pb;
            int mt = 0;
            for (int m = 0; m < M0; ++m)
                if (sv.nz[m] > K0 + 1 && sv.nz[m] > 3.0) ind[mt++] = m;      // :279
            mt_s = mt;
        }
        if (lane < M0)                                                        // :318
            a.pi[lane] = (sv.nz[lane] + a.pi_alpha - 1.0) / (double(d.n_total) + a.pi_alpha * M0 - M0);
    }
    __syncthreads();
    if (!a.do_lambda) return;
    const int mt = mt_s;
    if (mt > 1) {
        if (warp < K0) {
            const int k = warp;
            // lambda[-ind_m, ] <- 1/M0  (:288)
            if (lane < M0) {
                bool in = false;
                for (int q = 0; q < mt; ++q) in |= (ind[q] == lane);
                if (!in) a.lam[lane * K0 + k] = 1.0 / M0;
            }
            if (!colnz[k]) {
                if (lane < M0) a.lam[lane * K0 + k] = 1.0 / M0;               // :293-294
            } else {
                double E = 1.0, c = 1.0, x = 1.0;
                if (lane < mt) {
                    const int m = ind[lane];
                    E = sv.S[(size_t(m) * K0 + k) * K0 + k] + a.Mm[(size_t(m) * K0 + k) * K0 + k] * sv.nz[m];   // :282-285
                    c = sv.nz[m] - 2.0;
                    x = (E + 1.0) / (sv.nz[m] + 1.0);                         // :289
                }
                const double s = ordered_sum(lane < mt ? x : 0.0, mt);
                const double T = double(mt) / M0;
                x = x / s * mt / M0;                                          // :290
                lambda_solve_warp(E, c, x, mt, T, lane);
                if (lane < mt) a.lam[ind[lane] * K0 + k] = x;
            }
        }
    } else {
        for (int e = threadIdx.x; e < M0 * K0; e += blockDim.x) a.lam[e] = 1.0 / M0;   // :302
    }
}

constexpr int MG = 256;

// mu update, phase 1: partial sums of C_m = B' diag(1/s) B and rhs_m = B' (v / s) over the gene tiles gs, gs + GS, ...
// (s_g = nz_m + sigma_g / sigma0, v = Y z_m).  grid (M0, GS); partials are summed in a fixed order by k_mu_fin.
__global__ void __launch_bounds__(256) k_mu_acc(MstepArgs a, int GS) {
    extern __shared__ __align__(16) double sm[];
    const Dims& d = a.d;
    const int K0 = d.K0, M0 = d.M0, ldG = d.ldG;
    const StatView sv = stat_view(a.stats, d, a.general);
    const int m = blockIdx.x, gs = blockIdx.y;
    const int ucol = a.general ? K0 * M0 + m : K0 + m;
    const double nzm = sv.nz[m];
    double* sB = sm;               // [MG][K0]
    double* sw = sB + MG * K0;     // 1/s
    double* sv_ = sw + MG;         // v/s
    const int NE = K0 * K0 + K0;
    double acc[5] = {0, 0, 0, 0, 0};
    for (int g0 = gs * MG; g0 < ldG; g0 += GS * MG) {
        __syncthreads();
        for (int idx = threadIdx.x; idx < MG * K0; idx += blockDim.x)
            sB[idx] = (g0 + idx / K0 < ldG) ? a.beta[size_t(g0) * K0 + idx] : 0.0;      // rows past ldG contribute 0
        if (threadIdx.x < MG) {
            const int g = g0 + threadIdx.x;
            const bool v = g < ldG;
            const double s = nzm + (v ? a.sigma[size_t(g) * M0 + m] : 1.0) / a.sigma0;     // :308
            sw[threadIdx.x] = 1.0 / s;
            sv_[threadIdx.x] = v ? sv.C[size_t(g) * sv.ldc + ucol] / s : 0.0;
        }
        __syncthreads();
#pragma unroll
        for (int q = 0; q < 5; ++q) {
            const int e = threadIdx.x + q * 256;
            if (e >= NE) continue;
            double s = acc[q];
            if (e < K0 * K0) {
                const int j = e / K0, k = e - j * K0;
                for (int r = 0; r < MG; ++r) s = fma(sB[r * K0 + j] * sw[r], sB[r * K0 + k], s);   // b2' B
            } else {
                const int k = e - K0 * K0;
                for (int r = 0; r < MG; ++r) s = fma(sB[r * K0 + k], sv_[r], s);                   // b2' v
            }
            acc[q] = s;
        }
    }
#pragma unroll
    for (int q = 0; q < 5; ++q) {
        const int e = threadIdx.x + q * 256;
        if (e < NE) a.part_mu[(size_t(m) * GS + gs) * NE + e] = acc[q];
    }
}

// phase 2: per cluster the K0 x K0 solve (:311) and mu_gm = (v_g - b_g' x) / s_g for every gene (:313-315)
__global__ void __launch_bounds__(256) k_mu_fin(MstepArgs a, int GS) {
    __shared__ double sCm[MAXK * MAXK], srhs[MAXK];
    const Dims& d = a.d;
    const int K0 = d.K0, M0 = d.M0;
    const StatView sv = stat_view(a.stats, d, a.general);
    const int m = blockIdx.x;
    const int ucol = a.general ? K0 * M0 + m : K0 + m;
    const double nzm = sv.nz[m];
    const int NE = K0 * K0 + K0;
    for (int e = threadIdx.x; e < NE; e += blockDim.x) {
        double s = 0.0;
        for (int gs = 0; gs < GS; ++gs) s += a.part_mu[(size_t(m) * GS + gs) * NE + e];
        if (e < K0 * K0) {
            const int j = e / K0, k = e - j * K0;
            sCm[e] = s + ((j == k) ? a.sigma0 / a.lam[m * K0 + k] : 0.0);   // :311
        } else {
            srhs[e - K0 * K0] = s;
        }
    }
    __syncthreads();
    if (threadIdx.x < 32) {
        warp_cholesky(sCm, K0, threadIdx.x);
        warp_chol_solve(sCm, srhs, K0, threadIdx.x);
    }
    __syncthreads();
    for (int g = blockIdx.y * blockDim.x + threadIdx.x; g < d.G; g += gridDim.y * blockDim.x) {
        const double s = nzm + a.sigma[size_t(g) * M0 + m] / a.sigma0;
        double t = 0.0;
        for (int k = 0; k < K0; ++k) t += a.beta[size_t(g) * K0 + k] * srhs[k];
        a.mu[size_t(g) * M0 + m] = (sv.C[size_t(g) * sv.ldc + ucol] - t) / s;      // v/s - b2 x
    }
}

}  // namespace

int mstep_gene_ctas(int G) { return (G + GW - 1) / GW; }
int mu_gene_splits(int ldG) { const int t = (ldG + MG - 1) / MG; return t < 8 ? t : 8; }

void lasso_debug_fetch(int out[4], bool reset) {
    cudaMemcpyFromSymbol(out, g_lasso_dbg, sizeof(int) * 4);
    if (reset) {
        const int z[4] = {0, 0, 0, 0};
        cudaMemcpyToSymbol(g_lasso_dbg, z, sizeof z);
    }
}

void launch_mstep(const MstepArgs& a, cudaStream_t st) {
    const int K0 = a.d.K0, M0 = a.d.M0;
    {
        const size_t KK = size_t(K0) * K0;
        const size_t smem = sizeof(double) * (KK + size_t(M0) * K0 + GW + GW * KK * (a.general ? 2 : 1) + GW * K0) +
                            sizeof(int) * GW * K0;
        ensure_dyn_smem_k(k_gene_solve, smem);
        k_gene_solve<<<a.ncta_gene, GW * 32, smem, st>>>(a);
    }
    k_lmp<<<1, 1024, 0, st>>>(a);
    {
        const int GS = mu_gene_splits(a.d.ldG);
        const size_t smem = sizeof(double) * (size_t(MG) * K0 + 2 * MG);
        ensure_dyn_smem_k(k_mu_acc, smem);
        k_mu_acc<<<dim3(M0, GS), 256, smem, st>>>(a, GS);
        k_mu_fin<<<dim3(M0, (a.d.G + 2047) / 2048), 256, 0, st>>>(a, GS);
    }
}

void launch_lq_solve(const LqSolveArgs& a, cudaStream_t st) {
    const int K0 = a.d.K0, M0 = a.d.M0;
    const size_t KK = size_t(K0) * K0;
    const size_t smem = sizeof(double) * (KK + size_t(M0) * K0 + 2 * GW * KK + GW * K0) + sizeof(int) * GW * K0;
    ensure_dyn_smem_k(k_lq_solve, smem);
    k_lq_solve<<<(a.d.G + GW - 1) / GW, GW * 32, smem, st>>>(a);
}

}  // namespace sb
