// Counter-based variate tape (Philox4x32-10) and the inverse-CDF samplers that
// replace the reference's sequential R RNG calls (rmultinom / rmvnorm / rbinom /
// rnorm / msm::rtnorm at R/EM_impute.R:164-193, R/do_impute.R:107-131).
// Keyed by (seed; sweep, GLOBAL cell index, gene) so results do not depend on
// the number of GPUs.  Mirrored bit-for-bit by the test oracle.
#pragma once
#include <cstdint>
#include <cuda_runtime.h>

#include "math64.cuh"

namespace sb {

struct Philox4 {
    uint32_t x0, x1, x2, x3;
};

__device__ __forceinline__ Philox4 philox4x32_10(uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3, uint32_t k0,
                                                 uint32_t k1) {
    constexpr uint32_t M0 = 0xD2511F53u, M1 = 0xCD9E8D57u, W0 = 0x9E3779B9u, W1 = 0xBB67AE85u;
#pragma unroll
    for (int r = 0; r < 10; ++r) {
        const uint32_t hi0 = __umulhi(M0, c0), lo0 = M0 * c0;
        const uint32_t hi1 = __umulhi(M1, c2), lo1 = M1 * c2;
        const uint32_t n0 = hi1 ^ c1 ^ k0, n2 = hi0 ^ c3 ^ k1;
        c0 = n0;
        c1 = lo1;
        c2 = n2;
        c3 = lo0;
        k0 += W0;
        k1 += W1;
    }
    return Philox4{c0, c1, c2, c3};
}

// Philox4x32-10 with the round keys k_r = k_0 + r W precomputed (they are the same for every entry of a launch):
// kernels take them as parameters, so each round is two IMAD.WIDE + two LOP3 with a constant-bank operand.
// Bit-identical to philox4x32_10.
__device__ __forceinline__ Philox4 philox4x32_10_keys(uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3, const uint32_t (&rk)[20]) {
#pragma unroll
    for (int r = 0; r < 10; ++r) {
        const unsigned long long p0 = (unsigned long long)0xD2511F53u * c0;
        const unsigned long long p1 = (unsigned long long)0xCD9E8D57u * c2;
        const uint32_t n0 = uint32_t(p1 >> 32) ^ c1 ^ rk[2 * r], n2 = uint32_t(p0 >> 32) ^ c3 ^ rk[2 * r + 1];
        c0 = n0;
        c1 = uint32_t(p1);
        c2 = n2;
        c3 = uint32_t(p0);
    }
    return Philox4{c0, c1, c2, c3};
}
inline void philox_round_keys(uint64_t seed, uint32_t (&rk)[20]) {
    for (int r = 0; r < 10; ++r) {
        rk[2 * r] = uint32_t(seed) + uint32_t(r) * 0x9E3779B9u;
        rk[2 * r + 1] = uint32_t(seed >> 32) + uint32_t(r) * 0xBB67AE85u;
    }
}

// ((x >> 12) + 0.5) * 2^-52  in (0,1)
// (== [1 + m 2^-52] - 1 + 2^-53 with m = x >> 12 the top 52 bits; every step is exact)
__device__ __forceinline__ double u52(uint32_t hi, uint32_t lo) {
    const double one_m = __hiloint2double(0x3FF00000 | (hi >> 12), (hi << 20) | (lo >> 12));
    return (one_m - 1.0) + 1.1102230246251565e-16;
}

// per-entry Philox block: tag 0, counter (gene, cell_lo, sweep, cell_hi << 8); (x0,x1) -> ua, (x2,x3) -> ub
__device__ __forceinline__ Philox4 entry_philox(uint64_t seed, uint32_t sweep, uint64_t cell, uint32_t gene) {
    return philox4x32_10(gene, static_cast<uint32_t>(cell), sweep, static_cast<uint32_t>(cell >> 32) << 8,
                         static_cast<uint32_t>(seed), static_cast<uint32_t>(seed >> 32));
}

// per-entry uniforms: tag 0, counter (gene, cell_lo, sweep, cell_hi << 8)
__device__ __forceinline__ void entry_uniforms(uint64_t seed, uint32_t sweep, uint64_t cell, uint32_t gene, double& ua,
                                               double& ub) {
    const Philox4 p = philox4x32_10(gene, static_cast<uint32_t>(cell), sweep, static_cast<uint32_t>(cell >> 32) << 8,
                                    static_cast<uint32_t>(seed), static_cast<uint32_t>(seed >> 32));
    ua = u52(p.x0, p.x1);
    ub = u52(p.x2, p.x3);
}

// per-cell uniforms: tag 1, block j -> uniforms 2j, 2j+1
__device__ __forceinline__ void cell_uniform_pair(uint64_t seed, uint32_t sweep, uint64_t cell, uint32_t j, double& u0,
                                                  double& u1) {
    const Philox4 p = philox4x32_10(j, static_cast<uint32_t>(cell), sweep,
                                    (static_cast<uint32_t>(cell >> 32) << 8) | 1u, static_cast<uint32_t>(seed),
                                    static_cast<uint32_t>(seed >> 32));
    u0 = u52(p.x0, p.x1);
    u1 = u52(p.x2, p.x3);
}

// deep lower tail (r < -30): solve log Phi(z) = log u + log Phi(r) by Newton (erfcx form); never inlined
static __device__ __noinline__ double trunc_z_deep(double u, double r) {
    const double S2 = 1.4142135623730951, S2PI = 2.5066282746310002;
    const double lu = log(u);
    const double t = lu + log(0.5 * erfcx(-r / S2)) - 0.5 * r * r;
    double z = -sqrt(r * r - 2.0 * lu);
#pragma unroll 1
    for (int it = 0; it < 4; ++it) {
        const double e = 0.5 * erfcx(-z / S2);
        z = z - (log(e) - 0.5 * z * z - t) * e * S2PI;
    }
    return fmin(z, r);
}

// One zero entry: R/EM_impute.R:175-198.  ms/sds include gene_sd and gene_mean; isd = 1/sds.
//   prob = pnorm(cutoff, ms, sds); prob_drop = (1-p)/(prob p + 1-p); I_drop ~ Bern(prob_drop)   (:179-181)
//   dropout: N(ms, sds) (:186);  else TN(ms, sds; upper = cutoff) (:191-192), both by inverse CDF.
// The Bernoulli test ua < prob_drop is evaluated as ua (prob p + 1-p) < 1-p (no division).  When the
// cutoff is > 8 sd below the mean, prob < 6.3e-16 and the draw is a certain dropout unless ua is within
// 1e-9 of 1 or p > 0.999 -- those cases take the exact (cold) path, so the decision is always exact.
static __device__ __noinline__ double sample_entry(double ms, double sds, double isd, double p, double cutoff, double ua,
                                               double ub) {
    const double r = (cutoff - ms) * isd;
    const double q1 = 1.0 - p;
    double prob;
    if (r < -8.0 && q1 >= 1e-3 && ua <= 0.999999999) prob = 0.0;
    else prob = pnorm(r);
    const bool drop = ua * fma(prob, p, q1) < q1;
    double zz;
    if (!drop && r < -30.0) {
        zz = trunc_z_deep(ub, r);
    } else {
        zz = qnorm(drop ? ub : ub * prob);
        if (!drop) zz = fmin(zz, r);
    }
    return fma(sds, zz, ms);
}


// SIMPLES_MODE_EXPECT: E[x] = ms - sds p phi(r) / (p Phi(r) + 1 - p), r = (cutoff - ms) / sds (the Mills ratio of the
// textbook form cancels against 1 - prob_drop, so the tail needs no special case).
static __device__ __noinline__ double expect_entry(double ms, double sds, double isd, double p, double cutoff) {
    const double r = (cutoff - ms) * isd;
    const double phi = exp_mhalfsq(fabs(r)) * 0.398942280401432677939946059934;
    return ms - sds * p * phi / fma(p, pnorm(r), 1.0 - p);
}

// x = ms + sds Phi^-1(|v|) for an entry whose quantile argument is already known (k_sweep, dense blocks);
// v < 0 marks a truncated draw, clamped to the cutoff (R/EM_impute.R:191-192).
static __device__ __noinline__ double quantile_draw(double ms, double sds, double cutoff, double v) {
    double x = fma(sds, qnorm(fabs(v)), ms);
    if (v < 0.0) x = fmin(x, cutoff);
    return x;
}

}  // namespace sb
