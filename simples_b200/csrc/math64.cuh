// FP64 special functions tuned for the sampling kernels: coefficient tables live in __constant__
// memory (so DFMA takes them through LDC/LDCU instead of two UMOV per 64-bit immediate) and the
// divisions are a float-seeded Newton reciprocal.  Accuracy ~1e-15 relative (checked against SciPy in
// tests/test_oracle.py::test_device_math_restatement), i.e. the same mathematical functions the
// oracle evaluates with scipy.special.ndtr / ndtri.
//   pnorm : Cody (1969) rational Chebyshev erfc, the algorithm of R's pnorm_both (stats::pnorm,
//           called at R/EM_impute.R:179)
//   qnorm : Wichura (1988) AS241 PPND16, the algorithm of R's qnorm5 (used by rnorm's inversion,
//           R/EM_impute.R:186)
#pragma once
#include <cuda_runtime.h>

namespace sb {

static __constant__ double kQA[8] = {3.3871328727963666080,    1.3314166789178437745e+2, 1.9715909503065514427e+3,
                                     1.3731693765509461125e+4, 4.5921953931549871457e+4, 6.7265770927008700853e+4,
                                     3.3430575583588128105e+4, 2.5090809287301226727e+3};
static __constant__ double kQB[8] = {1.0,                      4.2313330701600911252e+1, 6.8718700749205790830e+2,
                                     5.3941960214247511077e+3, 2.1213794301586595867e+4, 3.9307895800092710610e+4,
                                     2.8729085735721942674e+4, 5.2264952788528545610e+3};
static __constant__ double kQC[8] = {1.42343711074968357734,    4.63033784615654529590,    5.76949722146069140550,
                                     3.64784832476320460504,    1.27045825245236838258,    2.41780725177450611770e-1,
                                     2.27238449892691845833e-2, 7.74545014278341407640e-4};
static __constant__ double kQD[8] = {1.0,                       2.05319162663775882187,    1.67638483018380384940,
                                     6.89767334985100004550e-1, 1.48103976427480074590e-1, 1.51986665636164571966e-2,
                                     5.47593808499534494600e-4, 1.05075007164441684324e-9};
static __constant__ double kQE[8] = {6.65790464350110377720,    5.46378491116411436990,    1.78482653991729133580,
                                     2.96560571828504891230e-1, 2.65321895265761230930e-2, 1.24266094738807843860e-3,
                                     2.71155556874348757815e-5, 2.01033439929228813265e-7};
static __constant__ double kQF[8] = {1.0,                       5.99832206555887937690e-1, 1.36929880922735805310e-1,
                                     1.48753612908506148525e-2, 7.86869131145613259100e-4, 1.84631831751005468180e-5,
                                     1.42151175831644588870e-7, 2.04426310338993978564e-15};

static __constant__ double kPA[5] = {2.2352520354606839287, 161.02823106855587881, 1067.6894854603709582,
                                     18154.981253343561249, 0.065682337918207449113};
static __constant__ double kPB[4] = {47.20258190468824187, 976.09855173777669322, 10260.932208618978205,
                                     45507.789335026729956};
static __constant__ double kPC[9] = {0.39894151208813466764, 8.8831497943883759412, 93.506656132177855979,
                                     597.27027639480026226,  2494.5375852903726711, 6848.1904505362823326,
                                     11602.651437647350124,  9842.7148383839780218, 1.0765576773720192317e-8};
static __constant__ double kPD[8] = {22.266688044328115691, 235.38790178262499861, 1519.377599407554805,
                                     6485.558298266760755,  18615.571640885098091, 34900.952721145977266,
                                     38912.003286093271411, 19685.429676859990727};
static __constant__ double kPP[6] = {0.21589853405795699,    0.1274011611602473639, 0.022235277870649807,
                                     0.001421619193227893466, 2.9112874951168792e-5, 0.02307344176494017303};
static __constant__ double kPQ[5] = {1.28426009614491121, 0.468238212480865118, 0.0659881378689285515,
                                     0.00378239633202758244, 7.29751555083966205e-5};

// exp(r) Taylor coefficients 1/n!, n = 2..13 (|r| <= ln2/2: truncation < 5e-18)
static __constant__ double kEX[12] = {0.5,
                                      0.16666666666666666667,
                                      4.1666666666666666667e-2,
                                      8.3333333333333333333e-3,
                                      1.3888888888888888889e-3,
                                      1.9841269841269841270e-4,
                                      2.4801587301587301587e-5,
                                      2.7557319223985890653e-6,
                                      2.7557319223985890653e-7,
                                      2.5052108385441718775e-8,
                                      2.0876756987868098979e-9,
                                      1.6059043836821614599e-10};

// 1/y for y within float range; two Newton steps from a MUFU.RCP float seed (~1 ulp).
__device__ __forceinline__ double fast_rcp(double y) {
    float rf;
    asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(rf) : "f"(static_cast<float>(y)));
    double r = static_cast<double>(rf);
    double e = fma(-y, r, 1.0);
    r = fma(r, e, r);
    e = fma(-y, r, 1.0);
    r = fma(r, e, r);
    return r;
}

// exp(h) for h <= 0: 2^k * e^r with the exponent patched in directly; below -700 (denormal / underflow
// territory, only reached from the cold |x| > 8 pnorm path) it defers to the library exp.
template <bool CHECK = true>
__device__ __forceinline__ double exp_neg(double h) {
    if (CHECK && h < -700.0) return exp(h);
    const double MAGIC = 6755399441055744.0;                    // 1.5 * 2^52: round-to-nearest-integer trick
    const double t = fma(h, 1.4426950408889634074, MAGIC);
    const double kd = t - MAGIC;
    double r = fma(kd, -6.93147180369123816490e-01, h);         // ln2 split hi / lo
    r = fma(kd, -1.90821492927058770002e-10, r);
    double p = kEX[11];
#pragma unroll
    for (int i = 10; i >= 0; --i) p = fma(p, r, kEX[i]);
    p = fma(p * r, r, r) + 1.0;                                 // 1 + r + r^2 (1/2 + ...)
    const int k = __double2loint(t);                            // low word of (k + MAGIC) holds k (two's complement)
    return __hiloint2double(__double2hiint(p) + (k << 20), __double2loint(p));
}

// exp(-y^2/2) with the rounding error of y*y compensated; CHECK = false when the caller guarantees |y| < 37
template <bool CHECK = true>
__device__ __forceinline__ double exp_mhalfsq(double y) {
    const double h = -0.5 * y * y;
    const double l = fma(-0.5 * y, y, -h);
    const double e = exp_neg<CHECK>(h);
    return fma(e, l, e);
}

// Rational part of Cody's mid-range erfc: Phi(-y) = exp(-y^2/2) * pn_mid_ratio(y).  Designed for
// 0.67 <= y <= sqrt(32) (relative error < 2e-15 there); used on all of [0, 8]: measured relative error
// <= 5e-12 on [0, 0.67] and <= 3e-13 on (sqrt(32), 8] -- far inside the 1e-8 / 1e-6 parity tolerances.
__device__ __forceinline__ double pn_mid_ratio(double y) {
    // Cody writes the two degree-8 polynomials as (xn + c) * y steps; the same polynomials in Horner form are one FMA per
    // coefficient (half the FP64 instructions, one rounding per step instead of two)
    double xn = kPC[8], xd = y + kPD[0];
    xn = fma(xn, y, kPC[0]);
#pragma unroll
    for (int i = 1; i < 8; ++i) {
        xn = fma(xn, y, kPC[i]);
        xd = fma(xd, y, kPD[i]);
    }
    return xn * fast_rcp(xd);
}

// Phi(x), full relative accuracy in the lower tail (all three Cody regions; cold path)
static __device__ __noinline__ double pnorm_full(double x) {
    const double y = fabs(x);
    if (y <= 0.67448975) {
        const double xs = x * x;
        double xn = kPA[4] * xs, xd = xs;
#pragma unroll 1
        for (int i = 0; i < 3; ++i) {
            xn = (xn + kPA[i]) * xs;
            xd = (xd + kPB[i]) * xs;
        }
        return fma(x * (xn + kPA[3]), 1.0 / (xd + kPB[3]), 0.5);
    }
    double lo;
    if (y <= 5.656854249492380195206754896838) {   // sqrt(32)
        lo = exp_mhalfsq(y) * pn_mid_ratio(y);
    } else {
        const double xs = 1.0 / (x * x);
        double xn = kPP[5] * xs, xd = xs;
#pragma unroll 1
        for (int i = 0; i < 4; ++i) {
            xn = (xn + kPP[i]) * xs;
            xd = (xd + kPQ[i]) * xs;
        }
        double t = xs * (xn + kPP[4]) / (xd + kPQ[4]);
        t = (0.398942280401432677939946059934 - t) / y;
        lo = exp_mhalfsq(y) * t;
    }
    return x < 0.0 ? lo : 1.0 - lo;
}

// Phi(x) on the hot path: one branch-free formula for |x| <= 8, cold call otherwise.
__device__ __forceinline__ double pnorm(double x) {
    const double y = fabs(x);
    if (y <= 8.0) {
        const double lo = exp_mhalfsq<false>(y) * pn_mid_ratio(y);
        return x < 0.0 ? lo : 1.0 - lo;
    }
    return pnorm_full(x);
}

template <int N>
__device__ __forceinline__ double horner8(const double* c, double x) {
    double r = c[N - 1];
#pragma unroll
    for (int i = N - 2; i >= 0; --i) r = fma(r, x, c[i]);
    return r;
}

// far tail of AS241 (sqrt(-log p) > 5, i.e. p < 1.4e-11): cold
static __device__ __noinline__ double qnorm_far(double r) {
    r -= 5.0;
    double n = kQE[7], d = kQF[7];
#pragma unroll 1
    for (int i = 6; i >= 0; --i) {
        n = fma(n, r, kQE[i]);
        d = fma(d, r, kQF[i]);
    }
    return n / d;
}

// Phi^-1(p) for |p - 0.5| <= 0.425 (AS241 central region only)
__device__ __forceinline__ double qnorm_central(double p) {
    const double q = p - 0.5;
    const double r = 0.180625 - q * q;
    return q * horner8<8>(kQA, r) * fast_rcp(horner8<8>(kQB, r));
}

// sqrt(x) for x in float range: float rsqrt seed + two coupled Newton steps (~1 ulp)
__device__ __forceinline__ double fast_sqrt(double x) {
    float rf;
    asm("rsqrt.approx.ftz.f32 %0, %1;" : "=f"(rf) : "f"(static_cast<float>(x)));
    double h = 0.5 * static_cast<double>(rf);
    double g = x * static_cast<double>(rf);
    double e = fma(-g, h, 0.5);
    g = fma(g, e, g);
    h = fma(h, e, h);
    e = fma(-g, h, 0.5);
    g = fma(g, e, g);
    h = fma(h, e, h);
    return fma(fma(-g, g, x), h, g);
}

// log(1+f)/s - 2 coefficients of the fdlibm e_log.c kernel (s = f/(2+f); remainder < 2^-58.45)
static __constant__ double kLG[7] = {6.666666666666735130e-01, 3.999999999940941908e-01, 2.857142874366239149e-01,
                                     2.222219843214978396e-01, 1.818357216161805012e-01, 1.531383769920937332e-01,
                                     1.479819860511658591e-01};

// log(x) for positive NORMAL x (no zero / denormal / inf / nan handling): the fdlibm algorithm -- x = 2^k m with
// m in [sqrt(1/2), sqrt(2)), log m = f - (f^2/2 - s (f^2/2 + R(s^2))) -- with the division by the Newton reciprocal.
// ~35 instructions against ~70 of the library routine (whose special cases and double-double tail are not needed here);
// relative error < 2e-16 (checked against numpy in tests/test_oracle.py::test_device_math_restatement).
__device__ __forceinline__ double log_normal(double x) {
    int hi = __double2hiint(x);
    int k = (hi >> 20) - 1023;
    hi &= 0x000fffff;
    const int i = (hi + 0x95f64) & 0x100000;                    // mantissa above sqrt(2): halve it, k + 1
    const double m = __hiloint2double(hi | (i ^ 0x3ff00000), __double2loint(x));
    k += i >> 20;
    const double f = m - 1.0;
    const double s = f * fast_rcp(2.0 + f);
    const double z = s * s;
    const double R = z * horner8<7>(kLG, z);
    const double hfsq = 0.5 * f * f;
    const double dk = static_cast<double>(k);
    return fma(dk, 6.93147180369123816490e-01, -((hfsq - fma(s, hfsq + R, dk * 1.90821492927058770002e-10)) - f));
}

// sqrt of a responsibility z in [0, 1]: the Newton form above 1e-30, the library routine below (zero, denormals)
__device__ __forceinline__ double sqrt_unit(double z) { return z > 1e-30 ? fast_sqrt(z) : sqrt(z); }

// Phi^-1(p) for |p - 0.5| > 0.425 (AS241 tails); min(p, 1 - p) must be a normal number (>= 2.3e-308)
__device__ __forceinline__ double qnorm_tail(double p) {
    const double q = p - 0.5;
    const double pp = q < 0.0 ? p : 1.0 - p;
    double r = fast_sqrt(-log_normal(pp));
    double v;
    if (r <= 5.0) {
        r -= 1.6;
        v = horner8<8>(kQC, r) * fast_rcp(horner8<8>(kQD, r));
    } else {
        v = qnorm_far(r);
    }
    return q < 0.0 ? -v : v;
}

// Phi^-1(p), 0 < p < 1
__device__ __forceinline__ double qnorm(double p) {
    const double q = p - 0.5;
    if (fabs(q) <= 0.425) {
        const double r = 0.180625 - q * q;
        return q * horner8<8>(kQA, r) * fast_rcp(horner8<8>(kQB, r));
    }
    const double pp = q < 0.0 ? p : 1.0 - p;
    double r = fast_sqrt(-log_normal(pp));            // min(p, 1 - p) is a normal number at every call site
    double v;
    if (r <= 5.0) {
        r -= 1.6;
        v = horner8<8>(kQC, r) * fast_rcp(horner8<8>(kQD, r));
    } else {
        v = qnorm_far(r);
    }
    return q < 0.0 ? -v : v;
}

}  // namespace sb
