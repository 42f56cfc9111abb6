// cluster_prep: per-cluster constants of an E-pass (R/EM_impute.R:123-125 = R/do_impute.R:70-72):
//   A_m = B' Sigma_m^-1 B + Lambda_m^-1,  M_m = A_m^-1,  dt_m = sum log sigma + sum log lambda - log det M_m,
//   Mh_m = M_m^(1/2) (symmetric), off_m = mu_m'(B/sigma_m), offq_m = mu_m'(B/sigma_q), cmu_m = sum mu^2/sigma,
// plus the gene-major GEMM operands Q (projection), isg, Qs (sweep).
#include "linalg.cuh"
#include "ptx.cuh"
#include "state.h"

namespace sb {
namespace {

__global__ void k_build_ops(PrepArgs a) {
    const Dims& d = a.d;
    const int g = blockIdx.x * blockDim.x + threadIdx.x;
    if (g >= d.ldG) return;
    const int K0 = d.K0, M0 = d.M0;
    const double sd = a.gsd[g], mean = a.gmean[g];
    for (int s = 0; s < d.NS; ++s) {
        const double inv = 1.0 / a.sigma[size_t(g) * M0 + s];
        a.isg[size_t(s) * d.ldG + g] = inv;
        double* q = a.Q + (size_t(s) * d.ldG + g) * d.QS;
        for (int k = 0; k < K0; ++k) q[k] = a.beta[size_t(g) * K0 + k] * inv;
        for (int m = 0; m < M0; ++m) q[K0 + m] = a.mu[size_t(g) * M0 + m] * inv;
        for (int c = K0 + M0; c < d.QS; ++c) q[c] = 0.0;
    }
    for (int m = 0; m < M0; ++m) {
        const double sdv = sqrt(a.sigma[size_t(g) * M0 + m]) * sd;
        a.sdsA[size_t(g) * M0 + m] = sdv;
        a.isdA[size_t(g) * M0 + m] = 1.0 / sdv;
        a.musdA[size_t(g) * M0 + m] = a.mu[size_t(g) * M0 + m] * sd;
    }
    a.igs[g] = 1.0 / sd;
    if (a.BM) {            // fused MC pass operands (cluster-shared sigma: column 0)
        double* bmr = a.BM + size_t(g) * d.QS;
        for (int k = 0; k < K0; ++k) bmr[k] = a.beta[size_t(g) * K0 + k];
        for (int m = 0; m < M0; ++m) bmr[K0 + m] = a.mu[size_t(g) * M0 + m];
        for (int c = K0 + M0; c < d.QS; ++c) bmr[c] = 0.0;
        const double sdv = sqrt(a.sigma[size_t(g) * M0]) * sd;
        a.gv[size_t(g) * 4 + 0] = 1.0 / sdv;
        a.gv[size_t(g) * 4 + 1] = sdv;
        a.gv[size_t(g) * 4 + 2] = mean;
        a.gv[size_t(g) * 4 + 3] = 1.0 / sd;
    }
    double* qs = a.Qs + size_t(g) * d.FS;
    for (int k = 0; k < K0; ++k) qs[k] = a.beta[size_t(g) * K0 + k] * sd;
    for (int c = K0; c < d.FS; ++c) qs[c] = 0.0;
}

constexpr int GT = 256;  // genes per smem tile

__global__ void __launch_bounds__(256) k_cluster(PrepArgs a) {
    extern __shared__ __align__(16) double sm[];
    const Dims& d = a.d;
    const int K0 = d.K0, M0 = d.M0, ldG = d.ldG;
    const int m = blockIdx.x;
    const int cs = (d.NS == 1) ? 0 : m;
    const int qcol = (d.NS == 1) ? 0 : M0 - 1;    // sigma column of the stale beta2 (Q1)
    double* sB = sm;                 // [GT][K0]
    double* sw = sB + GT * K0;       // [GT] 1/sigma_m
    double* smw = sw + GT;           // mu/sigma_m
    double* smq = smw + GT;          // mu/sigma_q
    double* sm2 = smq + GT;          // mu^2/sigma_m
    double* sls = sm2 + GT;          // log sigma_m
    double* sA = sls + GT;           // [K0][K0]
    double* sLi = sA + K0 * K0;      // [K0][K0]
    double* sV = sLi + K0 * K0;      // [K0][K0]
    double* sM = sV + K0 * K0;       // [K0][K0]
    double* sRed = sM + K0 * K0;     // [2]

    const int NE = K0 * K0 + 2 * K0 + 2;
    double acc[5] = {0, 0, 0, 0, 0};   // NE <= 1090 <= 5 * 256
    for (int g0 = 0; g0 < ldG; g0 += GT) {
        __syncthreads();
        for (int idx = threadIdx.x; idx < GT * K0; idx += blockDim.x)
            sB[idx] = (g0 + idx / K0 < ldG) ? a.beta[size_t(g0) * K0 + idx] : 0.0;      // rows past ldG contribute 0
        if (threadIdx.x < GT) {
            const int g = g0 + threadIdx.x;
            const bool v = g < ldG;
            const double sg = v ? a.sigma[size_t(g) * M0 + cs] : 1.0, sq = v ? a.sigma[size_t(g) * M0 + qcol] : 1.0;
            const double mu = v ? a.mu[size_t(g) * M0 + m] : 0.0;
            sw[threadIdx.x] = 1.0 / sg;
            smw[threadIdx.x] = mu / sg;
            smq[threadIdx.x] = mu / sq;
            sm2[threadIdx.x] = mu * mu / sg;
            sls[threadIdx.x] = log(sg);
        }
        __syncthreads();
#pragma unroll
        for (int q = 0; q < 5; ++q) {
            const int e = threadIdx.x + q * 256;
            if (e >= NE) continue;
            double s = acc[q];
            if (e < K0 * K0) {
                const int j = e / K0, k = e - j * K0;
                for (int r = 0; r < GT; ++r) s = fma(sB[r * K0 + j] * sw[r], sB[r * K0 + k], s);
            } else if (e < K0 * K0 + K0) {
                const int k = e - K0 * K0;
                for (int r = 0; r < GT; ++r) s = fma(smw[r], sB[r * K0 + k], s);
            } else if (e < K0 * K0 + 2 * K0) {
                const int k = e - K0 * K0 - K0;
                for (int r = 0; r < GT; ++r) s = fma(smq[r], sB[r * K0 + k], s);
            } else if (e == K0 * K0 + 2 * K0) {
                for (int r = 0; r < GT; ++r) s += sm2[r];
            } else {
                for (int r = 0; r < GT; ++r) s += sls[r];
            }
            acc[q] = s;
        }
    }
    __syncthreads();
#pragma unroll
    for (int q = 0; q < 5; ++q) {
        const int e = threadIdx.x + q * 256;
        if (e >= NE) continue;
        if (e < K0 * K0) {
            const int j = e / K0, k = e - j * K0;
            sA[e] = acc[q] + ((j == k) ? 1.0 / a.lam[m * K0 + k] : 0.0);
        } else if (e < K0 * K0 + K0) {
            a.cc.off[m * K0 + (e - K0 * K0)] = acc[q];
        } else if (e < K0 * K0 + 2 * K0) {
            a.cc.offq[m * K0 + (e - K0 * K0 - K0)] = acc[q];
        } else if (e == K0 * K0 + 2 * K0) {
            a.cc.cmu[m] = acc[q];
        } else {
            sRed[0] = acc[q];
        }
    }
    __syncthreads();
    if (threadIdx.x < 32) {
        const int lane = threadIdx.x;
        warp_cholesky(sA, K0, lane);
        double ld = 0.0, ll = 0.0;
        for (int k = 0; k < K0; ++k) {
            ld += log(sA[k * K0 + k]);
            ll += log(a.lam[m * K0 + k]);
        }
        // dt = sum log sigma + sum log lambda - log det M,  log det M = -2 sum log L_kk
        if (lane == 0) {
            a.cc.dt[m] = sRed[0] + ll + 2.0 * ld;
            a.cc.logpi[m] = log(a.pi[m]);
        }
        warp_tri_inverse(sA, sLi, K0, lane);
        warp_ltl(sLi, sM, K0, lane);
        for (int e = lane; e < K0 * K0; e += 32) {
            a.cc.Mm[size_t(m) * K0 * K0 + e] = sM[e];
            sA[e] = sM[e];
        }
        {
            const int KR = 4 * ((K0 + 3) / 4), KPS = 8 * ((K0 + 7) / 8) + 4;
            for (int e = lane; e < KR * KPS; e += 32) {
                const int k = e / KPS, c = e - k * KPS;
                a.cc.MmP[size_t(m) * KR * KPS + e] = (k < K0 && c < K0) ? sM[k * K0 + c] : 0.0;
            }
        }
        __syncwarp();
        warp_sym_sqrt(sA, sV, sLi, K0, lane);
        for (int e = lane; e < K0 * K0; e += 32) a.cc.Mh[size_t(m) * K0 * K0 + e] = sLi[e];
    }
}

}  // namespace

void launch_prep(const PrepArgs& a, cudaStream_t st, cudaStream_t st_cluster) {
    // the GEMM operands (needed by estep_project / impute_sweep) and the per-cluster constants (needed only by
    // cell_post) are independent: the latter may run on a side stream and overlap the first projection.
    k_build_ops<<<(a.d.ldG + 127) / 128, 128, 0, st>>>(a);
    const int K0 = a.d.K0;
    const size_t smem = sizeof(double) * (size_t(GT) * K0 + 5 * GT + 4 * K0 * K0 + 2);
    ensure_dyn_smem_k(k_cluster, smem);
    k_cluster<<<a.d.M0, 256, smem, st_cluster>>>(a);
}

}  // namespace sb
