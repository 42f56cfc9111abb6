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
        a.musdA[size_t(g) * M0 + m] = fma(a.mu[size_t(g) * M0 + m], sd, mean);
    }
    a.gmig[size_t(g) * 2 + 0] = mean;
    a.gmig[size_t(g) * 2 + 1] = 1.0 / sd;
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

constexpr int GT = 64;   // genes per CTA of k_gram

// Layout of one partial (and of the reduced vector): [G2 (NS K0 K0) | off (M0 K0) | offq (M0 K0) | cmu (M0) | lsig (NS)]
__host__ __device__ inline int prep_ne(int K0, int M0, int NS) { return NS * K0 * K0 + 2 * M0 * K0 + M0 + NS; }

// Gene-parallel part of cluster_prep: every CTA sums its GT genes' contributions to
//   G2_s = B' diag(1/sigma_s) B,  off_m = mu_m'(B/sigma_cs(m)),  offq_m = mu_m'(B/sigma_q),  cmu_m = sum mu^2/sigma,
//   lsig_s = sum log sigma_s.
// With cluster-shared sigma (NS = 1) the Gram matrix is the SAME for every cluster and is formed once -- round 1 had each
// of the M0 cluster CTAs walk all genes in eight barrier-separated tiles (0.28 ms, exposed in strongly scaled runs).
__global__ void __launch_bounds__(128) k_gram(PrepArgs a, double* part) {
    extern __shared__ __align__(16) double sm[];
    const Dims& d = a.d;
    const int K0 = d.K0, M0 = d.M0, NS = d.NS, ldG = d.ldG;
    const int qcol = (NS == 1) ? 0 : M0 - 1;        // sigma column of the stale beta2 (Q1)
    double* sB = sm;                  // [GT][K0]
    double* sI = sB + GT * K0;        // [GT][NS]  1/sigma_s
    double* sMu = sI + GT * NS;       // [GT][M0]
    double* sL = sMu + GT * M0;       // [GT][NS]  log sigma_s
    const int g0 = blockIdx.x * GT;
    for (int idx = threadIdx.x; idx < GT * K0; idx += blockDim.x) {
        const int r = idx / K0;
        sB[idx] = (g0 + r < ldG) ? a.beta[size_t(g0) * K0 + idx] : 0.0;
    }
    for (int idx = threadIdx.x; idx < GT * M0; idx += blockDim.x) {
        const int r = idx / M0;
        sMu[idx] = (g0 + r < ldG) ? a.mu[size_t(g0) * M0 + idx] : 0.0;
    }
    for (int idx = threadIdx.x; idx < GT * NS; idx += blockDim.x) {
        const int r = idx / NS, sc = idx - r * NS;
        const double sg = (g0 + r < ldG) ? a.sigma[size_t(g0 + r) * M0 + sc] : 1.0;
        sI[idx] = 1.0 / sg;
        sL[idx] = log(sg);
    }
    __syncthreads();
    const int nG2 = NS * K0 * K0, nOff = M0 * K0, NE = prep_ne(K0, M0, NS);
    double* out = part + size_t(blockIdx.x) * NE;
    for (int e = threadIdx.x; e < NE; e += blockDim.x) {
        double s = 0.0;
        if (e < nG2) {
            const int sc = e / (K0 * K0), jk = e - sc * K0 * K0, j = jk / K0, k = jk - j * K0;
            for (int r = 0; r < GT; ++r) s = fma(sB[r * K0 + j] * sI[r * NS + sc], sB[r * K0 + k], s);
        } else if (e < nG2 + 2 * nOff) {
            const int f = e - nG2, which = f / nOff, mk = f - which * nOff, m = mk / K0, k = mk - m * K0;
            const int sc = which ? qcol : ((NS == 1) ? 0 : m);
            for (int r = 0; r < GT; ++r) s = fma(sMu[r * M0 + m] * sI[r * NS + sc], sB[r * K0 + k], s);
        } else if (e < nG2 + 2 * nOff + M0) {
            const int m = e - nG2 - 2 * nOff, sc = (NS == 1) ? 0 : m;
            for (int r = 0; r < GT; ++r) s = fma(sMu[r * M0 + m] * sMu[r * M0 + m], sI[r * NS + sc], s);
        } else {
            const int sc = e - nG2 - 2 * nOff - M0;
            for (int r = 0; r < GT; ++r) s += sL[r * NS + sc];
        }
        out[e] = s;
    }
}

// Per-cluster part: fixed-order sum of the k_gram partials, then the K0 x K0 linear algebra by one warp.
__global__ void __launch_bounds__(128) k_cluster(PrepArgs a, const double* part, int nparts) {
    extern __shared__ __align__(16) double sm[];
    const Dims& d = a.d;
    const int K0 = d.K0, M0 = d.M0, NS = d.NS;
    const int m = blockIdx.x;
    const int cs = (NS == 1) ? 0 : m;
    double* sA = sm;                 // [K0][K0]
    double* sLi = sA + K0 * K0;      // [K0][K0]
    double* sV = sLi + K0 * K0;      // [K0][K0]
    double* sM = sV + K0 * K0;       // [K0][K0]
    double* sRed = sM + K0 * K0;     // [2]
    double* sW = sRed + 2;           // [32] rotation parameters of one Jacobi round
    const int nG2 = NS * K0 * K0, nOff = M0 * K0, NE = prep_ne(K0, M0, NS);
    auto total = [&](int e) {          // fixed order, four independent chains so that the L2 loads pipeline
        double s0 = 0.0, s1 = 0.0, s2 = 0.0, s3 = 0.0;
        int p = 0;
        for (; p + 4 <= nparts; p += 4) {
            s0 += part[size_t(p) * NE + e];
            s1 += part[size_t(p + 1) * NE + e];
            s2 += part[size_t(p + 2) * NE + e];
            s3 += part[size_t(p + 3) * NE + e];
        }
        for (; p < nparts; ++p) s0 += part[size_t(p) * NE + e];
        return (s0 + s1) + (s2 + s3);
    };
    for (int e = threadIdx.x; e < K0 * K0 + 2 * K0 + 2; e += blockDim.x) {
        if (e < K0 * K0) {
            const int j = e / K0, k = e - j * K0;
            sA[e] = total(cs * K0 * K0 + e) + ((j == k) ? 1.0 / a.lam[m * K0 + k] : 0.0);
        } else if (e < K0 * K0 + K0) {
            const int k = e - K0 * K0;
            a.cc.off[m * K0 + k] = total(nG2 + m * K0 + k);
        } else if (e < K0 * K0 + 2 * K0) {
            const int k = e - K0 * K0 - K0;
            a.cc.offq[m * K0 + k] = total(nG2 + nOff + m * K0 + k);
        } else if (e == K0 * K0 + 2 * K0) {
            a.cc.cmu[m] = total(nG2 + 2 * nOff + m);
        } else {
            sRed[0] = total(nG2 + 2 * nOff + M0 + cs);
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
        // warm start from the previous call's eigenvectors (M_m moves little between EM iterations); a cold start every
        // 8th call bounds the loss of orthogonality that accumulates through the rotations
        const int jst = a.cc.jst ? a.cc.jst[m] : 0;
        const bool warm = jst > 0 && jst < 8;
        warp_sym_sqrt(sA, sV, sLi, sW, K0, lane, warm ? a.cc.Vj + size_t(m) * K0 * K0 : nullptr);
        for (int e = lane; e < K0 * K0; e += 32) a.cc.Mh[size_t(m) * K0 * K0 + e] = sLi[e];
        if (a.cc.jst) {
            for (int e = lane; e < K0 * K0; e += 32) a.cc.Vj[size_t(m) * K0 * K0 + e] = sV[e];
            if (lane == 0) a.cc.jst[m] = warm ? jst + 1 : 1;
        }
    }
}

}  // namespace

int prep_parts(int ldG) { return (ldG + GT - 1) / GT; }
size_t prep_scratch_doubles(const Dims& d, int NS_max) { return size_t(prep_parts(d.ldG)) * prep_ne(d.K0, d.M0, NS_max); }

void launch_prep(const PrepArgs& a, cudaStream_t st, cudaStream_t st_cluster) {
    // the GEMM operands (needed by estep_project / impute_sweep) and the per-cluster constants (needed only by
    // cell_post) are independent: the latter may run on a side stream and overlap the first projection.
    k_build_ops<<<(a.d.ldG + 127) / 128, 128, 0, st>>>(a);
    const int K0 = a.d.K0, M0 = a.d.M0, NS = a.d.NS;
    const int nparts = prep_parts(a.d.ldG);
    {
        const size_t smem = sizeof(double) * size_t(GT) * (K0 + 2 * NS + M0);
        ensure_dyn_smem_k(k_gram, smem);
        k_gram<<<nparts, 128, smem, st_cluster>>>(a, a.part_prep);
    }
    const size_t smem = sizeof(double) * (4 * size_t(K0) * K0 + 2 + 32);
    k_cluster<<<M0, 128, smem, st_cluster>>>(a, a.part_prep, nparts);
}

}  // namespace sb
