// Per-cluster second moments of the factor posterior means:
//   S_m = sum_i z_im W_m[i,] W_m[i,]^T  (K0 x K0),   a_m = sum_i z_im W_m[i,]
// and the deterministic assembly of the all-reduce payload
//   stats = [C (ldG x NVP) | R2 (ldG) | S | a | nz | c | tot].
#include "ptx.cuh"
#include "state.h"

namespace sb {
namespace {

constexpr int CH = 256;  // cells per smem chunk

__global__ void __launch_bounds__(256) k_sstats(SstatsArgs a, int cps) {
    extern __shared__ __align__(16) double sm[];
    const int K0 = a.d.K0, M0 = a.d.M0, n = a.d.n;
    double* sW = sm;               // [CH][K0]
    double* sZ = sW + CH * K0;     // [CH]
    const int m = blockIdx.x, sp = blockIdx.y;
    const int cbeg = sp * cps, cend = min(n, cbeg + cps);
    const int NE = K0 * K0 + K0;
    double acc[5] = {0, 0, 0, 0, 0};   // NE <= 1056 <= 5 * 256
    for (int i0 = cbeg; i0 < cend; i0 += CH) {
        const int lim = min(CH, cend - i0);
        __syncthreads();
        for (int idx = threadIdx.x; idx < lim * K0; idx += blockDim.x) sW[idx] = a.W[(size_t(m) * n + i0) * K0 + idx];
        for (int idx = threadIdx.x; idx < lim; idx += blockDim.x) sZ[idx] = a.z[size_t(i0 + idx) * M0 + m];
        __syncthreads();
#pragma unroll
        for (int q = 0; q < 5; ++q) {
            const int e = threadIdx.x + q * 256;
            if (e < NE) {
                double s = acc[q];
                if (e < K0 * K0) {
                    const int j = e / K0, k = e - j * K0;
                    for (int r = 0; r < lim; ++r) s = fma(sZ[r] * sW[r * K0 + j], sW[r * K0 + k], s);
                } else {
                    const int k = e - K0 * K0;
                    for (int r = 0; r < lim; ++r) s = fma(sZ[r], sW[r * K0 + k], s);
                }
                acc[q] = s;
            }
        }
    }
#pragma unroll
    for (int q = 0; q < 5; ++q) {
        const int e = threadIdx.x + q * 256;
        if (e < NE) a.part[(size_t(sp) * M0 + m) * NE + e] = acc[q];
    }
}

__global__ void k_finalize_stats(FinalizeArgs a, size_t count) {
    const int K0 = a.d.K0, M0 = a.d.M0, ldG = a.d.ldG, NVP = a.d.NVP;
    const size_t nC = a.skip_head ? a.head : size_t(ldG) * NVP, nR = a.skip_head ? 0 : ldG;
    const size_t nS = size_t(M0) * K0 * K0, nA = size_t(M0) * K0;
    const int NE = K0 * K0 + K0, RC = 2 * M0 + 1;
    for (size_t i = blockIdx.x * size_t(blockDim.x) + threadIdx.x; i < count; i += size_t(gridDim.x) * blockDim.x) {
        double s = 0.0;
        if (a.skip_head && i < nC) continue;
        if (i < nC) {
            for (int p = 0; p < a.splitK; ++p) s += a.part_ms[size_t(p) * nC + i];
        } else if (i < nC + nR) {
            const size_t j = i - nC;
            const double* pr = a.part_ms + size_t(a.splitK) * nC;
            for (int p = 0; p < a.splitK; ++p) s += pr[size_t(p) * nR + j];
        } else if (i < nC + nR + nS) {
            const size_t j = i - nC - nR;
            const int m = int(j / (K0 * K0)), e = int(j - size_t(m) * K0 * K0);
            for (int p = 0; p < a.nsplit_ss; ++p) s += a.part_ss[(size_t(p) * M0 + m) * NE + e];
        } else if (i < nC + nR + nS + nA) {
            const size_t j = i - nC - nR - nS;
            const int m = int(j / K0), k = int(j - size_t(m) * K0);
            for (int p = 0; p < a.nsplit_ss; ++p) s += a.part_ss[(size_t(p) * M0 + m) * NE + K0 * K0 + k];
        } else {
            const size_t j = i - nC - nR - nS - nA;
            for (int p = 0; p < a.ncta_cell; ++p) s += a.part_cell[size_t(p) * RC + j];
        }
        a.stats[i] = s;
    }
}

}  // namespace

void launch_sstats(const SstatsArgs& a, cudaStream_t st) {
    const int cps = ((a.d.n + a.nsplit - 1) / a.nsplit + CH - 1) / CH * CH;
    const size_t smem = sizeof(double) * (size_t(CH) * a.d.K0 + CH);
    static size_t attr = 0;
    if (smem > 48 * 1024 && smem > attr) {
        cudaFuncSetAttribute(k_sstats, cudaFuncAttributeMaxDynamicSharedMemorySize, int(smem));
        attr = smem;
    }
    dim3 grid(a.d.M0, a.nsplit);
    k_sstats<<<grid, 256, smem, st>>>(a, cps);
}

void launch_finalize_stats(const FinalizeArgs& a, cudaStream_t st) {
    const size_t head = a.skip_head ? a.head : size_t(a.d.ldG) * a.d.NVP + a.d.ldG;
    const size_t count = head + size_t(a.d.M0) * a.d.K0 * a.d.K0 + size_t(a.d.M0) * a.d.K0 + 2 * a.d.M0 + 1;
    int grid = int((count + 255) / 256);
    if (grid > 592) grid = 592;
    k_finalize_stats<<<grid, 256, 0, st>>>(a, count);
}

}  // namespace sb
