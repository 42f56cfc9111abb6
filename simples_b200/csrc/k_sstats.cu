// Per-cluster second moments of the factor posterior means:
//   S_m = sum_i z_im W_m[i,] W_m[i,]^T  (K0 x K0),   a_m = sum_i z_im W_m[i,]
// and the deterministic assembly of the all-reduce payload
//   stats = [C (ldG x NVP) | R2 (ldG) | S | a | nz | c | tot].
#include "ptx.cuh"
#include "state.h"

namespace sb {
namespace {

constexpr int SSW = 4;    // warps per CTA, each with its own cell range and partial slot
constexpr int SCH = 32;   // cells per staged chunk

// One warp owns a contiguous cell range of cluster m and accumulates S_m = (z W)^T W as KT x KT DMMA tiles
// (m8n8k4: A = (z W)^T, B = W; both fragments are the same shared-memory load, A scaled by z) plus a_m = sum z W.
template <int KT>
__global__ void __launch_bounds__(SSW * 32) k_sstats(SstatsArgs a, int cells_per_warp) {
    constexpr int KS = KT * 8 + 4;                 // row stride == 4 (mod 8): conflict-free fragment loads
    __shared__ __align__(16) double sWall[SSW][SCH * KS];
    __shared__ double sZall[SSW][SCH];
    const int K0 = a.d.K0, M0 = a.d.M0, n = a.d.n;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, g = lane >> 2, t = lane & 3;
    const int m = blockIdx.x, slot = blockIdx.y * SSW + warp;
    if (slot >= a.nsplit) return;
    double* sW = sWall[warp];
    double* sZ = sZall[warp];
    const int cbeg = slot * cells_per_warp, cend = min(n, cbeg + cells_per_warp);
    double acc[KT][KT][2], sa[KT];
#pragma unroll
    for (int i = 0; i < KT; ++i) {
        sa[i] = 0.0;
#pragma unroll
        for (int j = 0; j < KT; ++j) acc[i][j][0] = acc[i][j][1] = 0.0;
    }
    for (int i0 = cbeg; i0 < cend; i0 += SCH) {
        const int lim = min(SCH, cend - i0);
        __syncwarp();
        for (int idx = lane; idx < SCH * KT * 8; idx += 32) {
            const int r = idx / (KT * 8), k = idx - r * (KT * 8);
            sW[r * KS + k] = (r < lim && k < K0) ? a.W[(size_t(m) * n + i0 + r) * K0 + k] : 0.0;
        }
        sZ[lane] = lane < lim ? a.z[size_t(i0 + lane) * M0 + m] : 0.0;
        __syncwarp();
#pragma unroll 2
        for (int c = 0; c < SCH; c += 4) {
            const double zc = sZ[c + t];
            double w[KT], zw[KT];
#pragma unroll
            for (int i = 0; i < KT; ++i) {
                w[i] = sW[(c + t) * KS + i * 8 + g];
                zw[i] = zc * w[i];
                sa[i] += zw[i];
            }
#pragma unroll
            for (int i = 0; i < KT; ++i)
#pragma unroll
                for (int j = 0; j < KT; ++j) dmma(acc[i][j][0], acc[i][j][1], zw[i], w[j]);
        }
    }
    const int NE = K0 * K0 + K0;
    double* out = a.part + (size_t(slot) * M0 + m) * NE;
#pragma unroll
    for (int i = 0; i < KT; ++i) {
        const int row = i * 8 + g;
#pragma unroll
        for (int j = 0; j < KT; ++j) {
            const int col = j * 8 + 2 * t;
            if (row < K0 && col < K0) out[row * K0 + col] = acc[i][j][0];
            if (row < K0 && col + 1 < K0) out[row * K0 + col + 1] = acc[i][j][1];
        }
        double s = sa[i];
        s += __shfl_xor_sync(0xffffffffu, s, 1);
        s += __shfl_xor_sync(0xffffffffu, s, 2);
        if (t == 0 && row < K0) out[K0 * K0 + row] = s;
    }
}

// fixed-order sum of `np` partials with stride `stride`: four independent accumulators so that the loads pipeline
// (one running sum is a chain of np dependent FP64 adds behind np L2 round trips)
__device__ __forceinline__ double sum_parts(const double* p, size_t stride, int np) {
    double s0 = 0.0, s1 = 0.0, s2 = 0.0, s3 = 0.0;
    int q = 0;
    for (; q + 4 <= np; q += 4) {
        s0 += p[size_t(q) * stride];
        s1 += p[size_t(q + 1) * stride];
        s2 += p[size_t(q + 2) * stride];
        s3 += p[size_t(q + 3) * stride];
    }
    for (; q < np; ++q) s0 += p[size_t(q) * stride];
    return (s0 + s1) + (s2 + s3);
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
            s = sum_parts(a.part_ms + i, nC, a.splitK);
        } else if (i < nC + nR) {
            const size_t j = i - nC;
            const double* pr = a.part_ms + size_t(a.splitK) * nC;
            s = sum_parts(pr + j, nR, a.splitK);
        } else if (i < nC + nR + nS) {
            const size_t j = i - nC - nR;
            const int m = int(j / (K0 * K0)), e = int(j - size_t(m) * K0 * K0);
            s = sum_parts(a.part_ss + size_t(m) * NE + e, size_t(M0) * NE, a.nsplit_ss);
        } else if (i < nC + nR + nS + nA) {
            const size_t j = i - nC - nR - nS;
            const int m = int(j / K0), k = int(j - size_t(m) * K0);
            s = sum_parts(a.part_ss + size_t(m) * NE + K0 * K0 + k, size_t(M0) * NE, a.nsplit_ss);
        } else {
            const size_t j = i - nC - nR - nS - nA;
            s = sum_parts(a.part_cell + j, size_t(RC), a.ncta_cell);
        }
        a.stats[i] = s;
    }
}

}  // namespace

template <int KT>
static void launch_sstats_kt(const SstatsArgs& a, cudaStream_t st) {
    const int cpw = ((a.d.n + a.nsplit - 1) / a.nsplit + SCH - 1) / SCH * SCH;
    dim3 grid(a.d.M0, (a.nsplit + SSW - 1) / SSW);
    k_sstats<KT><<<grid, SSW * 32, 0, st>>>(a, cpw);
}

void launch_sstats(const SstatsArgs& a, cudaStream_t st) {
    switch ((a.d.K0 + 7) / 8) {
        case 1: launch_sstats_kt<1>(a, st); break;
        case 2: launch_sstats_kt<2>(a, st); break;
        case 3: launch_sstats_kt<3>(a, st); break;
        default: launch_sstats_kt<4>(a, st); break;
    }
}

void launch_finalize_stats(const FinalizeArgs& a, cudaStream_t st) {
    const size_t head = a.skip_head ? a.head : size_t(a.d.ldG) * a.d.NVP + a.d.ldG;
    const size_t count = head + size_t(a.d.M0) * a.d.K0 * a.d.K0 + size_t(a.d.M0) * a.d.K0 + 2 * a.d.M0 + 1;
    int grid = int((count + 255) / 256);
    if (grid > 592) grid = 592;
    k_finalize_stats<<<grid, 256, 0, st>>>(a, count);
}

}  // namespace sb
