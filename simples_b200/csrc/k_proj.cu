// estep_project: P = Y^T Q (n x NQP) and S2_i = sum_g Y_gi^2 / sigma_g, the skinny FP64 GEMM of
// the E-pass (replaces the n*M0 R closures at R/EM_impute.R:126-136 and the Wm GEMM at :155-157;
// SURVEY.md 3.3).  FP64 DMMA (mma.sync m8n8k4) fed by a bulk-async-copy (TMA engine) mbarrier ring.
//
// CTA = 8 consumer warps (16 cells each) + 1 producer warp; tile = 128 cells x all genes, streamed
// in 32-gene chunks: stage = Y[128][32] (row stride 36 doubles => conflict-free fragment loads),
// Q[32][QS], isg[32].  Persistent grid over cell tiles.
#include "ptx.cuh"
#include "state.h"

namespace sb {

namespace {
constexpr int TC = 128;     // cells per tile
constexpr int GC = 32;      // genes per chunk
constexpr int YS = 36;      // smem row stride of the Y tile (== 4 mod 16)
constexpr int NCW = 8;      // consumer warps
constexpr int THREADS = (NCW + 1) * 32;

__host__ __device__ constexpr size_t stage_bytes(int QS) {
    return size_t(TC) * YS * 8 + size_t(GC) * QS * 8 + GC * 8;
}

template <int NT>
__global__ void __launch_bounds__(THREADS, 1) k_proj(ProjArgs a, int stages, int ntiles) {
    constexpr int QS = NT * 8 + 4;
    extern __shared__ __align__(128) unsigned char smem[];
    uint64_t* full = reinterpret_cast<uint64_t*>(smem);
    uint64_t* empty = full + 8;
    unsigned char* stage0 = smem + 128;
    const size_t SB = stage_bytes(QS);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (threadIdx.x == 0) {
        for (int s = 0; s < stages; ++s) {
            mbar_init(&full[s], 1);
            mbar_init(&empty[s], NCW);
        }
        mbar_fence_init();
    }
    __syncthreads();

    const int nchunks = a.ldG / GC;

    if (warp == NCW) {
        // ---------------- producer ----------------
        int stage = 0;
        uint32_t phase = 0;
        for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
            const int cell0 = tile * TC;
            const int rows = min(TC, a.n - cell0);
            for (int c = 0; c < nchunks; ++c) {
                mbar_wait(&empty[stage], phase ^ 1);
                unsigned char* sp = stage0 + stage * SB;
                double* Ys = reinterpret_cast<double*>(sp);
                double* Qc = Ys + TC * YS;
                double* Ic = Qc + GC * QS;
                if (lane == 0) {
                    mbar_arrive_expect_tx(&full[stage], uint32_t(rows) * GC * 8 + GC * QS * 8 + GC * 8);
                    bulk_g2s(Qc, a.Q + size_t(c) * GC * QS, GC * QS * 8, &full[stage]);
                    bulk_g2s(Ic, a.isg + size_t(c) * GC, GC * 8, &full[stage]);
                }
                __syncwarp();
                for (int r = lane; r < rows; r += 32)
                    bulk_g2s(Ys + r * YS, a.Y + size_t(cell0 + r) * a.ldG + size_t(c) * GC, GC * 8, &full[stage]);
                if (++stage == stages) { stage = 0; phase ^= 1; }
            }
        }
    } else {
        // ---------------- consumers ----------------
        const int g = lane >> 2, t = lane & 3;
        int stage = 0;
        uint32_t phase = 0;
        for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
            const int cell0 = tile * TC;
            double acc[2][NT][2];
            double s2[2] = {0.0, 0.0};
#pragma unroll
            for (int mt = 0; mt < 2; ++mt)
#pragma unroll
                for (int nt = 0; nt < NT; ++nt) acc[mt][nt][0] = acc[mt][nt][1] = 0.0;

            for (int c = 0; c < nchunks; ++c) {
                mbar_wait(&full[stage], phase);
                const unsigned char* sp = stage0 + stage * SB;
                const double* Ys = reinterpret_cast<const double*>(sp);
                const double* Qc = Ys + TC * YS;
                const double* Ic = Qc + GC * QS;
                const double* ya = Ys + (warp * 16 + g) * YS + t;
#pragma unroll
                for (int ks = 0; ks < GC / 4; ++ks) {
                    double b[NT];
#pragma unroll
                    for (int nt = 0; nt < NT; ++nt) b[nt] = Qc[(ks * 4 + t) * QS + nt * 8 + g];
                    const double w = Ic[ks * 4 + t];
#pragma unroll
                    for (int mt = 0; mt < 2; ++mt) {
                        const double av = ya[mt * 8 * YS + ks * 4];
#pragma unroll
                        for (int nt = 0; nt < NT; ++nt) dmma(acc[mt][nt][0], acc[mt][nt][1], av, b[nt]);
                        s2[mt] = fma(av * av, w, s2[mt]);
                    }
                }
                __syncwarp();
                if (lane == 0) mbar_arrive(&empty[stage]);
                if (++stage == stages) { stage = 0; phase ^= 1; }
            }
            // epilogue: C[g][2t,2t+1] per (mt, nt)
#pragma unroll
            for (int mt = 0; mt < 2; ++mt) {
                const int cell = cell0 + warp * 16 + mt * 8 + g;
                double s = s2[mt];
                s += __shfl_xor_sync(0xffffffffu, s, 1);
                s += __shfl_xor_sync(0xffffffffu, s, 2);
                if (cell < a.n) {
                    double* prow = a.P + size_t(cell) * a.NQP;
#pragma unroll
                    for (int nt = 0; nt < NT; ++nt)
                        *reinterpret_cast<double2*>(prow + nt * 8 + 2 * t) = make_double2(acc[mt][nt][0], acc[mt][nt][1]);
                    if (t == 0) a.S2[cell] = s;
                }
            }
        }
    }
}

template <int NT>
void launch_nt(const ProjArgs& a, cudaStream_t st) {
    constexpr int QS = NT * 8 + 4;
    const size_t SB = stage_bytes(QS);
    int stages = int((227 * 1024 - 128) / SB);
    if (stages > 4) stages = 4;
    const size_t smem = 128 + stages * SB;
    static bool attr_set = false;
    if (!attr_set) {
        cudaFuncSetAttribute(k_proj<NT>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
        attr_set = true;
    }
    const int ntiles = (a.n + TC - 1) / TC;
    const int grid = ntiles < NUM_SMS_B200 ? ntiles : NUM_SMS_B200;
    k_proj<NT><<<grid, THREADS, smem, st>>>(a, stages, ntiles);
}
}  // namespace

void launch_proj(const ProjArgs& a, cudaStream_t st) {
    switch (a.NQP / 8) {
        case 1: launch_nt<1>(a, st); break;
        case 2: launch_nt<2>(a, st); break;
        case 3: launch_nt<3>(a, st); break;
        case 4: launch_nt<4>(a, st); break;
        case 5: launch_nt<5>(a, st); break;
        case 6: launch_nt<6>(a, st); break;
        case 7: launch_nt<7>(a, st); break;
        case 8: launch_nt<8>(a, st); break;
        default: break;   // validated by the engine (K0 + M0 <= 64)
    }
}

}  // namespace sb
