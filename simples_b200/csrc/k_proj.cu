// estep_project: P = Y^T Q (n x NQP) and S2_i = sum_g Y_gi^2 / sigma_g, the skinny FP64 GEMM of
// the E-pass (replaces the n*M0 R closures at R/EM_impute.R:126-136 and the Wm GEMM at :155-157;
// SURVEY.md 3.3).  FP64 DMMA (mma.sync m8n8k4) fed by tensor-map TMA through an mbarrier ring.
//
// CTA = 8 consumer warps (16 cells each) + 1 producer warp; tile = 128 cells x all genes, streamed
// in 32-gene chunks.  Stage = two TMA boxes [128 cells][16 genes] (SWIZZLE_128B, 16 KB per
// cp.async.bulk.tensor) + Q[32][QS] + isg[32].  The m-tile row -> cell map rho(g) = 2(g&3) + (g>>2)
// makes every DMMA A-fragment load hit 32 distinct banks under the 128B swizzle.
// (A ring of 256-byte 1-D bulk copies was TMA-issue bound at ~1 TB/s: profiles/r01_notes.md.)
#include <algorithm>

#include "ptx.cuh"
#include "state.h"
#include "tmap.h"

namespace sb {

namespace {
constexpr int TC = 128;     // cells per tile
constexpr int GC = 32;      // genes per chunk
constexpr int NCW = 8;      // consumer warps
constexpr int THREADS = (NCW + 1) * 32;
constexpr int YBOX = TC * 16;   // doubles per TMA box

__host__ __device__ constexpr size_t stage_bytes(int QS) {
    return size_t(2) * YBOX * 8 + size_t(GC) * QS * 8 + GC * 8 + 768;   // keep every stage 1024-B aligned
}
static_assert((2 * YBOX * 8) % 1024 == 0, "box alignment");

// P / S2 rows of the gene-split tail tiles = fixed-order sum of the S partial rows
__global__ void k_proj_reduce(ProjArgs a, int cell_begin, int rows, int S) {
    const int e = blockIdx.x * blockDim.x + threadIdx.x;
    const int W = a.NQP + 1;
    if (e >= rows * W) return;
    const int r = e / W, c = e - r * W;
    const int cell = cell_begin + r;
    if (cell >= a.n) return;
    double s = 0.0;
    for (int p = 0; p < S; ++p) s += a.part[(size_t(p) * rows + r) * W + c];
    if (c < a.NQP) a.P[size_t(cell) * a.NQP + c] = s;
    else a.S2[cell] = s;
}

template <int NT>
__global__ void __launch_bounds__(THREADS, 1) k_proj(const __grid_constant__ CUtensorMap tmapY, ProjArgs a, int stages,
                                                     int tile_begin, int nitems, int S) {
    // Work item w = (tile, part): tile = tile_begin + w / S over the gene-chunk range part of S.  S = 1 writes P / S2
    // directly; S > 1 (the last, partial round of tiles, split along the genes so that every SM has work) writes
    // partial sums to a.part, summed in a fixed order by k_proj_reduce.
    constexpr int QS = NT * 8 + 4;
    constexpr size_t SB = (stage_bytes(QS) + 1023) / 1024 * 1024;
    extern __shared__ __align__(1024) unsigned char smem_raw[];
    unsigned char* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);   // TMA swizzle needs 1024-B alignment
    unsigned char* stage0 = smem;
    uint64_t* full = reinterpret_cast<uint64_t*>(smem + size_t(stages) * SB);
    uint64_t* empty = full + 8;

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (threadIdx.x == 0) {
        for (int s = 0; s < stages; ++s) {
            mbar_init(&full[s], 1);
            mbar_init(&empty[s], NCW);
        }
        mbar_fence_init();
        tma_prefetch_desc(&tmapY);
    }
    __syncthreads();

    const int nchunks = a.ldG / GC;

    if (warp == NCW) {
        // ---------------- producer (one elected lane) ----------------
        if (lane == 0) {
            int stage = 0;
            uint32_t phase = 0;
            for (int w = blockIdx.x; w < nitems; w += gridDim.x) {
                const int tile = tile_begin + w / S, part = w - (w / S) * S;
                const int cell0 = tile * TC;
                const int cb = nchunks * part / S, ce = nchunks * (part + 1) / S;
                for (int c = cb; c < ce; ++c) {
                    mbar_wait(&empty[stage], phase ^ 1);
                    unsigned char* sp = stage0 + stage * SB;
                    double* Ys = reinterpret_cast<double*>(sp);
                    double* Qc = Ys + 2 * YBOX;
                    double* Ic = Qc + GC * QS;
                    mbar_arrive_expect_tx(&full[stage], 2 * YBOX * 8 + GC * QS * 8 + GC * 8);
                    tma_load_2d(Ys, &tmapY, c * GC, cell0, &full[stage]);
                    tma_load_2d(Ys + YBOX, &tmapY, c * GC + 16, cell0, &full[stage]);
                    bulk_g2s(Qc, a.Q + size_t(c) * GC * QS, GC * QS * 8, &full[stage]);
                    bulk_g2s(Ic, a.isg + size_t(c) * GC, GC * 8, &full[stage]);
                    if (++stage == stages) { stage = 0; phase ^= 1; }
                }
            }
        }
    } else {
        // ---------------- consumers ----------------
        const int g = lane >> 2, t = lane & 3;
        const int rho = 2 * (g & 3) + (g >> 2);          // row -> cell permutation (bank-conflict-free under the swizzle)
        const int r0 = warp * 16 + rho;                  // m-tile 0 row; m-tile 1 = r0 + 8
        // swizzled offsets of (row r0, gene 4*kk + t), kk = 0..3, inside a box
        int off[4];
#pragma unroll
        for (int kk = 0; kk < 4; ++kk) off[kk] = swz16(r0, 4 * kk + t);
        int stage = 0;
        uint32_t phase = 0;
        for (int w = blockIdx.x; w < nitems; w += gridDim.x) {
            const int tile = tile_begin + w / S, part = w - (w / S) * S;
            const int cell0 = tile * TC;
            const int cb = nchunks * part / S, ce = nchunks * (part + 1) / S;
            double acc[2][NT][2];
            double s2[2] = {0.0, 0.0};
#pragma unroll
            for (int mt = 0; mt < 2; ++mt)
#pragma unroll
                for (int nt = 0; nt < NT; ++nt) acc[mt][nt][0] = acc[mt][nt][1] = 0.0;

            for (int c = cb; c < ce; ++c) {
                mbar_wait(&full[stage], phase);
                const unsigned char* sp = stage0 + stage * SB;
                const double* Ys = reinterpret_cast<const double*>(sp);
                const double* Qc = Ys + 2 * YBOX;
                const double* Ic = Qc + GC * QS;
#pragma unroll
                for (int ks = 0; ks < GC / 4; ++ks) {
                    double b[NT];
#pragma unroll
                    for (int nt = 0; nt < NT; ++nt) b[nt] = Qc[(ks * 4 + t) * QS + nt * 8 + g];
                    const double w = Ic[ks * 4 + t];
                    const double* yb = Ys + (ks >> 2) * YBOX + off[ks & 3];
#pragma unroll
                    for (int mt = 0; mt < 2; ++mt) {
                        const double av = yb[mt * 8 * 16];     // row + 8: same (row & 7) => same swizzle
#pragma unroll
                        for (int nt = 0; nt < NT; ++nt) dmma(acc[mt][nt][0], acc[mt][nt][1], av, b[nt]);
                        s2[mt] = fma(av * av, w, s2[mt]);
                    }
                }
                __syncwarp();
                if (lane == 0) mbar_arrive(&empty[stage]);
                if (++stage == stages) { stage = 0; phase ^= 1; }
            }
            // epilogue: C[row g][2t, 2t+1] per (mt, nt); row g of m-tile mt is cell cell0 + 16 warp + 8 mt + rho
#pragma unroll
            for (int mt = 0; mt < 2; ++mt) {
                const int cell = cell0 + r0 + mt * 8;
                double s = s2[mt];
                s += __shfl_xor_sync(0xffffffffu, s, 1);
                s += __shfl_xor_sync(0xffffffffu, s, 2);
                if (S > 1) {       // partial row: [part][(tile - tile_begin) * TC + row][NQP + 1]
                    const size_t prows = size_t(nitems / S) * TC;
                    double* prow = a.part + (size_t(part) * prows + size_t(tile - tile_begin) * TC + (r0 + mt * 8)) * (a.NQP + 1);
#pragma unroll
                    for (int nt = 0; nt < NT; ++nt) {
                        prow[nt * 8 + 2 * t] = acc[mt][nt][0];
                        prow[nt * 8 + 2 * t + 1] = acc[mt][nt][1];
                    }
                    if (t == 0) prow[a.NQP] = s;
                } else if (cell < a.n) {
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

// gene split of the last partial round: S in 1..PROJ_MAX_SPLIT minimising ceil(rem S / NUM_SMS) / S, >= 4 chunks per part
int proj_pick_split(int rem, int nchunks) {
    if (rem <= 0) return 1;
    int best = 1;
    double cost = 1.0;                    // S = 1: one whole round
    for (int S = 2; S <= PROJ_MAX_SPLIT && nchunks / S >= 4; ++S) {
        const double c = double((rem * S + NUM_SMS_B200 - 1) / NUM_SMS_B200) / S + 0.004 * S;     // small cost per extra partial row
        if (c < cost - 1e-9) {
            cost = c;
            best = S;
        }
    }
    return best;
}

template <int NT>
void launch_nt(const ProjArgs& a, const CUtensorMap& tm, cudaStream_t st) {
    constexpr int QS = NT * 8 + 4;
    const size_t SB = (stage_bytes(QS) + 1023) / 1024 * 1024;
    int stages = int((227 * 1024 - 256 - 1024) / SB);
    if (stages > 5) stages = 5;
    const size_t smem = stages * SB + 256 + 1024;
    ensure_dyn_smem_k(k_proj<NT>, size_t(227 * 1024));
    const int ntiles = (a.n + TC - 1) / TC;
    const int nchunks = a.ldG / GC;
    // Full rounds of NUM_SMS tiles run as they are; a last partial round of `rem` tiles is split along the genes into
    // S parts per tile so that it costs ceil(rem S / NUM_SMS) / S of a round instead of a whole one (a small shard of a
    // strongly scaled run is ALL tail: 98 tiles at 12.5k cells -> 3 parts, 2 rounds of a third).
    const int head = (ntiles / NUM_SMS_B200) * NUM_SMS_B200, rem = ntiles - head;
    const int S = a.part ? proj_pick_split(rem, nchunks) : 1;
    if (S <= 1) {
        const int grid = ntiles < NUM_SMS_B200 ? ntiles : NUM_SMS_B200;
        k_proj<NT><<<grid, THREADS, smem, st>>>(tm, a, stages, 0, ntiles, 1);
        return;
    }
    if (head > 0) k_proj<NT><<<NUM_SMS_B200, THREADS, smem, st>>>(tm, a, stages, 0, head, 1);
    k_proj<NT><<<std::min(rem * S, NUM_SMS_B200), THREADS, smem, st>>>(tm, a, stages, head, rem * S, S);
    const int rows = rem * TC;
    k_proj_reduce<<<(rows * (a.NQP + 1) + 255) / 256, 256, 0, st>>>(a, head * TC, rows, S);
}
}  // namespace

// kernels enqueued by one launch_proj (the tail split adds a second k_proj and the reduction)
int proj_launches(int n, int ldG, bool have_scratch) {
    const int ntiles = (n + TC - 1) / TC, nchunks = ldG / GC;
    const int head = (ntiles / NUM_SMS_B200) * NUM_SMS_B200, rem = ntiles - head;
    const int S = have_scratch ? proj_pick_split(rem, nchunks) : 1;
    return S <= 1 ? 1 : (head > 0 ? 3 : 2);
}

void launch_proj(const ProjArgs& a, const CUtensorMap& tmapY, cudaStream_t st) {
    switch (a.NQP / 8) {
        case 1: launch_nt<1>(a, tmapY, st); break;
        case 2: launch_nt<2>(a, tmapY, st); break;
        case 3: launch_nt<3>(a, tmapY, st); break;
        case 4: launch_nt<4>(a, tmapY, st); break;
        case 5: launch_nt<5>(a, tmapY, st); break;
        case 6: launch_nt<6>(a, tmapY, st); break;
        case 7: launch_nt<7>(a, tmapY, st); break;
        case 8: launch_nt<8>(a, tmapY, st); break;
        default: break;   // validated by the engine (K0 + M0 <= 64)
    }
}

}  // namespace sb
