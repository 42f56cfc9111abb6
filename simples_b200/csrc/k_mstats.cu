// mstep_stats: the cell-additive sufficient statistics of the M-step that replace the
// per-gene (M0 n + K0) x K0 augmented designs of R/EM_impute.R:238-253:
//     C  = Y V      (ldG x NVP),  V_i = [Wbar_i | z_i | sqrt z_i]   ->  Tbar, U, Uh
//     R2 = Y.^2 zsum                                                 ->  sum_m U2[,m]
// FP64 DMMA split-K GEMM over cells, fed by a bulk-async-copy mbarrier ring.  Each CTA owns a
// 64-gene tile and one cell range; partials are reduced in fixed order (deterministic).
#include "ptx.cuh"
#include "state.h"

namespace sb {
namespace {

constexpr int MT = 64;      // genes per CTA
constexpr int KC = 32;      // cells per stage
constexpr int YS = 68;      // smem row stride of the Y stage (== 4 mod 16)
constexpr int NCW = 4;      // consumer warps (16 genes each)
constexpr int THREADS = (NCW + 1) * 32;
constexpr int MAX_STAGES = 4;   // ring depth: 4, or fewer when that lets two CTAs share an SM (wide operands)

__host__ __device__ constexpr size_t stage_bytes(int VS) { return size_t(KC) * YS * 8 + size_t(KC) * VS * 8 + KC * 8; }

template <int NT>
__global__ void __launch_bounds__(THREADS) k_mstats(MstatsArgs a, int cps, int STAGES) {
    constexpr int VS = NT * 8 + 4;
    extern __shared__ __align__(128) unsigned char smem[];
    uint64_t* full = reinterpret_cast<uint64_t*>(smem);
    uint64_t* empty = full + 8;
    unsigned char* stage0 = smem + 128;
    constexpr size_t SB = stage_bytes(VS);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (threadIdx.x == 0) {
        for (int s = 0; s < STAGES; ++s) {
            mbar_init(&full[s], 1);
            mbar_init(&empty[s], NCW);
        }
        mbar_fence_init();
    }
    __syncthreads();

    const int m0 = blockIdx.x * MT;
    const int sk = blockIdx.y;
    const int cbeg = sk * cps;
    const int cend = min(cbeg + cps, (a.n + KC - 1) / KC * KC);   // buffers are zero-padded to a multiple of 128 cells
    const int nst = cend > cbeg ? (cend - cbeg) / KC : 0;

    if (warp == NCW) {
        int stage = 0;
        uint32_t phase = 0;
        for (int s = 0; s < nst; ++s) {
            const int c0 = cbeg + s * KC;
            mbar_wait(&empty[stage], phase ^ 1);
            unsigned char* sp = stage0 + stage * SB;
            double* Ys = reinterpret_cast<double*>(sp);
            double* Vs = Ys + KC * YS;
            double* Zs = Vs + KC * VS;
            if (lane == 0) {
                mbar_arrive_expect_tx(&full[stage], KC * MT * 8 + KC * VS * 8 + KC * 8);
                bulk_g2s(Vs, a.V + size_t(c0) * VS, KC * VS * 8, &full[stage]);
                bulk_g2s(Zs, a.zsum + c0, KC * 8, &full[stage]);
            }
            __syncwarp();
            bulk_g2s(Ys + lane * YS, a.Y + size_t(c0 + lane) * a.ldG + m0, MT * 8, &full[stage]);
            if (++stage == STAGES) { stage = 0; phase ^= 1; }
        }
    } else {
        const int g = lane >> 2, t = lane & 3;
        double acc[2][NT][2];
        double r2[2] = {0.0, 0.0};
#pragma unroll
        for (int mt = 0; mt < 2; ++mt)
#pragma unroll
            for (int nt = 0; nt < NT; ++nt) acc[mt][nt][0] = acc[mt][nt][1] = 0.0;
        int stage = 0;
        uint32_t phase = 0;
        for (int s = 0; s < nst; ++s) {
            mbar_wait(&full[stage], phase);
            const unsigned char* sp = stage0 + stage * SB;
            const double* Ys = reinterpret_cast<const double*>(sp);
            const double* Vs = Ys + KC * YS;
            const double* Zs = Vs + KC * VS;
            const double* ya = Ys + t * YS + warp * 16 + g;
#pragma unroll
            for (int ks = 0; ks < KC / 4; ++ks) {
                double b[NT];
#pragma unroll
                for (int nt = 0; nt < NT; ++nt) b[nt] = Vs[(ks * 4 + t) * VS + nt * 8 + g];
                const double w = Zs[ks * 4 + t];
#pragma unroll
                for (int mt = 0; mt < 2; ++mt) {
                    const double av = ya[ks * 4 * YS + mt * 8];
#pragma unroll
                    for (int nt = 0; nt < NT; ++nt) dmma(acc[mt][nt][0], acc[mt][nt][1], av, b[nt]);
                    r2[mt] = fma(av * av, w, r2[mt]);
                }
            }
            __syncwarp();
            if (lane == 0) mbar_arrive(&empty[stage]);
            if (++stage == STAGES) { stage = 0; phase ^= 1; }
        }
        double* pc = a.part + size_t(sk) * a.ldG * a.NVP;
        double* pr = a.part + size_t(a.splitK) * a.ldG * a.NVP + size_t(sk) * a.ldG;
#pragma unroll
        for (int mt = 0; mt < 2; ++mt) {
            const int gene = m0 + warp * 16 + mt * 8 + g;
            double s = r2[mt];
            s += __shfl_xor_sync(0xffffffffu, s, 1);
            s += __shfl_xor_sync(0xffffffffu, s, 2);
#pragma unroll
            for (int nt = 0; nt < NT; ++nt)
                *reinterpret_cast<double2*>(pc + size_t(gene) * a.NVP + nt * 8 + 2 * t) =
                    make_double2(acc[mt][nt][0], acc[mt][nt][1]);
            if (t == 0) pr[gene] = s;
        }
    }
}

template <int NT>
void launch_nt(const MstatsArgs& a, cudaStream_t st) {
    constexpr int VS = NT * 8 + 4;
    // two resident CTAs (8 consumer warps) keep the DMMA pipe fed; with one CTA per SM the K0 = M0 = 20 operand (NT = 6)
    // ran at 52 % of the DGEMM rate against 76 % at NT = 3 (profiles/r02_notes.md)
    int stages = MAX_STAGES;
    while (stages > 2 && 2 * (128 + stages * stage_bytes(VS) + 1024) > 227 * 1024) --stages;
    const size_t smem = 128 + stages * stage_bytes(VS);
    ensure_dyn_smem_k(k_mstats<NT>, size_t(int(smem)));
    const int cps = (((a.n + a.splitK - 1) / a.splitK) + KC - 1) / KC * KC;
    dim3 grid(a.ldG / MT, a.splitK);
    k_mstats<NT><<<grid, THREADS, smem, st>>>(a, cps, stages);
}

__global__ void k_reduce_parts(const double* part, double* out, size_t count, int nparts) {
    for (size_t i = blockIdx.x * size_t(blockDim.x) + threadIdx.x; i < count; i += size_t(gridDim.x) * blockDim.x) {
        double s = 0.0;
        for (int p = 0; p < nparts; ++p) s += part[size_t(p) * count + i];
        out[i] = s;
    }
}

// generic (slow) path: C[g][c] = sum_i f(Y[i][g]) V[i][c], thread per gene, 8 columns per pass
__global__ void k_ygemm_simple(const double* Y, const double* V, double* part, int n, int ldG, int ncol, int vstride,
                               int square, int cps) {
    const int gene = blockIdx.x * blockDim.x + threadIdx.x;
    const int sk = blockIdx.y;
    const int cbeg = sk * cps, cend = min(n, cbeg + cps);
    __shared__ double sv[32][8];
    for (int c0 = 0; c0 < ncol; c0 += 8) {
        double acc[8] = {0, 0, 0, 0, 0, 0, 0, 0};
        for (int i0 = cbeg; i0 < cend; i0 += 32) {
            __syncthreads();
            for (int idx = threadIdx.x; idx < 256; idx += blockDim.x) {
                const int r = idx >> 3, c = idx & 7;
                sv[r][c] = (i0 + r < cend && c0 + c < ncol) ? V[size_t(i0 + r) * vstride + c0 + c] : 0.0;
            }
            __syncthreads();
            const int lim = min(32, cend - i0);
            for (int r = 0; r < lim; ++r) {
                double y = Y[size_t(i0 + r) * ldG + gene];
                if (square) y *= y;
#pragma unroll
                for (int c = 0; c < 8; ++c) acc[c] = fma(y, sv[r][c], acc[c]);
            }
        }
        for (int c = 0; c < 8 && c0 + c < ncol; ++c) part[(size_t(sk) * ldG + gene) * ncol + c0 + c] = acc[c];
    }
}

}  // namespace

int mstats_pick_splitk(int ldG) {
    const int mtiles = ldG / MT;
    int s = (2 * NUM_SMS_B200) / mtiles;
    return s < 1 ? 1 : s;
}

void launch_mstats(const MstatsArgs& a, cudaStream_t st) {
    switch (a.NVP / 8) {
        case 1: launch_nt<1>(a, st); break;
        case 2: launch_nt<2>(a, st); break;
        case 3: launch_nt<3>(a, st); break;
        case 4: launch_nt<4>(a, st); break;
        case 5: launch_nt<5>(a, st); break;
        case 6: launch_nt<6>(a, st); break;
        case 7: launch_nt<7>(a, st); break;
        case 8: launch_nt<8>(a, st); break;
        case 9: launch_nt<9>(a, st); break;
        case 10: launch_nt<10>(a, st); break;
        case 11: launch_nt<11>(a, st); break;
        case 12: launch_nt<12>(a, st); break;
        default: break;
    }
}

void launch_reduce_parts(const double* part, double* out, size_t count, int nparts, cudaStream_t st) {
    int grid = int((count + 255) / 256);
    if (grid > 1184) grid = 1184;
    if (grid < 1) grid = 1;
    k_reduce_parts<<<grid, 256, 0, st>>>(part, out, count, nparts);
}

void launch_ygemm_simple(const double* Y, const double* V, double* part, int n, int ldG, int ncol, int vstride, int square,
                         int nsplit, cudaStream_t st) {
    const int cps = ((n + nsplit - 1) / nsplit + 31) / 32 * 32;
    dim3 grid(ldG / 64, nsplit);
    k_ygemm_simple<<<grid, 64, 0, st>>>(Y, V, part, n, ldG, ncol, vstride, square, cps);
}

}  // namespace sb
