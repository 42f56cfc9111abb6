// mc_pass: one Monte-Carlo imputation sweep FUSED with the projection of the E-pass that follows it
// (R/EM_impute.R:168-199 + :202-213, R/do_impute.R:109-136 + :73-80 of the next sweep).
//
// The round-1 path ran impute_sweep (scattered 8-byte read-modify-write of Y: 2.9 GB of DRAM traffic per
// launch) and then estep_project (re-reads all of Y: 1.6 GB).  Here Y crosses the SM exactly once per MC
// pass in each direction:
//   * a WARP owns 8 cells and walks the genes in 32-gene chunks.  Each chunk is two tensor-map TMA boxes
//     [8 cells][16 genes] (SWIZZLE_128B, SASS UTMALDG) into the warp's PRIVATE two-stage ring; there is no
//     CTA-level barrier, producer warp or shared parameter buffer anywhere, so the 24 warps of an SM are
//     fully decoupled (round 1: every chunk cost the slowest of the 8 warps sharing a parameter buffer).
//   * conditional means ms = (f B' + mu_m(i)) sd + mean by FP64 DMMA (depth K0, A-fragments in registers)
//     are written INTO the tile at the positions of the zero entries (their old values are dead);
//   * the zero entries are compacted (ballot/popc) and sampled exactly as in round 1 (FP32 screen of the
//     Bernoulli test, dense class lists, table-driven FP64 pnorm / qnorm) -- the result is written to the
//     shared-memory tile, not to global memory;
//   * the updated tile feeds P += (Y/sigma)' [B | mu] (FP64 DMMA) and S2 += Y^2/sigma while it is still in
//     shared memory, and goes back to HBM as two TMA box stores (UTMASTG): full 128-byte lines, no
//     read-modify-write in L2.
//   * gene-chunk parameters are read through L1 (ld.global.nc) from one packed row [B_g | mu_g] per gene
//     and a few per-gene vectors: warps that drift apart simply hit different lines.
//   * 8-cell tasks are handed out through one atomic counter (persistent warps), so the 148 SMs stay
//     balanced to within one task; results do not depend on which warp ran which task.
// Requires cluster-shared sigma (NS == 1: always true after the first M-step, R/EM_impute.R:269); the
// general path keeps the separate impute_sweep / estep_project kernels.
#pragma once
#include <cstdlib>

#include "ptx.cuh"
#include "rng.cuh"
#include "state.h"

namespace sb {
namespace mcp {

constexpr int RING_BYTES = 4096;      // per warp: 2 stages x 2 boxes x [8][16] doubles
constexpr int WB_DUB = 0, WB_LIST = 2048, WB_DCODE = 2304, WB_CI = 2560, WB_BAR = 2592, WBLK = 2624;
constexpr size_t smem_bytes(int MW) { return size_t(MW) * RING_BYTES + size_t(MW) * WBLK + 1024; }

// list code of tile entry (row r, gene gl): [box:1][r:3][chunk16:3][lo:1]; the swizzled tile index is
// code ^ (r << 1) (the 16-byte chunk index XOR the row, as written by TMA with SWIZZLE_128B).
__device__ __forceinline__ int code_row(int code) { return (code >> 4) & 7; }
__device__ __forceinline__ int code_gene(int code) { return ((code >> 3) & 16) | (code & 15); }
__device__ __forceinline__ int code_tix(int code) { return code ^ ((code >> 3) & 14); }

// exact FP64 Bernoulli test + draw of one entry inside the screening band; returns the centred value
static __device__ __noinline__ double uncertain_entry(double ms, const double* gq, double p, double cutoff, double ua, double ub,
                                                      int clip, double lower, double upper) {
    double x = sample_entry(ms, gq[1], gq[0], p, cutoff, ua, ub);
    if (clip) x = fmax(fmin(x, upper), lower);
    return (x - gq[2]) * gq[3];
}

template <int NKS, int NT, int MW, int MINB, int CU>
__global__ void __launch_bounds__(MW * 32, MINB)
k_mcpass(const __grid_constant__ CUtensorMap tm, const __grid_constant__ McArgs a) {
    extern __shared__ __align__(1024) unsigned char smraw[];
    unsigned char* smb = smraw + ((1024u - (smem_u32(smraw) & 1023u)) & 1023u);
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31, g = lane >> 2, t = lane & 3;
    double* yring = reinterpret_cast<double*>(smb + warp * RING_BYTES);
    unsigned char* wb = smb + MW * RING_BYTES + warp * WBLK;
    double* dub = reinterpret_cast<double*>(wb + WB_DUB);
    unsigned char* list = wb + WB_LIST;
    unsigned char* dcode = wb + WB_DCODE;
    int* cellinfo = reinterpret_cast<int*>(wb + WB_CI);
    uint64_t* bars = reinterpret_cast<uint64_t*>(wb + WB_BAR);

    const int M0 = a.d.M0, MB = a.d.MB, K0 = a.d.K0, ldG = a.d.ldG, n = a.d.n, FS = a.d.FS, NQP = a.d.NQP;
    constexpr int QS = NT * 8 + 4;
    const int KST = a.d.NFP / 4;
    const int nchunks = ldG / 32;
    const int ntasks = (n + 7) / 8;
    const unsigned FULL = 0xffffffffu;

    if (lane == 0) {
        mbar_init(&bars[0], 1);
        mbar_init(&bars[1], 1);
        mbar_fence_init();
        tma_prefetch_desc(&tm);
    }
    __syncwarp();

    const int rho = 2 * (g & 3) + (g >> 2);     // fragment row g <-> tile row (cell) rho: conflict-free A loads
    const int offA0 = swz16(rho, t);            // A-fragment element (row rho, gene 4 kk + t) = offA0 ^ (kk << 2)
    const int offC0 = swz16(rho, 2 * t), offC1 = swz16(rho, 8 + 2 * t);
    const int codelane = ((lane & 16) << 3) | (lane & 15);
    const unsigned ltmask = (1u << lane) - 1u;

    auto next_task = [&]() -> int {
        int v = 0;
        if (lane == 0) v = int(atomicAdd(a.counter, 1u));
        return __shfl_sync(FULL, v, 0);
    };
    auto issue_load = [&](int buf, int task, int chunk) {     // lane 0 only
        double* dst = yring + buf * 256;
        mbar_arrive_expect_tx(&bars[buf], 2048u);
        tma_load_2d(dst, &tm, chunk * 32, task * 8, &bars[buf]);
        tma_load_2d(dst + 128, &tm, chunk * 32 + 16, task * 8, &bars[buf]);
    };

    unsigned q = 0;                      // ring position of this warp (steps done)
    int task = next_task();
    if (task < ntasks && lane == 0) issue_load(0, task, 0);
    unsigned nword = 0u;                 // mask word of the NEXT step (lanes 0..7)
    if (task < ntasks && lane < 8) nword = a.mask[size_t(task) * 8 + lane];

    while (task < ntasks) {
        const int ntask = next_task();
        const int wc0 = task * 8;
        double fr[NKS];                  // A-fragments f[cell rho][4 ks + t]
#pragma unroll
        for (int ks = 0; ks < NKS; ++ks) fr[ks] = (ks < KST) ? a.Fh[size_t(wc0 + rho) * FS + ks * 4 + t] : 0.0;
        __syncwarp();                    // previous task's readers of cellinfo are done
        if (lane < 8) {
            const int cell = wc0 + lane;
            cellinfo[lane] = cell < n ? (a.im[cell] | (a.celltype0[cell] << 8)) : 0;
        }
        __syncwarp();
        const int imr = cellinfo[rho] & 0xff;
        const unsigned long long gcell = (unsigned long long)(a.d.cell_offset + wc0);      // wc0 is a multiple of 8: + r never carries
        const uint32_t cell_lo = (uint32_t)gcell, cell_hi8 = (uint32_t)(gcell >> 32) << 8;
        double acc[NT][2];
#pragma unroll
        for (int nt = 0; nt < NT; ++nt) acc[nt][0] = acc[nt][1] = 0.0;
        double s2 = 0.0;

        for (int c = 0; c < nchunks; ++c, ++q) {
            const int b = int(q & 1u);
            double* Yt = yring + b * 256;
            const unsigned word = nword;
            {   // mask word of the next step
                const bool more = c + 1 < nchunks;
                const int t2 = more ? task : ntask, c2 = more ? c + 1 : 0;
                nword = 0u;
                if (lane < 8 && t2 < ntasks) nword = a.mask[size_t(c2) * a.n_alloc + size_t(t2) * 8 + lane];
            }
            const int gbase = c * 32;
            const double* bm = a.BM + size_t(gbase) * QS;
            const double* gvc = a.gv + size_t(gbase) * 4;       // per-entry parameters of this chunk
            const double* pgc = a.pg + size_t(gbase) * MB;
            const bool anyz = __ballot_sync(FULL, word != 0u) != 0u;
            mbar_wait(&bars[b], (q >> 1) & 1u);      // this step's tile has landed

            int nDc = 0, nS = 0, total = 0;
            if (anyz) {
                // ---- 1. conditional means of the block, written at the zero entries' positions ----
                {
                    double cm[4][2];
#pragma unroll
                    for (int nt = 0; nt < 4; ++nt) cm[nt][0] = cm[nt][1] = 0.0;
#pragma unroll
                    for (int ks = 0; ks < NKS; ++ks) {
                        if (ks < KST) {
#pragma unroll
                            for (int nt = 0; nt < 4; ++nt)
                                dmma(cm[nt][0], cm[nt][1], fr[ks], __ldg(bm + (nt * 8 + g) * QS + ks * 4 + t));
                        }
                    }
                    const unsigned wrow = __shfl_sync(FULL, word, rho);
#pragma unroll
                    for (int nt = 0; nt < 4; ++nt) {
                        const int j0 = nt * 8 + 2 * t;
                        const unsigned bits = (wrow >> j0) & 3u;
                        if (bits) {
                            const double2 gm2 = __ldg(reinterpret_cast<const double2*>(a.gmean + gbase + j0));
                            const double2 sd2 = __ldg(reinterpret_cast<const double2*>(a.gsd + gbase + j0));
                            const double m0v = fma(cm[nt][0] + __ldg(bm + j0 * QS + K0 + imr), sd2.x, gm2.x);
                            const double m1v = fma(cm[nt][1] + __ldg(bm + (j0 + 1) * QS + K0 + imr), sd2.y, gm2.y);
                            double* dst = Yt + (nt >> 1) * 128 + ((nt & 1) ? offC1 : offC0);
                            if (bits & 1u) dst[0] = m0v;
                            if (bits & 2u) dst[1] = m1v;
                        }
                    }
                }
                // ---- 2. compact the zero entries of the block into a dense list ----
#pragma unroll
                for (int r = 0; r < 8; ++r) {
                    const unsigned w = __shfl_sync(FULL, word, r);
                    if ((w >> lane) & 1u) list[total + __popc(w & ltmask)] = (unsigned char)(codelane | (r << 4));
                    total += __popc(w);
                }
                __syncwarp();

                // ---- 3a. classify every zero entry with an FP32 screen of the Bernoulli test (see k_sweep.cu) ----
                const int rounds = (total + 31) >> 5;
#pragma unroll CU
                for (int it = 0; it < rounds; ++it) {
                    const int e = it * 32 + lane;
                    const bool valid = e < total;
                    const int code = valid ? list[e] : 0;
                    const int r = code_row(code), gl = code_gene(code), tix = code_tix(code);
                    const int ct = cellinfo[r] >> 8;
                    const double ms = Yt[tix];
                    const Philox4 px = philox4x32_10_keys((uint32_t)(gbase + gl), cell_lo + (uint32_t)r, a.sweep, cell_hi8, a.rk);
                    const double ub = u52(px.x2, px.x3);
                    const double isd = __ldg(gvc + gl * 4);
                    const float rf = (float)((a.cutoff - ms) * isd);
                    const float ax = fabsf(rf) * 0.70710678f;
                    float tt, ex;
                    asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(tt) : "f"(fmaf(0.3275911f, ax, 1.0f)));
                    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(ex) : "f"(-1.4426950f * ax * ax));
                    float po = fmaf(tt, 1.061405429f, -1.453152027f);
                    po = fmaf(tt, po, 1.421413741f);
                    po = fmaf(tt, po, -0.284496736f);
                    po = fmaf(tt, po, 0.254829592f);
                    const float hc = 0.5f * po * tt * ex;                        // 0.5 erfc(|r| / sqrt 2)
                    const float pf = rf < 0.0f ? hc : 1.0f - hc;
                    const float uaf = (float)(px.x0 >> 8) * 5.9604644775390625e-8f;
                    const float pfl = (float)__ldg(pgc + gl * MB + ct), q1f = 1.0f - pfl;
                    const float d = fmaf(uaf, fmaf(pf, pfl, q1f), -q1f);
                    const bool isD = valid && (d < -4e-6f);
                    const bool isT = valid && (d > 4e-6f);
                    const bool cen = fabs(ub - 0.5) <= 0.425;
                    const bool toDc = isD && cen, toS = (isD && !cen) || isT;
                    const unsigned bDc = __ballot_sync(FULL, toDc);
                    const unsigned bS = __ballot_sync(FULL, toS);
                    if (toDc) {
                        const int pos = nDc + __popc(bDc & ltmask);
                        dcode[pos] = (unsigned char)code;
                        dub[pos] = ub;
                    } else if (toS) {
                        const int pos = 255 - (nS + __popc(bS & ltmask));
                        dcode[pos] = (unsigned char)code;
                        dub[pos] = isT ? -ub : ub;
                    } else if (valid) {
                        // inside the screening band (~1e-5 of the entries): exact FP64 test and draw, right here (cold)
                        const double v = uncertain_entry(ms, gvc + gl * 4, __ldg(pgc + gl * MB + ct), a.cutoff, u52(px.x0, px.x1), ub,
                                                         a.clip, a.lower, a.upper);
                        Yt[tix] = v;
                        if (a.rec1) {
                            const size_t o = size_t(wc0 + r) * ldG + gbase + gl;
                            a.rec1[o] += v;
                            a.rec2[o] += v * v;
                        }
                    }
                    nDc += __popc(bDc);
                    nS += __popc(bS);
                }
                __syncwarp();
            }

            // ---- prefetch the next step's tile into the other buffer (its store was issued a step ago) ----
            if (lane == 0) {
                const bool more = c + 1 < nchunks;
                const int t2 = more ? task : ntask, c2 = more ? c + 1 : 0;
                if (t2 < ntasks) {
                    bulk_wait_read<0>();
                    issue_load(b ^ 1, t2, c2);
                }
            }

            if (anyz) {
                auto finish = [&](int code, double x) {
                    const int gl = code_gene(code);
                    if (a.clip) x = fmax(fmin(x, a.upper), a.lower);
                    const double2 gi = __ldg(reinterpret_cast<const double2*>(gvc + gl * 4 + 2));
                    const double v = (x - gi.x) * gi.y;
                    Yt[code_tix(code)] = v;
                    if (a.rec1) {
                        const size_t o = size_t(wc0 + code_row(code)) * ldG + gbase + gl;
                        a.rec1[o] += v;
                        a.rec2[o] += v * v;
                    }
                };
                // ---- 3b. T-stage: quantile argument ub * Phi(r) of the certain truncated draws, re-filed by region ----
                const bool inplace = nDc + 2 * nS > 256;
                int nTail = 0;
                {
                    const int srounds = (nS + 31) >> 5;
#pragma unroll 1
                    for (int it = 0; it < srounds; ++it) {
                        const int e = it * 32 + lane;
                        const bool valid = e < nS;
                        const int code = valid ? dcode[255 - e] : 0;
                        const double v = valid ? dub[255 - e] : 1.0;
                        const bool isT = v < 0.0;
                        double arg = v;
                        bool toC = false, toTail = valid;
                        if (isT) {
                            const double ms = Yt[code_tix(code)];
                            const double2 gs = __ldg(reinterpret_cast<const double2*>(gvc + code_gene(code) * 4));
                            const double rr = (a.cutoff - ms) * gs.x;
                            if (rr < -30.0) {                // deep tail: log-space Newton, finished right here (cold)
                                finish(code, fma(gs.y, trunc_z_deep(-v, rr), ms));
                                toTail = false;
                            } else {
                                const double qq = -v * pnorm(rr);
                                arg = -qq;
                                toC = fabs(qq - 0.5) <= 0.425;
                                toTail = !toC;
                            }
                        }
                        if (inplace) {                       // warp-uniform
                            if (valid) dub[255 - e] = (toC || toTail) ? arg : 0.0;      // 0 marks "already finished"
                            continue;
                        }
                        __syncwarp();
                        const unsigned bC = __ballot_sync(FULL, toC);
                        const unsigned bT = __ballot_sync(FULL, toTail);
                        if (toC) {
                            const int pos = nDc + __popc(bC & ltmask);
                            dcode[pos] = (unsigned char)code;
                            dub[pos] = arg;
                        } else if (toTail) {
                            const int pos = 255 - (nTail + __popc(bT & ltmask));
                            dcode[pos] = (unsigned char)code;
                            dub[pos] = arg;
                        }
                        nDc += __popc(bC);
                        nTail += __popc(bT);
                        __syncwarp();
                    }
                }
                // ---- 3c. central quantiles: x = ms + sds Phi^-1(arg)   (R/EM_impute.R:186 / :191-192) ----
#pragma unroll CU
                for (int e = lane; e < nDc; e += 32) {
                    const int code = dcode[e];
                    const double v = dub[e];
                    const double sds = __ldg(gvc + code_gene(code) * 4 + 1);
                    double x = fma(sds, qnorm_central(fabs(v)), Yt[code_tix(code)]);
                    if (v < 0.0) x = fmin(x, a.cutoff);
                    finish(code, x);
                }
                // ---- 3c'. tail quantiles ----
#pragma unroll 1
                for (int e = lane; e < nTail; e += 32) {
                    const int code = dcode[255 - e];
                    const double v = dub[255 - e];
                    const double sds = __ldg(gvc + code_gene(code) * 4 + 1);
                    double x = fma(sds, qnorm_tail(fabs(v)), Yt[code_tix(code)]);
                    if (v < 0.0) x = fmin(x, a.cutoff);
                    finish(code, x);
                }
                // ---- 3c''. dense block: the whole back list, arguments in place, either quantile region ----
                if (inplace) {
                    __syncwarp();
#pragma unroll 1
                    for (int e = lane; e < nS; e += 32) {
                        const int code = dcode[255 - e];
                        const double v = dub[255 - e];
                        if (v != 0.0) {
                            const double sds = __ldg(gvc + code_gene(code) * 4 + 1);
                            finish(code, quantile_draw(Yt[code_tix(code)], sds, a.cutoff, v));
                        }
                    }
                }
                fence_proxy_async();       // the tile was written through the generic proxy; the store reads it through the async one
                __syncwarp();
                if (lane == 0) {
                    tma_store_2d(&tm, Yt, gbase, wc0);
                    tma_store_2d(&tm, Yt + 128, gbase + 16, wc0);
                    bulk_commit();
                }
            }

            // ---- 4. projection of the updated tile: P += (Y / sigma)' [B | mu],  S2 += Y^2 / sigma ----
#pragma unroll
            for (int ks = 0; ks < 8; ++ks) {
                const double w = __ldg(a.isg + gbase + ks * 4 + t);
                const double av = Yt[(ks >> 2) * 128 + (offA0 ^ ((ks & 3) << 2))];
                const double aw = av * w;
                s2 = fma(aw, av, s2);
                const double* brow = bm + (ks * 4 + t) * QS + g;
#pragma unroll
                for (int nt = 0; nt < NT; ++nt) dmma(acc[nt][0], acc[nt][1], aw, __ldg(brow + nt * 8));
            }
            __syncwarp();                  // every lane has read the tile before the next step may overwrite the other one
        }

        // ---- task epilogue: P / S2 rows of the 8 cells ----
        {
            const int cell = wc0 + rho;
            double s = s2;
            s += __shfl_xor_sync(FULL, s, 1);
            s += __shfl_xor_sync(FULL, s, 2);
            if (cell < n) {
                double* prow = a.P + size_t(cell) * NQP;
#pragma unroll
                for (int nt = 0; nt < NT; ++nt)
                    if (nt * 8 < NQP)
                        *reinterpret_cast<double2*>(prow + nt * 8 + 2 * t) = make_double2(acc[nt][0], acc[nt][1]);
                if (t == 0) a.S2[cell] = s;
            }
        }
        task = ntask;
    }
    if (lane == 0) bulk_wait_read<0>();    // shared memory must outlive the last box store
}

template <int NKS, int NT, int MW, int MINB, int CU>
void launch_mcpass_cfg(const McArgs& a, const CUtensorMap& tm, cudaStream_t st) {
    constexpr size_t smem = smem_bytes(MW);
    ensure_dyn_smem_k(k_mcpass<NKS, NT, MW, MINB, CU>, smem);
    const int ntasks = (a.d.n + 7) / 8;
    int grid = NUM_SMS_B200 * MINB;
    if (grid * MW > ntasks) grid = (ntasks + MW - 1) / MW;
    k_mcpass<NKS, NT, MW, MINB, CU><<<grid, MW * 32, smem, st>>>(tm, a);
}

template <int NKS, int NT>
void launch_mcpass_variant(const McArgs& a, const CUtensorMap& tm, cudaStream_t st) {
#ifdef SIMPLES_MC_EXPERIMENT
    if (NKS == 3 && NT == 3) {           // tuning hook: SIMPLES_MC_CFG = index of a (warps/CTA, min CTAs/SM, unroll) variant
        const char* e = getenv("SIMPLES_MC_CFG");
        switch (e ? atoi(e) : 0) {
            case 1: launch_mcpass_cfg<3, 3, 8, 2, 2>(a, tm, st); return;     // 128 regs, 16 warps / SM
            case 2: launch_mcpass_cfg<3, 3, 10, 2, 2>(a, tm, st); return;    //  96 regs, 20 warps
            case 3: launch_mcpass_cfg<3, 3, 8, 3, 1>(a, tm, st); return;     //  80 regs, 24 warps, no unroll
            case 4: launch_mcpass_cfg<3, 3, 8, 2, 1>(a, tm, st); return;     // 128 regs, 16 warps, no unroll
            case 5: launch_mcpass_cfg<3, 3, 10, 2, 1>(a, tm, st); return;    //  96 regs, 20 warps, no unroll
            case 6: launch_mcpass_cfg<3, 3, 12, 1, 2>(a, tm, st); return;    // 168 regs, 12 warps
            case 7: launch_mcpass_cfg<3, 3, 7, 3, 2>(a, tm, st); return;     //  96 regs, 21 warps
            default: break;
        }
    }
#endif
    // 24 resident warps per SM and NO unrolling of the per-entry loops measured best at C3 (profiles/r02_notes.md):
    // occupancy beats registers here, and the 6 KB L0 / 32 KB L1.5 instruction caches thrash on unrolled code once
    // the warps are decoupled
    if (NKS <= 3 && NT <= 4) launch_mcpass_cfg<NKS, NT, 8, 3, 1>(a, tm, st);
    else launch_mcpass_cfg<NKS, NT, 8, 2, 1>(a, tm, st);
}

}  // namespace mcp
}  // namespace sb
