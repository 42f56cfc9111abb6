// impute_sweep: one Monte-Carlo imputation sweep over every entry with Y0 <= cutoff
// (R/EM_impute.R:168-199, R/do_impute.R:109-136).
//
// A warp owns 8 cells and walks the genes in 32-gene chunks (= one mask word per cell):
//   1. ms(8 cells x 32 genes) = Fh * Qs^T by FP64 DMMA (Fh_i = f_i, Qs_g = B_g sd_g; depth K0, A-fragments in
//      registers) + (mu_g,m(i) sd_g + mean_g) gathered from the chunk parameters -> per-warp shared-memory tile;
//   2. the set bits of the 8 mask words are compacted (ballot + popc) into a dense work list, so
//      the expensive per-entry sampler runs with full lanes instead of ~44 % (the zero fraction);
//   3. each list entry draws from the Philox tape and writes the centred value to Y.
// Gene-chunk parameters (Qs, sds, 1/sds, mu sd, pg, {mean, 1/sd}) travel through a three-stage shared-memory ring filled by
// the TMA engine (bulk copies completing on an mbarrier).  The warps of a CTA are not synchronised per chunk: the LAST warp
// to leave a stage (counted with a shared-memory atomic) refills it, so nobody ever waits for a free stage and a fast warp
// may run up to two chunks ahead of the slowest one.
#include <cmath>
#include <cstdlib>

#include "ptx.cuh"
#include "rng.cuh"
#include "state.h"

namespace sb {
namespace {

constexpr int SW = 8;          // default warps per CTA (template parameter SWT of the kernel)
constexpr int SG = 32;         // genes per chunk (= one mask word)
constexpr int MSS = 34;        // row stride of the per-warp ms tile (doubles)
constexpr int NST = 3;         // stages of the parameter ring (2 when a third one would cost a resident CTA)

struct ChunkBuf {
    double *Q, *sds, *isd, *musd, *pg;
    double2* gi;       // {gene mean, 1 / gene sd}
};

__device__ __forceinline__ ChunkBuf chunk_buf(double* base, int FS, int M0, int MB) {
    ChunkBuf b;
    b.Q = base;
    b.sds = b.Q + SG * FS;
    b.isd = b.sds + SG * M0;
    b.musd = b.isd + SG * M0;
    b.pg = b.musd + SG * M0;
    b.gi = reinterpret_cast<double2*>(b.pg + SG * MB);
    return b;
}
__host__ __device__ inline int chunk_doubles(int FS, int M0, int MB) { return SG * FS + 3 * SG * M0 + SG * MB + 2 * SG; }

// one thread: TMA bulk copies of the parameter spans of gene chunk `chunk`, completing on `bar`
__device__ __forceinline__ void load_chunk_bulk(const SweepArgs& a, const ChunkBuf& b, int chunk, uint64_t* bar) {
    const int g0 = chunk * SG, FS = a.d.FS, M0 = a.d.M0, MB = a.d.MB;
    mbar_arrive_expect_tx(bar, uint32_t(chunk_doubles(FS, M0, MB)) * 8u);
    bulk_g2s(b.Q, a.Qs + size_t(g0) * FS, SG * FS * 8, bar);
    bulk_g2s(b.sds, a.sdsA + size_t(g0) * M0, SG * M0 * 8, bar);
    bulk_g2s(b.isd, a.isdA + size_t(g0) * M0, SG * M0 * 8, bar);
    bulk_g2s(b.musd, a.musdA + size_t(g0) * M0, SG * M0 * 8, bar);
    bulk_g2s(b.pg, a.pg + size_t(g0) * MB, SG * MB * 8, bar);
    bulk_g2s(b.gi, a.gmig + size_t(g0) * 2, SG * 16, bar);
}

// mbarrier wait that sleeps between polls: a blocked warp must not spend issue slots the working warps need
__device__ __forceinline__ void mbar_wait_backoff(uint64_t* bar, uint32_t parity) {
    const uint32_t addr = smem_u32(bar);
    for (;;) {
        uint32_t done;
        asm volatile(
            "{\n"
            ".reg .pred p;\n"
            "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n"
            "selp.u32 %0, 1, 0, p;\n"
            "}\n"
            : "=r"(done)
            : "r"(addr), "r"(parity)
            : "memory");
        if (done) break;
        __nanosleep(200);
    }
}

__device__ __forceinline__ unsigned atom_add_acq_rel(unsigned* p, unsigned v) {
    unsigned old;
    asm volatile("atom.acq_rel.cta.shared::cta.add.u32 %0, [%1], %2;" : "=r"(old) : "r"(smem_u32(p)), "r"(v) : "memory");
    return old;
}

// NKS = number of K-steps of 4 held as register A-fragments (K0 <= 4 NKS); MINB = resident CTAs per SM aimed at;
// SWT = warps per CTA (8 cells each)
template <int NKS, int MINB, int SWT>
__global__ void __launch_bounds__(SWT * 32, MINB) k_sweep(const __grid_constant__ SweepArgs a) {
    constexpr int SW = SWT, SC = SWT * 8;
    extern __shared__ __align__(16) double sm[];
    const int FS = a.d.FS, M0 = a.d.M0, MB = a.d.MB, ldG = a.d.ldG, n = a.d.n;
    const int KST = a.d.NFP / 4, nst = a.nst;
    const int CD = chunk_doubles(FS, M0, MB);
    // Shared-memory layout: fixed-size blocks first (barriers, then one statically laid out block per warp) so that
    // all per-warp addressing is `warp base + compile-time offset`; the runtime-sized parameter buffers follow.
    constexpr int WBLK = 8 * MSS * 8 + 256 * 8 + 768 + 128;     // bytes per warp: ms tile | dub | codes | cell info
    unsigned char* smb = reinterpret_cast<unsigned char*>(sm);
    uint64_t* full = reinterpret_cast<uint64_t*>(smb);                            // full[NST]: the stage's chunk has landed
    unsigned* cnt = reinterpret_cast<unsigned*>(smb + 32);                        // cnt[NST]:  warps that have left the stage
    unsigned char* wblk = smb + 64 + (threadIdx.x >> 5) * WBLK;
    double* sBuf = reinterpret_cast<double*>(smb + 64 + SW * WBLK);               // [NST][CD] parameter chunks
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31, g = lane >> 2, t = lane & 3;
    const int cell0 = blockIdx.x * SC;
    // The 8-cell groups of the CTA ROTATE over the warps: in chunk ci warp w owns group (w + ci) mod SW.  The number of zero
    // entries is a property of the cell (capture efficiency), so with a fixed assignment the warp holding the sparsest
    // cells ran ahead until the ring stopped it and then spun on the barrier of every chunk (36 % of the waits blocked,
    // 6.8 % of all executed instructions: profiles/r02_notes.md section 8); rotating evens the work out within 8 chunks.
    // Cost: the A-fragments (3 doubles per lane, L1-resident) are re-read per chunk.
    if (lane < 8) {
        const int cell = cell0 + warp * 8 + lane;
        // per cell: cluster, cell type and the cell's two Philox counter words (global cell index: low word, high word << 8)
        const unsigned long long gc = (unsigned long long)(a.d.cell_offset + cell);
        reinterpret_cast<int4*>(wblk + 8 * MSS * 8 + 256 * 8 + 768)[lane] =
            make_int4(cell < n ? a.im[cell] : 0, cell < n ? a.celltype0[cell] : 0, int(uint32_t(gc)), int(uint32_t(gc >> 32) << 8));
    }
    // gridDim.y splits the gene range (wave-quantisation fix: more, smaller CTAs); this CTA walks chunks [cbeg, cend)
    const int nchunks_all = ldG / SG;
    const int cbeg = int((long long)nchunks_all * blockIdx.y / gridDim.y);
    const int cend = int((long long)nchunks_all * (blockIdx.y + 1) / gridDim.y);
    const int nchunks = cend - cbeg;
    if (tid == 0) {
#pragma unroll
        for (int s = 0; s < NST; ++s) {
            mbar_init(&full[s], 1);
            cnt[s] = 0u;
        }
        mbar_fence_init();
    }
    __syncthreads();
    if (tid == 0) {
        for (int s = 0; s < nst && s < nchunks; ++s) load_chunk_bulk(a, chunk_buf(sBuf + s * CD, FS, M0, MB), cbeg + s, &full[s]);
    }

    double* msW = reinterpret_cast<double*>(wblk);                     // [8][MSS] conditional means / deltas
    double* dub = msW + 8 * MSS;                                       // [256]    list values
    unsigned char* list = reinterpret_cast<unsigned char*>(dub + 256); // [256] all zero entries
    unsigned char* dcode = list + 256;                                 // [256] Dc (front) / back list codes
    unsigned char* xcode = list + 512;                                 // [256] uncertain entries
    const unsigned lt = (1u << lane) - 1u;

    // mask words of the group's 8 cells (chunk-major mask => 32 contiguous bytes), fetched one chunk ahead
    auto mask_word = [&](int chunk, int grp) {
        const int cell = cell0 + grp * 8 + lane;
        return (lane < 8 && cell < n) ? a.mask[size_t(chunk) * a.n_alloc + cell] : 0u;
    };
    unsigned wnext = nchunks > 0 ? mask_word(cbeg, warp) : 0u;
    int st = 0;                 // ring stage of the current chunk
    unsigned use = 0;           // how often the ring has wrapped (= ci / nst)
    int grp = warp;             // this chunk's cell group
    for (int ci = 0; ci < nchunks; ++ci, grp = (grp + 1 == SW ? 0 : grp + 1)) {
        const int chunk = cbeg + ci;                   // absolute gene-chunk index
        const int wc0 = cell0 + grp * 8;               // its first cell
        const int* cellinfo = reinterpret_cast<const int*>(smb + 64 + grp * WBLK + 8 * MSS * 8 + 256 * 8 + 768);   // [8][4]
        const unsigned word = wnext;
        wnext = ci + 1 < nchunks ? mask_word(chunk + 1, grp + 1 == SW ? 0 : grp + 1) : 0u;
        double fr[NKS];                                // A-fragments f[cell g][4 ks + t] of the group's 8 cells
        {
            const int cell = wc0 + g;
#pragma unroll
            for (int ks = 0; ks < NKS; ++ks) fr[ks] = (ks < KST && cell < n) ? __ldg(a.Fh + size_t(cell) * FS + ks * 4 + t) : 0.0;
        }
        mbar_wait_backoff(&full[st], use & 1u);        // this chunk's parameters have landed
        const ChunkBuf b = chunk_buf(sBuf + st * CD, FS, M0, MB);

        const size_t blk0 = size_t(wc0) * ldG + chunk * SG;
        double* const yblk = a.Y + blk0;
        const size_t rec0 = a.rec1 ? size_t(a.mi_gbase[wc0 >> 3]) + a.mi_off[size_t(wc0 >> 3) * (ldG / SG) + chunk] : 0;
        auto finish = [&](int r, int gl, double x, bool trunc) {
            // clamp to the cutoff (truncated draws, R/EM_impute.R:191-192) and to [lower, upper] (:196-197) -- the three
            // bounds are folded into two on the host (+-inf when the sweep does not clip); plain compare-and-select, the
            // operands are never NaN
            const double h = trunc ? a.hi_trunc : a.hi;
            x = x > h ? h : x;
            x = x < a.lo ? a.lo : x;
            const double2 gi = b.gi[gl];
            const double v = (x - gi.x) * gi.y;
            const int o = r * ldG + gl;                 // offset inside the warp's 8 x 32 block (32-bit arithmetic)
            yblk[o] = v;
            if (a.rec1) {      // do_impute: slot = block base + position of the entry in the block's list (row-major)
                const unsigned* wl = reinterpret_cast<const unsigned*>(list);
                const size_t slot = rec0 + wl[8 + r] + __popc(wl[r] & ((1u << gl) - 1u));
                double2* rp = reinterpret_cast<double2*>(a.rec1) + slot;      // {sum v, sum v^2}
                double2 q = *rp;
                q.x += v;
                q.y = fma(v, v, q.y);
                *rp = q;
            }
        };

        if (__ballot_sync(0xffffffffu, word != 0u) != 0u) {
        // ---- 1. conditional means of the 8 x 32 block ----
        double acc[4][2];
#pragma unroll
        for (int nt = 0; nt < 4; ++nt) acc[nt][0] = acc[nt][1] = 0.0;
#pragma unroll
        for (int ks = 0; ks < NKS; ++ks) {
            if (ks < KST) {
#pragma unroll
                for (int nt = 0; nt < 4; ++nt) dmma(acc[nt][0], acc[nt][1], fr[ks], b.Q[(nt * 8 + g) * FS + ks * 4 + t]);
            }
        }
        {   // + (mu_{g, m(i)} sd_g + mean_g) for the lane's cell g and its genes nt*8 + 2t, +1
            const int img = cellinfo[g * 4];
#pragma unroll
            for (int nt = 0; nt < 4; ++nt) {
                const int g0 = nt * 8 + 2 * t;
                acc[nt][0] += b.musd[g0 * M0 + img];
                acc[nt][1] += b.musd[(g0 + 1) * M0 + img];
            }
        }
#pragma unroll
        for (int nt = 0; nt < 4; ++nt)
            *reinterpret_cast<double2*>(msW + g * MSS + nt * 8 + 2 * t) = make_double2(acc[nt][0], acc[nt][1]);

        // ---- 2. compact the zero entries of the block into a dense list ----
        int total = 0;
#pragma unroll
        for (int r = 0; r < 8; ++r) {
            const unsigned w = __shfl_sync(0xffffffffu, word, r);
            if ((w >> lane) & 1u) list[total + __popc(w & lt)] = (unsigned char)((r << 5) | lane);
            total += __popc(w);
        }
        __syncwarp();

        if (a.expect) {      // SIMPLES_MODE_EXPECT (warp-uniform): every entry gets its conditional expectation, no variates
#pragma unroll 1
            for (int e = lane; e < total; e += 32) {
                const int code = list[e];
                const int r = code >> 5, gl = code & 31;
                const int im = cellinfo[r * 4], ct = cellinfo[r * 4 + 1];
                finish(r, gl, expect_entry(msW[r * MSS + gl], b.sds[gl * M0 + im], b.isd[gl * M0 + im], b.pg[gl * MB + ct], a.cutoff),
                       false);
            }
        } else {
        // ---- 3a. classify every zero entry with an FP32 screen of the Bernoulli test ----
        // Phi(r) in float by Abramowitz-Stegun 7.1.26 (|err| < 5e-7 incl. MUFU.RCP / MUFU.EX2); an entry whose
        // test statistic clears the +-4e-6 band has a CERTAIN outcome:
        //   Dc: certain dropout, central quantile  -> x = ms + sds qnorm_central(ub), no FP64 pnorm at all
        //   S : certain dropout with a tail quantile (ub stored) or certain "expressed but <= cutoff" (-ub stored)
        //   U : inside the band (~1e-5 of the entries) -> exact FP64 decision
        int nDc = 0, nS = 0, nU = 0;
        const int rounds = (total + 31) >> 5;
#pragma unroll 2
        for (int it = 0; it < rounds; ++it) {
            const int e = it * 32 + lane;
            const bool valid = e < total;
            const int code = valid ? list[e] : 0;
            const int r = code >> 5, gl = code & 31;
            const int4 cinf = reinterpret_cast<const int4*>(cellinfo)[r];
            const int im = cinf.x, ct = cinf.y;
            const double ms = msW[r * MSS + gl];
            const Philox4 px = philox4x32_10_keys((uint32_t)(chunk * SG + gl), (uint32_t)cinf.z, a.sweep, (uint32_t)cinf.w, a.rk);
            const double ub = u52(px.x2, px.x3);
            const float rf = (float)((a.cutoff - ms) * b.isd[gl * M0 + im]);
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
            const float pfl = (float)b.pg[gl * MB + ct], q1f = 1.0f - pfl;
            const float d = fmaf(uaf, fmaf(pf, pfl, q1f), -q1f);
            const bool isD = valid && (d < -4e-6f);
            const bool isT = valid && (d > 4e-6f);
            const bool cen = fabs(ub - 0.5) <= 0.425;
            const bool toDc = isD && cen, toS = (isD && !cen) || isT;
            const unsigned bDc = __ballot_sync(0xffffffffu, toDc);
            const unsigned bS = __ballot_sync(0xffffffffu, toS);
            const unsigned bV = __ballot_sync(0xffffffffu, valid);
            {   // front list (Dc) grows upwards from 0, back list (S) downwards from 255: one store pair for either
                const int p = __popc((toDc ? bDc : bS) & lt);
                const int pos = toDc ? nDc + p : 255 - nS - p;
                if (toDc || toS) {
                    dcode[pos] = (unsigned char)code;
                    dub[pos] = isT ? -ub : ub;
                }
            }
            const unsigned bU = bV & ~(bDc | bS);
            if (bU) {                                               // rare (~1e-5 of the entries)
                if ((bU >> lane) & 1u) xcode[nU + __popc(bU & lt)] = (unsigned char)code;
                nU += __popc(bU);
            }
            nDc += __popc(bDc);
            nS += __popc(bS);
        }
        __syncwarp();
        if (a.rec1) {        // the list is consumed: keep the 8 mask words and their exclusive prefix counts in its place
            unsigned pre = 0;
#pragma unroll
            for (int r = 0; r < 8; ++r) {
                const unsigned w = __shfl_sync(0xffffffffu, word, r);
                if (lane == r) {
                    reinterpret_cast<unsigned*>(list)[r] = w;
                    reinterpret_cast<unsigned*>(list)[8 + r] = pre;
                }
                pre += __popc(w);
            }
            __syncwarp();
        }

        // ---- 3b. T-stage: for the certain "expressed but <= cutoff" entries of the back list compute the quantile
        //      argument ub * Phi(r) (the only place the FP64 pnorm runs) and re-file them by quantile region:
        //      central -> appended to the Dc list, tail -> compacted back list.  Stored value < 0 marks a
        //      truncated draw (clamped to the cutoff, R/EM_impute.R:191-192).
        // Re-filing appends to the front list while back-list slots are still unread, which is only safe when the
        // two lists cannot meet: in a block dense in zeros (nDc + 2 nS > 256) the arguments are written IN PLACE
        // instead and the back list is drained entry by entry afterwards (3c'').
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
                const int r = code >> 5, gl = code & 31;
                const int im = cellinfo[r * 4];
                const bool isT = v < 0.0;
                double arg = v;                      // tail-quantile dropout: stays a tail entry as it is
                bool toC = false, toTail = valid;
                if (isT) {
                    const double ms = msW[r * MSS + gl];
                    const double rr = (a.cutoff - ms) * b.isd[gl * M0 + im];
                    if (rr < -30.0) {                // deep tail: log-space Newton, finished right here (cold)
                        finish(r, gl, fma(b.sds[gl * M0 + im], trunc_z_deep(-v, rr), ms), false);
                        toTail = false;
                    } else {
                        const double q = -v * pnorm(rr);
                        arg = -q;
                        toC = fabs(q - 0.5) <= 0.425;
                        toTail = !toC;
                    }
                }
                if (inplace) {                       // warp-uniform
                    if (valid) dub[255 - e] = (toC || toTail) ? arg : 0.0;      // 0 marks "already finished"
                    continue;
                }
                __syncwarp();                        // every lane has read its slot before slots are rewritten
                const unsigned bC = __ballot_sync(0xffffffffu, toC);
                const unsigned bT = __ballot_sync(0xffffffffu, toTail);
                {
                    const int p = __popc((toC ? bC : bT) & lt);
                    const int pos = toC ? nDc + p : 255 - nTail - p;
                    if (toC || toTail) {
                        dcode[pos] = (unsigned char)code;
                        dub[pos] = arg;
                    }
                }
                nDc += __popc(bC);
                nTail += __popc(bT);
                __syncwarp();
            }
        }
        // ---- 3c. central quantiles: x = ms + sds Phi^-1(arg)   (R/EM_impute.R:186 / :191-192) ----
#pragma unroll 1
        for (int e = lane; e < nDc; e += 32) {
            const int code = dcode[e];
            const double v = dub[e];
            const int r = code >> 5, gl = code & 31;
            finish(r, gl, fma(b.sds[gl * M0 + cellinfo[r * 4]], qnorm_central(fabs(v)), msW[r * MSS + gl]), v < 0.0);
        }
        // ---- 3c'. tail quantiles ----
#pragma unroll 1
        for (int e = lane; e < nTail; e += 32) {
            const int code = dcode[255 - e];
            const double v = dub[255 - e];
            const int r = code >> 5, gl = code & 31;
            finish(r, gl, fma(b.sds[gl * M0 + cellinfo[r * 4]], qnorm_tail(fabs(v)), msW[r * MSS + gl]), v < 0.0);
        }
        // ---- 3c''. dense block: the whole back list, arguments in place, either quantile region ----
        if (inplace) {
            __syncwarp();
#pragma unroll 1
            for (int e = lane; e < nS; e += 32) {
                const int code = dcode[255 - e];
                const double v = dub[255 - e];
                const int r = code >> 5, gl = code & 31;
                if (v != 0.0) finish(r, gl, quantile_draw(msW[r * MSS + gl], b.sds[gl * M0 + cellinfo[r * 4]], a.cutoff, v), v < 0.0);
            }
        }
        // ---- 3d. inside the screening band: exact FP64 Bernoulli test and draw ----
#pragma unroll 1
        for (int e = lane; e < nU; e += 32) {
            const int code = xcode[e];
            const int r = code >> 5, gl = code & 31;
            const int im = cellinfo[r * 4], ct = cellinfo[r * 4 + 1];
            double ua, ub;
            entry_uniforms(a.seed, a.sweep, (unsigned long long)(a.d.cell_offset + wc0 + r), (uint32_t)(chunk * SG + gl), ua, ub);
            finish(r, gl, sample_entry(msW[r * MSS + gl], b.sds[gl * M0 + im], b.isd[gl * M0 + im], b.pg[gl * MB + ct],
                                       a.cutoff, ua, ub), false);
        }
        }   // sampling mode
        }   // block has zero entries
        // ---- leave the stage; the last warp out refills it with the chunk `nst` further on ----
        __syncwarp();
        if (lane == 0) {
            const unsigned old = atom_add_acq_rel(&cnt[st], 1u);
            if (old == SW * (use + 1u) - 1u && ci + nst < nchunks) load_chunk_bulk(a, b, chunk + nst, &full[st]);
        }
        if (++st == nst) {
            st = 0;
            ++use;
        }
    }
}

}  // namespace

template <int NKS, int MINB, int SWT>
static void launch_variant(const SweepArgs& a, int tiles, int gsplit, size_t smem, cudaStream_t st) {
    ensure_dyn_smem_k(k_sweep<NKS, MINB, SWT>, smem);
    k_sweep<NKS, MINB, SWT><<<dim3(tiles, gsplit), SWT * 32, smem, st>>>(a);
}

void launch_sweep(SweepArgs a, cudaStream_t st) {
    philox_round_keys(a.seed, a.rk);
    a.hi = a.clip ? a.upper : HUGE_VAL;
    a.lo = a.clip ? a.lower : -HUGE_VAL;
    a.hi_trunc = std::min(a.cutoff, a.hi);
    const int kst = a.d.NFP / 4;
    // warps per CTA: 8 (3 CTAs of <= 80 registers per SM).  The 64-register builds with 10 x 3 and 16 x 2 resident warps
    // were measured at C3 (profiles/r02_notes.md section 8): 1.70 ms and 1.57 ms against 1.57 ms -- the kernel is bound by
    // dispatch (FP64 instructions take two slots), not by latency, so more resident warps buy nothing.
    const int sw = SW;
    auto smem_for = [&](int nst) {
        return sizeof(double) * (size_t(sw) * 8 * MSS + nst * chunk_doubles(a.d.FS, a.d.M0, a.d.MB)) + sw * 256 * sizeof(double) +
               sw * 768 + sw * 32 * sizeof(int) + 64;
    };
    auto resident_for = [](size_t smem) { return int(std::min<size_t>(8, (227 * 1024) / (smem + 1024))); };
    a.nst = resident_for(smem_for(NST)) >= 3 || resident_for(smem_for(NST)) == resident_for(smem_for(2)) ? NST : 2;
    const size_t smem = smem_for(a.nst);
    const int resident = resident_for(smem);
    const int tiles = (a.d.n + sw * 8 - 1) / (sw * 8);
    // gridDim.y splits the gene range: more, smaller CTAs even out the tail of the last wave.  Measured on B200 at
    // 12.5k-100k cells x 2000 genes (profiles/r02_notes.md): 5-8 parts are within 1 % of each other and up to 15 % faster
    // than the wave-count model of round 1 at the small shards of a strongly scaled run; keep >= 8 chunks per CTA.
    int gsplit = std::min(6, std::max(1, (a.d.ldG / SG) / 8));
    if (const char* e = getenv("SIMPLES_SWEEP_GSPLIT")) {       // tuning override
        const int v = atoi(e);
        if (v >= 1 && (a.d.ldG / SG) / v >= 1) gsplit = v;
    }
    if (kst <= 3) launch_variant<3, 3, SW>(a, tiles, gsplit, smem, st);
    else if (kst <= 6 && resident >= 3) launch_variant<6, 3, SW>(a, tiles, gsplit, smem, st);
    else if (kst <= 6) launch_variant<6, 2, SW>(a, tiles, gsplit, smem, st);
    else launch_variant<8, 2, SW>(a, tiles, gsplit, smem, st);
}

}  // namespace sb
