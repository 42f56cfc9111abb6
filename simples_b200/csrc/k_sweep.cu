// impute_sweep: one Monte-Carlo imputation sweep over every entry with Y0 <= cutoff
// (R/EM_impute.R:168-199, R/do_impute.R:109-136).  The conditional means of a 64-cell x 32-gene
// block are one small FP64 DMMA product  ms = Fh * Qs^T  (Fh_i = [f_i | onehot(m_i) | 1],
// Qs_g = [B_g sd_g | mu_g. sd_g | mean_g]) so that no per-cell gathers of mu/beta are needed; each
// thread then samples the zero entries of its accumulator fragment from the Philox tape and writes
// the centred value to Y (and the do_impute accumulators when recording).
#include "ptx.cuh"
#include "rng.cuh"
#include "state.h"

namespace sb {
namespace {

constexpr int SC = 64;    // cells per CTA
constexpr int SG = 32;    // genes per chunk (= one mask word)

__global__ void __launch_bounds__(256) k_sweep(SweepArgs a) {
    extern __shared__ __align__(16) double sm[];
    const int FS = a.d.FS, M0 = a.d.M0, MB = a.d.MB, ldG = a.d.ldG, n = a.d.n;
    const int KST = a.d.NFP / 4;
    double* sF = sm;                     // [SC][FS]
    double* sQ = sF + SC * FS;           // [SG][FS]
    double* sSig = sQ + SG * FS;         // [SG][M0]  (sqrt(sigma) * sd)
    double* sPg = sSig + SG * M0;        // [SG][MB]
    double* sGm = sPg + SG * MB;         // [SG]
    double* sGs = sGm + SG;              // [SG]

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, g = lane >> 2, t = lane & 3;
    const int cell0 = blockIdx.x * SC;
    for (int idx = threadIdx.x; idx < SC * FS; idx += blockDim.x) {
        const int r = idx / FS, c = idx - r * FS;
        sF[idx] = (cell0 + r < n) ? a.Fh[size_t(cell0 + r) * FS + c] : 0.0;
    }
    const int cell = cell0 + warp * 8 + g;
    const bool valid = cell < n;
    const int im = valid ? a.im[cell] : 0;
    const int ct = valid ? a.celltype0[cell] : 0;
    const unsigned long long gcell = (unsigned long long)(a.d.cell_offset + cell);
    const int nwords = ldG / 32;

    for (int chunk = 0; chunk < nwords; ++chunk) {
        const int gene0 = chunk * SG;
        __syncthreads();
        for (int idx = threadIdx.x; idx < SG * FS; idx += blockDim.x) sQ[idx] = a.Qs[size_t(gene0) * FS + idx];
        for (int idx = threadIdx.x; idx < SG * M0; idx += blockDim.x) {
            const int r = idx / M0;
            sSig[idx] = sqrt(a.sigma[size_t(gene0) * M0 + idx]) * a.gsd[gene0 + r];
        }
        for (int idx = threadIdx.x; idx < SG * MB; idx += blockDim.x) sPg[idx] = a.pg[size_t(gene0) * MB + idx];
        if (threadIdx.x < SG) {
            sGm[threadIdx.x] = a.gmean[gene0 + threadIdx.x];
            sGs[threadIdx.x] = a.gsd[gene0 + threadIdx.x];
        }
        __syncthreads();
        const uint32_t word = valid ? a.mask[size_t(cell) * nwords + chunk] : 0u;
        if (__ballot_sync(0xffffffffu, word != 0u) == 0u) continue;

        double acc[4][2];
#pragma unroll
        for (int nt = 0; nt < 4; ++nt) acc[nt][0] = acc[nt][1] = 0.0;
        const double* fa = sF + (warp * 8 + g) * FS + t;
        for (int ks = 0; ks < KST; ++ks) {
            const double av = fa[ks * 4];
#pragma unroll
            for (int nt = 0; nt < 4; ++nt) dmma(acc[nt][0], acc[nt][1], av, sQ[(nt * 8 + g) * FS + ks * 4 + t]);
        }
#pragma unroll
        for (int nt = 0; nt < 4; ++nt) {
#pragma unroll
            for (int j = 0; j < 2; ++j) {
                const int gl = nt * 8 + 2 * t + j;
                if ((word >> gl) & 1u) {
                    const int gene = gene0 + gl;
                    const double ms = acc[nt][j];
                    const double sds = sSig[gl * M0 + im];
                    const double p = sPg[gl * MB + ct];
                    double ua, ub;
                    entry_uniforms(a.seed, a.sweep, gcell, (uint32_t)gene, ua, ub);
                    double x = sample_entry(ms, sds, p, a.cutoff, ua, ub);
                    if (a.clip) {
                        x = fmin(x, a.upper);
                        x = fmax(x, a.lower);
                    }
                    const double v = (x - sGm[gl]) / sGs[gl];
                    const size_t o = size_t(cell) * ldG + gene;
                    a.Y[o] = v;
                    if (a.rec1) {
                        a.rec1[o] += v;
                        a.rec2[o] += v * v;
                    }
                }
            }
        }
    }
}

}  // namespace

void launch_sweep(const SweepArgs& a, cudaStream_t st) {
    const size_t smem = sizeof(double) * (size_t(SC) * a.d.FS + size_t(SG) * a.d.FS + SG * a.d.M0 + SG * a.d.MB + 2 * SG);
    static size_t attr = 0;
    if (smem > 48 * 1024 && smem > attr) {
        cudaFuncSetAttribute(k_sweep, cudaFuncAttributeMaxDynamicSharedMemorySize, int(smem));
        attr = smem;
    }
    const int grid = (a.d.n + SC - 1) / SC;
    k_sweep<<<grid, 256, smem, st>>>(a);
}

}  // namespace sb
