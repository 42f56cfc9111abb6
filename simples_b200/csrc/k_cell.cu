// cell_post: per-cell epilogue of an E-pass.  From the projections P = Y^T Q and S2 it forms the
// Woodbury Mahalanobis distance, log-likelihoods, responsibilities z, log-lik total, factor posterior
// means W_m = x2_m M_m (R/EM_impute.R:126-157), the M-step operand row V_i, and -- when an MC sweep
// follows -- the cell-level draws of that sweep (membership + factor vector, R/EM_impute.R:163-171).
// One warp per cell; lane k owns factor k, lane m owns cluster m.  Deterministic per-CTA partials.
#include "ptx.cuh"
#include "rng.cuh"
#include "state.h"

namespace sb {
namespace {

constexpr int CW = 8;   // warps per CTA

__global__ void __launch_bounds__(CW * 32) k_cell(CellArgs a) {
    extern __shared__ __align__(16) double sm[];
    const int M0 = a.d.M0, K0 = a.d.K0, NS = a.d.NS, n = a.d.n, NQP = a.d.NQP;
    double* sMm = sm;                        // [M0][K0][K0]
    double* sOff = sMm + M0 * K0 * K0;       // [M0][K0]
    double* sOffq = sOff + M0 * K0;          // [M0][K0]
    double* sCmu = sOffq + M0 * K0;          // [M0]
    double* sDt = sCmu + M0;                 // [M0]
    double* sLp = sDt + M0;                  // [M0]
    double* sRed = sLp + M0;                 // [CW][2*M0+1]
    for (int i = threadIdx.x; i < M0 * K0 * K0; i += blockDim.x) sMm[i] = a.cc.Mm[i];
    for (int i = threadIdx.x; i < M0 * K0; i += blockDim.x) {
        sOff[i] = a.cc.off[i];
        sOffq[i] = a.cc.offq[i];
    }
    for (int i = threadIdx.x; i < M0; i += blockDim.x) {
        sCmu[i] = a.cc.cmu[i];
        sDt[i] = a.cc.dt[i];
        sLp[i] = a.cc.logpi[i];
    }
    __syncthreads();

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const bool kl = lane < K0, ml = lane < M0;
    double nz_acc = 0.0, c_acc = 0.0, tot_acc = 0.0;
    const size_t pstride = size_t(n) * NQP;

    for (int i = blockIdx.x * CW + warp; i < n; i += gridDim.x * CW) {
        const double* prow = a.P + size_t(i) * NQP;
        double llv = -INFINITY;
        // ---- pass 1: log-likelihoods ----
        for (int m = 0; m < M0; ++m) {
            const int cs = (NS == 1) ? 0 : m;
            const int qs = a.stale_q ? NS - 1 : cs;
            const double* off = a.stale_q ? sOffq : sOff;
            const double x2 = kl ? prow[qs * pstride + lane] - off[m * K0 + lane] : 0.0;
            double t = 0.0;
            const double* Mm = sMm + m * K0 * K0;
            for (int j = 0; j < K0; ++j) {
                const double xj = __shfl_sync(0xffffffffu, x2, j);
                if (kl) t = fma(xj, Mm[j * K0 + lane], t);
            }
            const double qd = warp_sum(kl ? t * x2 : 0.0);
            const double quad = a.S2[size_t(cs) * n + i] - 2.0 * prow[cs * pstride + K0 + m] + sCmu[m];
            const double dt = a.stale_q ? sDt[M0 - 1] : sDt[m];
            const double ll = (-(quad - qd) - dt - a.ll_const) / 2.0 + sLp[m];
            if (lane == m) llv = ll;
        }
        const double mx = warp_max(llv);
        const double lik = ml ? exp(llv - mx) : 0.0;
        const double lsum = warp_sum(lik);
        double zz;
        if (a.update_z) {
            zz = ml ? lik / lsum : 0.0;
            if (ml) a.z[size_t(i) * M0 + lane] = zz;
        } else {
            zz = ml ? a.z[size_t(i) * M0 + lane] : 0.0;
        }
        const double zsq = sqrt(zz);
        nz_acc += zz;
        c_acc += zsq;
        tot_acc += log(lsum) + mx;

        // ---- membership draw for the coming sweep ----
        int im = 0;
        if (a.sample && a.draw_membership) {
            double u0, u1;
            cell_uniform_pair(a.seed, a.sweep, (unsigned long long)(a.d.cell_offset + i), 0u, u0, u1);
            double cum = 0.0;
            for (int m = 0; m < M0; ++m) {
                cum += __shfl_sync(0xffffffffu, zz, m);
                im += (u0 >= cum) ? 1 : 0;
            }
            im = min(im, M0 - 1);
        }

        // ---- pass 2: factor posterior means ----
        double wbar = 0.0, wsel = 0.0;
        for (int m = 0; m < M0; ++m) {
            const int cs = (NS == 1) ? 0 : m;
            const double xc = kl ? prow[cs * pstride + lane] - sOff[m * K0 + lane] : 0.0;
            double w = 0.0;
            const double* Mm = sMm + m * K0 * K0;
            for (int j = 0; j < K0; ++j) {
                const double xj = __shfl_sync(0xffffffffu, xc, j);
                if (kl) w = fma(xj, Mm[j * K0 + lane], w);
            }
            const double zm = __shfl_sync(0xffffffffu, zz, m);
            wbar = fma(zm, w, wbar);
            if (m == im) wsel = w;
            if (kl) {
                if (a.write_W) a.W[(size_t(m) * n + i) * K0 + lane] = w;
                if (a.write_Vg) a.Vg[size_t(i) * (K0 * M0 + 2 * M0) + m * K0 + lane] = zm * w;
            }
        }
        // ---- M-step operand row ----
        double* vrow = a.V + size_t(i) * a.d.VS;
        if (kl) vrow[lane] = wbar;
        if (ml) {
            vrow[K0 + lane] = zz;
            vrow[K0 + M0 + lane] = zsq;
            if (a.write_Vg) {
                double* vg = a.Vg + size_t(i) * (K0 * M0 + 2 * M0) + K0 * M0;
                vg[lane] = zz;
                vg[M0 + lane] = zsq;
            }
        }
        const double zs = warp_sum(zz);
        if (lane == 0) a.zsum[i] = zs;

        // ---- factor draw: f = W_im[i,] + eps Mh_im  (symmetric sqrt, as mixtools::rmvnorm) ----
        if (a.sample) {
            double ua, ub;
            cell_uniform_pair(a.seed, a.sweep, (unsigned long long)(a.d.cell_offset + i), unsigned(1 + lane) >> 1, ua, ub);
            const double eps = kl ? qnorm(((1 + lane) & 1) ? ub : ua) : 0.0;
            const double* Mh = a.cc.Mh + size_t(im) * K0 * K0;
            double f = wsel;
            for (int j = 0; j < K0; ++j) {
                const double ej = __shfl_sync(0xffffffffu, eps, j);
                if (kl) f = fma(ej, Mh[j * K0 + lane], f);
            }
            double* frow = a.Fh + size_t(i) * a.d.FS;
            if (kl) frow[lane] = f;
            if (ml) frow[K0 + lane] = (lane == im) ? 1.0 : 0.0;
            if (lane == 0) {
                frow[K0 + M0] = 1.0;
                a.im[i] = im;
            }
        }
    }

    // ---- deterministic CTA reduction of (nz, c, tot) ----
    const int R = 2 * M0 + 1;
    if (ml) {
        sRed[warp * R + lane] = nz_acc;
        sRed[warp * R + M0 + lane] = c_acc;
    }
    if (lane == 0) sRed[warp * R + 2 * M0] = tot_acc;
    __syncthreads();
    for (int c = threadIdx.x; c < R; c += blockDim.x) {
        double s = 0.0;
        for (int w = 0; w < CW; ++w) s += sRed[w * R + c];
        a.part[size_t(blockIdx.x) * R + c] = s;
    }
}

__global__ void k_finalize_cell(const double* part, int ncta, int R, double* out) {
    for (int c = threadIdx.x; c < R; c += blockDim.x) {
        double s = 0.0;
        for (int b = 0; b < ncta; ++b) s += part[size_t(b) * R + c];
        out[c] = s;
    }
}

}  // namespace

void launch_cell(const CellArgs& a, cudaStream_t st) {
    const int M0 = a.d.M0, K0 = a.d.K0;
    const size_t smem = sizeof(double) * (size_t(M0) * K0 * K0 + 2 * M0 * K0 + 3 * M0 + CW * (2 * M0 + 1));
    static size_t attr = 0;
    if (smem > 48 * 1024 && smem > attr) {
        cudaFuncSetAttribute(k_cell, cudaFuncAttributeMaxDynamicSharedMemorySize, int(smem));
        attr = smem;
    }
    k_cell<<<a.ncta, CW * 32, smem, st>>>(a);
}

void launch_finalize_cell(const double* part_cell, int ncta, int M0, double* out, cudaStream_t st) {
    k_finalize_cell<<<1, 64, 0, st>>>(part_cell, ncta, 2 * M0 + 1, out);
}

}  // namespace sb
