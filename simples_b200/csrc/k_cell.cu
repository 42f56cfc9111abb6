// cell_post: per-cell epilogue of an E-pass.  From the projections P = Y^T Q and S2 it forms the
// Woodbury Mahalanobis distance, log-likelihoods, responsibilities z, the log-lik total, the factor
// posterior means W_m = x2_m M_m (R/EM_impute.R:126-157), the M-step operand row V_i, and -- when an
// MC sweep follows -- the cell-level draws of that sweep (membership + factor vector,
// R/EM_impute.R:163-171).
//
// A warp owns 8 cells (one DMMA m-tile): T_m = X2_m M_m is a tiny FP64 DMMA product per cluster
// (K0 x K0, padded to 4 x 8 granules), so the K0^2 M0 work per cell runs on the tensor path instead
// of shuffle loops.  Deterministic per-CTA partial sums (nz, sum sqrt z, tot).
#include "ptx.cuh"
#include "rng.cuh"
#include "state.h"

namespace sb {
namespace {

constexpr int CW = 8;   // warps per CTA

struct CellSmem {
    double *Mp, *off, *offq, *cmu, *dt, *lp, *red;
    double *Pa, *Pb, *ll, *zz, *ws, *eps;   // per-warp
    int* im;
};

template <int KSX, int NTX>
__global__ void __launch_bounds__(CW * 32, (KSX <= 4 ? 4 : 2)) k_cell(CellArgs a, int mm_in_smem, int pw_doubles) {
    extern __shared__ __align__(16) double sm[];
    const int M0 = a.d.M0, K0 = a.d.K0, NS = a.d.NS, n = a.d.n, NQP = a.d.NQP;
    const int KS = (K0 + 3) / 4, NT = (K0 + 7) / 8, KR = 4 * KS, KPS = 8 * NT + 4;   // padded M_m: [KR][KPS]
    const int MMSZ = KR * KPS;
    // ---- shared memory carve-up ----
    double* p = sm;
    const double* sMp;
    if (mm_in_smem) {
        for (int i = threadIdx.x; i < M0 * MMSZ; i += blockDim.x) p[i] = a.cc.MmP[i];
        sMp = p;
        p += M0 * MMSZ;
    } else {
        sMp = a.cc.MmP;
    }
    double* sOff = p;      p += M0 * KR;     // [M0][KR] (zero padded)
    double* sOffq = p;     p += M0 * KR;
    double* sCmu = p;      p += M0;
    double* sDt = p;       p += M0;
    double* sLp = p;       p += M0;
    double* sRed = p;      p += CW * (2 * M0 + 1);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, g = lane >> 2, t = lane & 3;
    const int PWS = NQP + 4;
    double* Pa = p + warp * pw_doubles;                                      // [8][PWS]  x2 source for the quadratic form
    double* Pb = Pa + 8 * PWS;                                               // [8][PWS]  x2 source for W, mu column (general path only)
    double* llw = Pa + (NS == 1 ? 1 : 2) * 8 * PWS;                          // [8][M0]
    double* zw = llw + 8 * M0;                                               // [8][M0]
    double* sqw = zw + 8 * M0;                                               // [8][M0]   sqrt(z), formed once per (cell, cluster)
    double* wsw = sqw + 8 * M0;                                              // [8][KPS]  W of the sampled cluster
    double* epw = wsw + 8 * KPS;                                             // [8][KR]   factor normals
    p += CW * pw_doubles;
    int* imw = reinterpret_cast<int*>(p) + warp * 8;

    for (int i = threadIdx.x; i < M0 * KR; i += blockDim.x) {
        const int m = i / KR, k = i - m * KR;
        sOff[i] = k < K0 ? a.cc.off[m * K0 + k] : 0.0;
        sOffq[i] = k < K0 ? a.cc.offq[m * K0 + k] : 0.0;
    }
    for (int i = threadIdx.x; i < M0; i += blockDim.x) {
        sCmu[i] = a.cc.cmu[i];
        sDt[i] = a.cc.dt[i];
        sLp[i] = a.cc.logpi[i];
    }
    __syncthreads();

    const bool ml = lane < M0;
    double nz_acc = 0.0, c_acc = 0.0, tot_acc = 0.0;
    const size_t pstride = size_t(n) * NQP;
    const int ngroups = (n + 7) / 8;
    const int VGS = K0 * M0 + 2 * M0;

    for (int grp = blockIdx.x * CW + warp; grp < ngroups; grp += gridDim.x * CW) {
        const int cell0 = grp * 8;
        const int cell = cell0 + g;             // this lane's row
        const bool rv = cell < n;
        // ---- stage the P rows of the 8 cells (shared-sigma path: one tile serves everything) ----
        if (NS == 1) {
            // the 8 rows are contiguous in P; (row, column) of the padded tile advance without a division
            for (int idx = lane, r = lane / NQP, c = lane - (lane / NQP) * NQP; idx < 8 * NQP; idx += 32) {
                Pa[r * PWS + c] = (cell0 + r < n) ? a.P[size_t(cell0) * NQP + idx] : 0.0;
                c += 32;
                while (c >= NQP) {
                    c -= NQP;
                    ++r;
                }
            }
        }
        const double* PbT = (NS == 1) ? Pa : Pb;
        __syncwarp();

        // ---- pass 1: log-likelihoods ----
        for (int m = 0; m < M0; ++m) {
            const int cs = (NS == 1) ? 0 : m;
            const int qs = a.stale_q ? NS - 1 : cs;
            if (NS > 1) {
                __syncwarp();
                for (int idx = lane; idx < 8 * NQP; idx += 32) {
                    const int r = idx / NQP, c = idx - r * NQP;
                    const bool v = cell0 + r < n;
                    Pa[r * PWS + c] = v ? a.P[qs * pstride + size_t(cell0 + r) * NQP + c] : 0.0;
                    Pb[r * PWS + c] = v ? a.P[cs * pstride + size_t(cell0 + r) * NQP + c] : 0.0;
                }
                __syncwarp();
            }
            const double* off = (a.stale_q ? sOffq : sOff) + m * KR;
            const double* Mp = sMp + m * MMSZ;
            double av[KSX];
#pragma unroll
            for (int ks = 0; ks < KSX; ++ks) {
                const int k = 4 * ks + t;
                av[ks] = (ks < KS && k < K0) ? Pa[g * PWS + k] - off[k] : 0.0;
            }
            double qd = 0.0;
#pragma unroll
            for (int nt = 0; nt < NTX; ++nt) {
                if (nt < NT) {
                    double c0 = 0.0, c1 = 0.0;
#pragma unroll
                    for (int ks = 0; ks < KSX; ++ks)
                        if (ks < KS) dmma(c0, c1, av[ks], Mp[(4 * ks + t) * KPS + nt * 8 + g]);
                    const int k0 = nt * 8 + 2 * t;
                    const double x0 = (k0 < K0) ? Pa[g * PWS + k0] - off[k0] : 0.0;
                    const double x1 = (k0 + 1 < K0) ? Pa[g * PWS + k0 + 1] - off[k0 + 1] : 0.0;
                    qd = fma(c0, x0, fma(c1, x1, qd));
                }
            }
            qd += __shfl_xor_sync(0xffffffffu, qd, 1);
            qd += __shfl_xor_sync(0xffffffffu, qd, 2);
            if (t == 0) {
                const double s2 = rv ? a.S2[size_t(cs) * n + cell] : 0.0;
                const double quad = s2 - 2.0 * PbT[g * PWS + K0 + m] + sCmu[m];
                const double dt = a.stale_q ? sDt[M0 - 1] : sDt[m];
                llw[g * M0 + m] = (-(quad - qd) - dt - a.ll_const) / 2.0 + sLp[m];
            }
        }
        __syncwarp();

        // ---- responsibilities: lane (g, t) covers clusters t, t+4, ... of cell g ----
        double mx = -INFINITY;
        for (int m = t; m < M0; m += 4) mx = fmax(mx, llw[g * M0 + m]);
        mx = fmax(mx, __shfl_xor_sync(0xffffffffu, mx, 1));
        mx = fmax(mx, __shfl_xor_sync(0xffffffffu, mx, 2));
        double ls = 0.0;
        for (int m = t; m < M0; m += 4) {
            const double e = exp_neg(llw[g * M0 + m] - mx);
            zw[g * M0 + m] = e;
            ls += e;
        }
        // sum over the 4 lanes of the row in ascending-cluster order is not required for parity (1e-16)
        ls += __shfl_xor_sync(0xffffffffu, ls, 1);
        ls += __shfl_xor_sync(0xffffffffu, ls, 2);
        for (int m = t; m < M0; m += 4) {
            double z;
            if (a.update_z) {
                z = zw[g * M0 + m] / ls;
                if (rv) a.z[size_t(cell) * M0 + m] = z;
            } else {
                z = rv ? a.z[size_t(cell) * M0 + m] : 0.0;
            }
            z = rv ? z : 0.0;
            zw[g * M0 + m] = z;
            sqw[g * M0 + m] = sqrt_unit(z);
        }
        if (t == 0 && rv) tot_acc += log(ls) + mx;
        __syncwarp();
        if (ml) {
            double s = 0.0, c = 0.0;
            for (int r = 0; r < 8; ++r) {
                s += zw[r * M0 + lane];
                c += sqw[r * M0 + lane];
            }
            nz_acc += s;
            c_acc += c;
        }

        // ---- membership draw for the coming sweep (lanes 0..7: one cell each) ----
        if (a.sample) {
            if (lane < 8) {
                int im = 0;
                if (a.expect) {                 // most probable cluster (first maximum)
                    double best = zw[lane * M0];
                    for (int m = 1; m < M0; ++m)
                        if (zw[lane * M0 + m] > best) {
                            best = zw[lane * M0 + m];
                            im = m;
                        }
                } else if (a.draw_membership) {
                    double u0, u1;
                    cell_uniform_pair(a.seed, a.sweep, (unsigned long long)(a.d.cell_offset + cell0 + lane), 0u, u0, u1);
                    double cum = 0.0;
                    for (int m = 0; m < M0; ++m) {
                        cum += zw[lane * M0 + m];
                        im += (u0 >= cum) ? 1 : 0;
                    }
                    im = min(im, M0 - 1);
                }
                imw[lane] = im;
                if (cell0 + lane < n) a.im[cell0 + lane] = im;
            }
            __syncwarp();
        }
        const int im_g = a.sample ? imw[g] : -1;

        // ---- pass 2: factor posterior means W_m = x2_m M_m, Wbar = sum_m z_m W_m ----
        double wb[NTX][2];
#pragma unroll
        for (int nt = 0; nt < NTX; ++nt) wb[nt][0] = wb[nt][1] = 0.0;
        for (int m = 0; m < M0; ++m) {
            const int cs = (NS == 1) ? 0 : m;
            if (NS > 1) {
                __syncwarp();
                for (int idx = lane; idx < 8 * NQP; idx += 32) {
                    const int r = idx / NQP, c = idx - r * NQP;
                    Pb[r * PWS + c] = (cell0 + r < n) ? a.P[cs * pstride + size_t(cell0 + r) * NQP + c] : 0.0;
                }
                __syncwarp();
            }
            const double* off = sOff + m * KR;
            const double* Mp = sMp + m * MMSZ;
            const double zm = zw[g * M0 + m];
            double av[KSX];
#pragma unroll
            for (int ks = 0; ks < KSX; ++ks) {
                const int k = 4 * ks + t;
                av[ks] = (ks < KS && k < K0) ? PbT[g * PWS + k] - off[k] : 0.0;
            }
#pragma unroll
            for (int nt = 0; nt < NTX; ++nt) {
                if (nt < NT) {
                    double c0 = 0.0, c1 = 0.0;
#pragma unroll
                    for (int ks = 0; ks < KSX; ++ks)
                        if (ks < KS) dmma(c0, c1, av[ks], Mp[(4 * ks + t) * KPS + nt * 8 + g]);
                    wb[nt][0] = fma(zm, c0, wb[nt][0]);
                    wb[nt][1] = fma(zm, c1, wb[nt][1]);
                    const int k0 = nt * 8 + 2 * t;
                    if (m == im_g) {
                        wsw[g * KPS + k0] = c0;
                        wsw[g * KPS + k0 + 1] = c1;
                    }
                    if (rv) {
                        if (a.write_W) {
                            double* wr = a.W + (size_t(m) * n + cell) * K0;
                            if (k0 < K0) wr[k0] = c0;
                            if (k0 + 1 < K0) wr[k0 + 1] = c1;
                        }
                        if (a.write_Vg) {
                            double* vg = a.Vg + size_t(cell) * VGS + m * K0;
                            if (k0 < K0) vg[k0] = zm * c0;
                            if (k0 + 1 < K0) vg[k0 + 1] = zm * c1;
                        }
                    }
                }
            }
        }
        // ---- M-step operand row ----
        if (rv) {
            double* vrow = a.V + size_t(cell) * a.d.VS;
#pragma unroll
            for (int nt = 0; nt < NTX; ++nt) {
                const int k0 = nt * 8 + 2 * t;
                if (nt < NT) {
                    if (k0 < K0) vrow[k0] = wb[nt][0];
                    if (k0 + 1 < K0) vrow[k0 + 1] = wb[nt][1];
                }
            }
            double zs = 0.0, zq = 0.0;
            for (int m = t; m < M0; m += 4) {
                const double z = zw[g * M0 + m];
                vrow[K0 + m] = z;
                if (a.write_Vg) {
                    double* vg = a.Vg + size_t(cell) * VGS + K0 * M0;
                    vg[m] = z;
                    vg[M0 + m] = sqw[g * M0 + m];
                }
            }
            for (int m = 0; m < M0; ++m) {
                zs += zw[g * M0 + m];
                zq += sqw[g * M0 + m];
            }
            if (t == 0) {
                a.zsum[cell] = zs;
                vrow[K0 + M0] = zq;        // sum_m sqrt(z_im): with cluster-shared sigma only sum_m Uh[,m] is needed
            }
        }

        // ---- factor draw: f = W_im[i,] + eps Mh_im  (symmetric sqrt, as mixtools::rmvnorm) ----
        if (a.sample) {
            __syncwarp();
            for (int id = lane; id < 8 * K0; id += 32) {
                const int r = id / K0, k = id - r * K0;
                double ua, ub;
                if (a.expect) {
                    epw[r * KR + k] = 0.0;      // factor at its posterior mean
                    continue;
                }
                cell_uniform_pair(a.seed, a.sweep, (unsigned long long)(a.d.cell_offset + cell0 + r), unsigned(1 + k) >> 1, ua, ub);
                epw[r * KR + k] = qnorm(((1 + k) & 1) ? ub : ua);
            }
            __syncwarp();
            for (int id = lane; id < 8 * K0; id += 32) {
                const int r = id / K0, k = id - r * K0;
                if (cell0 + r < n) {
                    const double* Mh = a.cc.Mh + size_t(imw[r]) * K0 * K0;
                    double f = wsw[r * KPS + k];
                    for (int j = 0; j < K0; ++j) f = fma(epw[r * KR + j], Mh[j * K0 + k], f);
                    a.Fh[size_t(cell0 + r) * a.d.FS + k] = f;
                }
            }
        }
        __syncwarp();
    }

    // ---- deterministic CTA reduction of (nz, c, tot) ----
    tot_acc = warp_sum(tot_acc);
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

// Cluster-shared sigma (NS == 1, the default model): the same computation with every tile dimension a compile-time
// constant (KS k-steps: KR = 4 KS rows, NT = ceil(KS / 2) n-tiles, KPS = 8 NT + 4) and no `k < K0` predicates -- the padded
// rows / columns of M_m are zero, so whatever finite value the padded operand positions hold (the mu columns of P) drops
// out.  What depends only on the cell (its P row fragments) is read once per group instead of once per cluster, the M_m
// fragments come through LDG with immediate offsets, the offset rows are padded to 8 NT so that the quadratic form reads
// them as 16-byte pairs.  Same arithmetic, in the same order, as k_cell; ~190 -> ~55 instructions per cluster in pass 1.
template <int KS>
__global__ void __launch_bounds__(CW * 32, (KS <= 4 ? 4 : 2)) k_cell_fast(CellArgs a, int pw_doubles) {
    constexpr int NT = (KS + 1) / 2, KR = 4 * KS, KPS = 8 * NT + 4, MMSZ = KR * KPS, OS = 8 * NT;
    extern __shared__ __align__(16) double sm[];
    const int M0 = a.d.M0, K0 = a.d.K0, n = a.d.n, NQP = a.d.NQP;
    double* p = sm;
    double* sOff = p;      p += M0 * OS;     // [M0][OS] (zero padded)
    double* sOffq = p;     p += M0 * OS;
    double* sCmu = p;      p += M0;
    double* sDt = p;       p += M0;
    double* sLp = p;       p += M0;
    double* sRed = p;      p += CW * (2 * M0 + 1);
    p += (reinterpret_cast<uintptr_t>(p) & 8) ? 1 : 0;        // 16-byte alignment of the per-warp tiles
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, g = lane >> 2, t = lane & 3;
    const int PWS = NQP + 4;
    double* Pa = p + warp * pw_doubles;                                      // [8][PWS]
    double* llw = Pa + 8 * PWS;                                              // [8][M0]
    double* zw = llw + 8 * M0;                                               // [8][M0]
    double* sqw = zw + 8 * M0;                                               // [8][M0]   sqrt(z)
    double* wsw = sqw + 8 * M0;                                              // [8][KPS]  W of the sampled cluster
    double* epw = wsw + 8 * KPS;                                             // [8][KR]   factor normals
    p += CW * pw_doubles;
    int* imw = reinterpret_cast<int*>(p) + warp * 8;

    for (int i = threadIdx.x; i < M0 * OS; i += blockDim.x) {
        const int m = i / OS, k = i - m * OS;
        sOff[i] = k < K0 ? a.cc.off[m * K0 + k] : 0.0;
        sOffq[i] = k < K0 ? a.cc.offq[m * K0 + k] : 0.0;
    }
    for (int i = threadIdx.x; i < M0; i += blockDim.x) {
        sCmu[i] = a.cc.cmu[i];
        sDt[i] = a.cc.dt[i];
        sLp[i] = a.cc.logpi[i];
    }
    __syncthreads();

    const bool ml = lane < M0;
    double nz_acc = 0.0, c_acc = 0.0, tot_acc = 0.0;
    const int ngroups = (n + 7) / 8;
    const int VGS = K0 * M0 + 2 * M0;
    const double* off1 = a.stale_q ? sOffq : sOff;                           // pass-1 offsets (Q1: stale beta2 in MC passes)
    // this lane's B-fragment corner of M_0.  The padded M_m tables (19 KB at K0 = M0 = 10) stay in global memory, read with
    // an L1 evict-last policy; a shared-memory copy per CTA was measured equal (0.0999 vs 0.0997 ms) and costs a resident CTA
    const double* MpL = a.cc.MmP + t * KPS + g;
    const int lane_r0 = lane / NQP, lane_c0 = lane - lane_r0 * NQP;
    bool kv[NT][2];                                                          // column < K0 (global stores only)
#pragma unroll
    for (int nt = 0; nt < NT; ++nt) {
        kv[nt][0] = nt * 8 + 2 * t < K0;
        kv[nt][1] = nt * 8 + 2 * t + 1 < K0;
    }

    for (int grp = blockIdx.x * CW + warp; grp < ngroups; grp += gridDim.x * CW) {
        const int cell0 = grp * 8;
        const int cell = cell0 + g;             // this lane's row
        const bool rv = cell < n;
        // ---- stage the P rows of the 8 cells (contiguous in P; (row, column) of the padded tile advance without a division) ----
        for (int idx = lane, r = lane_r0, c = lane_c0; idx < 8 * NQP; idx += 32) {
            Pa[r * PWS + c] = (cell0 + r < n) ? ldg_stream(a.P + size_t(cell0) * NQP + idx) : 0.0;
            c += 32;
            while (c >= NQP) {
                c -= NQP;
                ++r;
            }
        }
        const double s2v = (t == 0 && rv) ? a.S2[cell] : 0.0;
        __syncwarp();
        const double* prow = Pa + g * PWS;
        double pa[KS], px[NT][2];               // the cell's A-fragment and C-fragment columns of x = P[, 0:K0)
#pragma unroll
        for (int ks = 0; ks < KS; ++ks) pa[ks] = prow[4 * ks + t];
#pragma unroll
        for (int nt = 0; nt < NT; ++nt) {
            const double2 v = *reinterpret_cast<const double2*>(prow + nt * 8 + 2 * t);
            px[nt][0] = v.x;
            px[nt][1] = v.y;
        }

        // ---- pass 1: log-likelihoods ----
        {
            const double* off = off1;
            const double* Mp = MpL;
            for (int m = 0; m < M0; ++m, off += OS, Mp += MMSZ) {
                double av[KS];
#pragma unroll
                for (int ks = 0; ks < KS; ++ks) av[ks] = pa[ks] - off[4 * ks + t];
                double qd = 0.0;
#pragma unroll
                for (int nt = 0; nt < NT; ++nt) {
                    double c0 = 0.0, c1 = 0.0;
#pragma unroll
                    for (int ks = 0; ks < KS; ++ks) dmma(c0, c1, av[ks], ldg_keep(Mp + 4 * ks * KPS + nt * 8));
                    const double2 o2 = *reinterpret_cast<const double2*>(off + nt * 8 + 2 * t);
                    qd = fma(c0, px[nt][0] - o2.x, fma(c1, px[nt][1] - o2.y, qd));
                }
                qd += __shfl_xor_sync(0xffffffffu, qd, 1);
                qd += __shfl_xor_sync(0xffffffffu, qd, 2);
                if (t == 0) {
                    const double quad = s2v - 2.0 * prow[K0 + m] + sCmu[m];
                    const double dt = a.stale_q ? sDt[M0 - 1] : sDt[m];
                    llw[g * M0 + m] = (-(quad - qd) - dt - a.ll_const) / 2.0 + sLp[m];
                }
            }
        }
        __syncwarp();

        // ---- responsibilities: lane (g, t) covers clusters t, t+4, ... of cell g ----
        double mx = -INFINITY;
        for (int m = t; m < M0; m += 4) mx = fmax(mx, llw[g * M0 + m]);
        mx = fmax(mx, __shfl_xor_sync(0xffffffffu, mx, 1));
        mx = fmax(mx, __shfl_xor_sync(0xffffffffu, mx, 2));
        double ls = 0.0;
        for (int m = t; m < M0; m += 4) {
            const double e = exp_neg(llw[g * M0 + m] - mx);
            zw[g * M0 + m] = e;
            ls += e;
        }
        ls += __shfl_xor_sync(0xffffffffu, ls, 1);
        ls += __shfl_xor_sync(0xffffffffu, ls, 2);
        for (int m = t; m < M0; m += 4) {
            double z;
            if (a.update_z) {
                z = zw[g * M0 + m] / ls;
                if (rv) a.z[size_t(cell) * M0 + m] = z;
            } else {
                z = rv ? a.z[size_t(cell) * M0 + m] : 0.0;
            }
            z = rv ? z : 0.0;
            zw[g * M0 + m] = z;
            sqw[g * M0 + m] = sqrt_unit(z);
        }
        if (t == 0 && rv) tot_acc += log(ls) + mx;
        __syncwarp();
        if (ml) {
            double s = 0.0, c = 0.0;
            for (int r = 0; r < 8; ++r) {
                s += zw[r * M0 + lane];
                c += sqw[r * M0 + lane];
            }
            nz_acc += s;
            c_acc += c;
        }

        // ---- membership draw for the coming sweep (lanes 0..7: one cell each) ----
        if (a.sample) {
            if (lane < 8) {
                int im = 0;
                if (a.expect) {                 // most probable cluster (first maximum)
                    double best = zw[lane * M0];
                    for (int m = 1; m < M0; ++m)
                        if (zw[lane * M0 + m] > best) {
                            best = zw[lane * M0 + m];
                            im = m;
                        }
                } else if (a.draw_membership) {
                    double u0, u1;
                    cell_uniform_pair(a.seed, a.sweep, (unsigned long long)(a.d.cell_offset + cell0 + lane), 0u, u0, u1);
                    double cum = 0.0;
                    for (int m = 0; m < M0; ++m) {
                        cum += zw[lane * M0 + m];
                        im += (u0 >= cum) ? 1 : 0;
                    }
                    im = min(im, M0 - 1);
                }
                imw[lane] = im;
                if (cell0 + lane < n) a.im[cell0 + lane] = im;
            }
            __syncwarp();
        }
        const int im_g = a.sample ? imw[g] : -1;

        // ---- pass 2: factor posterior means W_m = x2_m M_m, Wbar = sum_m z_m W_m ----
        double wb[NT][2];
#pragma unroll
        for (int nt = 0; nt < NT; ++nt) wb[nt][0] = wb[nt][1] = 0.0;
        {
            const double* off = sOff;
            const double* Mp = MpL;
            double* wr = a.W + size_t(cell) * K0 + 2 * t;                    // row of cluster 0; clusters are n K0 apart
            double* vg = a.Vg + size_t(cell) * VGS + 2 * t;
            for (int m = 0; m < M0; ++m, off += OS, Mp += MMSZ, wr += size_t(n) * K0, vg += K0) {
                const double zm = zw[g * M0 + m];
                double av[KS];
#pragma unroll
                for (int ks = 0; ks < KS; ++ks) av[ks] = pa[ks] - off[4 * ks + t];
#pragma unroll
                for (int nt = 0; nt < NT; ++nt) {
                    double c0 = 0.0, c1 = 0.0;
#pragma unroll
                    for (int ks = 0; ks < KS; ++ks) dmma(c0, c1, av[ks], ldg_keep(Mp + 4 * ks * KPS + nt * 8));
                    wb[nt][0] = fma(zm, c0, wb[nt][0]);
                    wb[nt][1] = fma(zm, c1, wb[nt][1]);
                    if (m == im_g) *reinterpret_cast<double2*>(wsw + g * KPS + nt * 8 + 2 * t) = make_double2(c0, c1);
                    if (rv) {
                        if (a.write_W) {
                            if (kv[nt][0]) __stcs(wr + nt * 8, c0);
                            if (kv[nt][1]) __stcs(wr + nt * 8 + 1, c1);
                        }
                        if (a.write_Vg) {
                            if (kv[nt][0]) __stcs(vg + nt * 8, zm * c0);
                            if (kv[nt][1]) __stcs(vg + nt * 8 + 1, zm * c1);
                        }
                    }
                }
            }
        }
        // ---- M-step operand row ----
        if (rv) {
            double* vrow = a.V + size_t(cell) * a.d.VS;
#pragma unroll
            for (int nt = 0; nt < NT; ++nt) {
                if (kv[nt][0]) vrow[nt * 8 + 2 * t] = wb[nt][0];
                if (kv[nt][1]) vrow[nt * 8 + 2 * t + 1] = wb[nt][1];
            }
            double zs = 0.0, zq = 0.0;
            for (int m = t; m < M0; m += 4) {
                const double z = zw[g * M0 + m];
                vrow[K0 + m] = z;
                if (a.write_Vg) {
                    double* vgz = a.Vg + size_t(cell) * VGS + K0 * M0;
                    vgz[m] = z;
                    vgz[M0 + m] = sqw[g * M0 + m];
                }
            }
            for (int m = 0; m < M0; ++m) {
                zs += zw[g * M0 + m];
                zq += sqw[g * M0 + m];
            }
            if (t == 0) {
                a.zsum[cell] = zs;
                vrow[K0 + M0] = zq;        // sum_m sqrt(z_im): with cluster-shared sigma only sum_m Uh[,m] is needed
            }
        }

        // ---- factor draw: f = W_im[i,] + eps Mh_im  (symmetric sqrt, as mixtools::rmvnorm) ----
        if (a.sample) {
            __syncwarp();
            for (int id = lane; id < 8 * K0; id += 32) {
                const int r = id / K0, k = id - r * K0;
                double ua, ub;
                if (a.expect) {
                    epw[r * KR + k] = 0.0;      // factor at its posterior mean
                    continue;
                }
                cell_uniform_pair(a.seed, a.sweep, (unsigned long long)(a.d.cell_offset + cell0 + r), unsigned(1 + k) >> 1, ua, ub);
                epw[r * KR + k] = qnorm(((1 + k) & 1) ? ub : ua);
            }
            __syncwarp();
            for (int id = lane; id < 8 * K0; id += 32) {
                const int r = id / K0, k = id - r * K0;
                if (cell0 + r < n) {
                    const double* Mh = a.cc.Mh + size_t(imw[r]) * K0 * K0;
                    double f = wsw[r * KPS + k];
                    for (int j = 0; j < K0; ++j) f = fma(epw[r * KR + j], Mh[j * K0 + k], f);
                    a.Fh[size_t(cell0 + r) * a.d.FS + k] = f;
                }
            }
        }
        __syncwarp();
    }

    // ---- deterministic CTA reduction of (nz, c, tot) ----
    tot_acc = warp_sum(tot_acc);
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

// one warp per column: lane-strided partial sums then a fixed shuffle tree (deterministic)
__global__ void k_finalize_cell(const double* part, int ncta, int R, double* out) {
    const int c = blockIdx.x, lane = threadIdx.x;
    double s = 0.0;
    for (int b = lane; b < ncta; b += 32) s += part[size_t(b) * R + c];
    s = warp_sum(s);
    if (lane == 0) out[c] = s;
}

}  // namespace

template <int KS>
static void launch_cell_fast(const CellArgs& a, cudaStream_t st) {
    constexpr int NT = (KS + 1) / 2, KR = 4 * KS, KPS = 8 * NT + 4, OS = 8 * NT;
    const int M0 = a.d.M0, PWS = a.d.NQP + 4;
    int pw_doubles = 8 * PWS + 8 * M0 * 3 + 8 * KPS + 8 * KR;
    pw_doubles += pw_doubles & 1;                                     // keep every warp's tiles 16-byte aligned
    const size_t smem = sizeof(double) * (size_t(2) * M0 * OS + 3 * M0 + CW * (2 * M0 + 1) + 1 + size_t(CW) * pw_doubles) +
                        sizeof(int) * CW * 8;
    ensure_dyn_smem_k(k_cell_fast<KS>, smem);
    k_cell_fast<KS><<<a.ncta, CW * 32, smem, st>>>(a, pw_doubles);
}

void launch_cell(const CellArgs& a, cudaStream_t st) {
    const int M0 = a.d.M0, K0 = a.d.K0;
    if (a.d.NS == 1) {      // cluster-shared sigma: fully specialised kernel
        switch ((K0 + 3) / 4) {
            case 1: launch_cell_fast<1>(a, st); return;
            case 2: launch_cell_fast<2>(a, st); return;
            case 3: launch_cell_fast<3>(a, st); return;
            case 4: launch_cell_fast<4>(a, st); return;
            case 5: launch_cell_fast<5>(a, st); return;
            case 6: launch_cell_fast<6>(a, st); return;
            case 7: launch_cell_fast<7>(a, st); return;
            case 8: launch_cell_fast<8>(a, st); return;
            default: break;
        }
    }
    const int KS = (K0 + 3) / 4, NT = (K0 + 7) / 8, KR = 4 * KS, KPS = 8 * NT + 4, PWS = a.d.NQP + 4;
    const int pw_doubles = (a.d.NS == 1 ? 1 : 2) * 8 * PWS + 8 * M0 * 3 + 8 * KPS + 8 * KR;
    const size_t fixed = sizeof(double) * (size_t(2) * M0 * KR + 3 * M0 + CW * (2 * M0 + 1) + size_t(CW) * pw_doubles) +
                         sizeof(int) * CW * 8;
    // The padded M_m tables (M0 * KR * KPS doubles, 19 KB at K0 = M0 = 10) are read through L1 instead of being
    // copied to shared memory by every CTA: same latency, and the smaller footprint lets 4-5 CTAs share an SM.
    const int mm_in_smem = 0;
    const size_t smem = fixed;
    if (K0 <= 16) {
        ensure_dyn_smem_k(k_cell<4, 2>, smem);
        k_cell<4, 2><<<a.ncta, CW * 32, smem, st>>>(a, mm_in_smem, pw_doubles);
    } else {
        ensure_dyn_smem_k(k_cell<8, 4>, smem);
        k_cell<8, 4><<<a.ncta, CW * 32, smem, st>>>(a, mm_in_smem, pw_doubles);
    }
}

void launch_finalize_cell(const double* part_cell, int ncta, int M0, double* out, cudaStream_t st) {
    k_finalize_cell<<<2 * M0 + 1, 32, 0, st>>>(part_cell, ncta, 2 * M0 + 1, out);
}

}  // namespace sb
