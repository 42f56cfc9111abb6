// Initial clustering of the cells, R/SIMPLE.R:258-266 (= R/SIMPLE_B.R:293-302):
//   Y2_scale = t(scale(t(X)));  s = irlba(Y2_scale, nv = K);  cellmat = s$v %*% diag(s$d);  kmeans(cellmat, M0, nstart)
// B200 formulation: the gene Gram matrix C = Ys Ys' (G x G, FP64 DMMA through mstep_stats in column blocks) ->
// top-K eigenvectors by orthogonal (subspace) iteration + Rayleigh-Ritz on the small matrix -> cellmat = Ys' U
// (estep_project) -> Lloyd k-means with `nstart` restarts advanced together.  Every reduction has a fixed order.
#include "linalg.cuh"
#include "ptx.cuh"
#include "rng.cuh"
#include "state.h"

namespace sb {
namespace {

// V[i][c] = Y[i][g0 + c] (c < ncol, zero beyond G), the cell-major operand block of mstep_stats
__global__ void k_extract_cols(const double* Y, double* V, int n, int ldG, int G, int g0, int ncol, int VS) {
    const size_t e = blockIdx.x * size_t(blockDim.x) + threadIdx.x;
    if (e >= size_t(n) * VS) return;
    const int i = int(e / VS), c = int(e - size_t(i) * VS);
    V[e] = (c < ncol && g0 + c < G) ? Y[size_t(i) * ldG + g0 + c] : 0.0;
}

// C[g][g0 + c] = blk[g][c]   (C has `rows` rows of stride ldc)
__global__ void k_scatter_block(const double* blk, double* C, int rows, int ldc, int NVP, int g0, int ncol) {
    const int e = blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= rows * ncol) return;
    const int g = e / ncol, c = e - g * ncol;
    if (g0 + c < ldc) C[size_t(g) * ldc + g0 + c] = blk[size_t(g) * NVP + c];
}

// sd_g = sqrt(ss_g / (n - 1)) (0 -> 1 so that constant genes scale to 0 instead of NaN), mean <- 0 for the 2nd centring pass
__global__ void k_sd_from_ss(const double* ss, double* sd, double* zero, int G, int ldG, double nm1) {
    const int g = blockIdx.x * blockDim.x + threadIdx.x;
    if (g >= ldG) return;
    const double v = g < G ? sqrt(ss[g] / nm1) : 1.0;
    sd[g] = v > 0.0 ? v : 1.0;
    zero[g] = 0.0;
}

// Q0: deterministic pseudo-random start of the subspace iteration
__global__ void k_sub_init(double* Q, int G, int ldG, int KP, unsigned long long seed) {
    const int e = blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= ldG * KP) return;
    const int g = e / KP, c = e - g * KP;
    double u0, u1;
    cell_uniform_pair(seed, 0x5eedu, (unsigned long long)g, (unsigned)c, u0, u1);
    Q[e] = g < G ? qnorm(u0) : 0.0;
}

// Z = C Q   (C [ldG][ldG] symmetric, Q [ldG][KP]); 8 rows x 32 columns per CTA
__global__ void __launch_bounds__(256) k_sub_mult(const double* __restrict__ C, const double* __restrict__ Q, double* Z,
                                                  int G, int ldG, int KP) {
    const int c = threadIdx.x & 31, g = blockIdx.x * 8 + (threadIdx.x >> 5);
    if (g >= ldG) return;
    double s = 0.0;
    if (c < KP && g < G) {
        const double* cr = C + size_t(g) * ldG;
        for (int h = 0; h < G; ++h) s = fma(cr[h], Q[size_t(h) * KP + c], s);
    }
    if (c < KP) Z[size_t(g) * KP + c] = s;
}

// Cholesky-QR of Z (G x KP): Gm = Z'Z, L = chol(Gm), Q = Z L^-T.  One CTA of 1024 threads.
__global__ void __launch_bounds__(1024) k_sub_orth(const double* Z, double* Q, int G, int KP) {
    __shared__ double Gm[32 * 32], Li[32 * 32];
    const int tid = threadIdx.x;
    for (int e = tid; e < KP * KP; e += 1024) {
        const int a = e / KP, b = e - a * KP;
        double s = 0.0;
        if (b <= a)
            for (int g = 0; g < G; ++g) s = fma(Z[size_t(g) * KP + a], Z[size_t(g) * KP + b], s);
        Gm[e] = s;
    }
    __syncthreads();
    if (tid < 32) {
        for (int j = 0; j < KP; ++j) {      // Cholesky with a floor on the pivots (rank-deficient starts stay finite)
            __syncwarp();
            double d = Gm[j * KP + j];
            for (int p = 0; p < j; ++p) d -= Gm[j * KP + p] * Gm[j * KP + p];
            d = sqrt(fmax(d, 1e-300));
            __syncwarp();
            if (tid == 0) Gm[j * KP + j] = d;
            const int i = j + 1 + tid;
            if (i < KP) {
                double s = Gm[i * KP + j];
                for (int p = 0; p < j; ++p) s -= Gm[i * KP + p] * Gm[j * KP + p];
                Gm[i * KP + j] = s / d;
            }
        }
        __syncwarp();
        warp_tri_inverse(Gm, Li, KP, tid);
    }
    __syncthreads();
    for (int e = tid; e < G * KP; e += 1024) {
        const int g = e / KP, c = e - g * KP;
        double s = 0.0;
        for (int p = 0; p <= c; ++p) s = fma(Z[size_t(g) * KP + p], Li[c * KP + p], s);
        Q[e] = s;
    }
}

// Rayleigh-Ritz: T = Q'(C Q) -> Jacobi eigen-decomposition -> the K leading Ritz vectors as the projection operand
// A [ldG][QS] (zero padded) and the singular values d_k = sqrt(lambda_k).
__global__ void __launch_bounds__(1024) k_sub_ritz(const double* Q, const double* Z, double* A, double* dvals, int G,
                                                   int ldG, int KP, int K, int QS) {
    __shared__ double T[32 * 32], V[32 * 32];
    __shared__ int order[32];
    const int tid = threadIdx.x;
    for (int e = tid; e < KP * KP; e += 1024) {
        const int a = e / KP, b = e - a * KP;
        double s = 0.0;
        for (int g = 0; g < G; ++g) s = fma(Q[size_t(g) * KP + a], Z[size_t(g) * KP + b], s);
        T[e] = s;
    }
    __syncthreads();
    for (int e = tid; e < KP * KP; e += 1024) {
        const int a = e / KP, b = e - a * KP;
        if (b < a) {
            const double s = 0.5 * (T[a * KP + b] + T[b * KP + a]);
            V[a * KP + b] = s;                     // V as scratch for the symmetrised copy
            V[b * KP + a] = s;
        } else if (a == b) {
            V[e] = T[e];
        }
    }
    __syncthreads();
    for (int e = tid; e < KP * KP; e += 1024) T[e] = V[e];
    __syncthreads();
    if (tid < 32) {
        const int lane = tid;
        for (int e = lane; e < KP * KP; e += 32) V[e] = ((e / KP) == (e % KP)) ? 1.0 : 0.0;
        __syncwarp();
        for (int sweep = 0; sweep < 40; ++sweep) {          // cyclic Jacobi (same rotations as warp_sym_sqrt)
            int nrot = 0;
            for (int p = 0; p < KP - 1; ++p) {
                for (int q = p + 1; q < KP; ++q) {
                    const double apq = T[p * KP + q], app = T[p * KP + p], aqq = T[q * KP + q];
                    __syncwarp();
                    if (fabs(apq) <= 4e-16 * sqrt(fabs(app * aqq)) || fabs(apq) <= 1e-300) continue;
                    ++nrot;
                    const double theta = (aqq - app) / (2.0 * apq);
                    const double tt = (theta >= 0.0 ? 1.0 : -1.0) / (fabs(theta) + sqrt(theta * theta + 1.0));
                    const double c = 1.0 / sqrt(tt * tt + 1.0), s = tt * c;
                    if (lane < KP) {
                        const int i = lane;
                        const double aip = T[i * KP + p], aiq = T[i * KP + q];
                        T[i * KP + p] = c * aip - s * aiq;
                        T[i * KP + q] = s * aip + c * aiq;
                        const double vip = V[i * KP + p], viq = V[i * KP + q];
                        V[i * KP + p] = c * vip - s * viq;
                        V[i * KP + q] = s * vip + c * viq;
                    }
                    __syncwarp();
                    if (lane < KP) {
                        const int j = lane;
                        const double apj = T[p * KP + j], aqj = T[q * KP + j];
                        T[p * KP + j] = c * apj - s * aqj;
                        T[q * KP + j] = s * apj + c * aqj;
                    }
                    __syncwarp();
                }
            }
            if (nrot == 0) break;
        }
        __syncwarp();
        if (lane < KP) {                                     // rank of each eigenvalue, descending (ties by index)
            const double mine = T[lane * KP + lane];
            int rk = 0;
            for (int o = 0; o < KP; ++o) {
                const double ot = T[o * KP + o];
                rk += (ot > mine || (ot == mine && o < lane)) ? 1 : 0;
            }
            order[rk] = lane;
        }
        __syncwarp();
        if (lane < K) dvals[lane] = sqrt(fmax(T[order[lane] * KP + order[lane]], 0.0));
    }
    __syncthreads();
    for (int e = tid; e < ldG * QS; e += 1024) {
        const int g = e / QS, k = e - g * QS;
        double s = 0.0;
        if (g < G && k < K) {
            const int col = order[k];
            for (int c = 0; c < KP; ++c) s = fma(Q[size_t(g) * KP + c], V[c * KP + col], s);
        }
        A[e] = s;
    }
}

// ------------------------------------------------------------------------------------------ k-means
// initial centres of restart r: M0 distinct cells drawn from the Philox cell stream (sweep = r, "cell" = j, block = try)
__global__ void k_km_init(KmArgs a, unsigned long long seed, long long n_total, long long cell_offset) {
    const int r = blockIdx.x * blockDim.x + threadIdx.x;
    if (r >= a.R) return;
    long long picks[32];
    for (int j = 0; j < a.M0; ++j) {
        for (unsigned t = 0;; ++t) {
            double u0, u1;
            cell_uniform_pair(seed, (unsigned)r, (unsigned long long)j, t, u0, u1);
            long long c = (long long)(u0 * double(n_total));
            if (c >= n_total) c = n_total - 1;
            bool dup = false;
            for (int p = 0; p < j; ++p) dup |= (picks[p] == c);
            if (!dup || t >= 64u || n_total <= j) {
                picks[j] = c;
                break;
            }
        }
        const long long loc = picks[j] - cell_offset;
        double* dst = a.cent + (size_t(r) * a.M0 + j) * a.K;
        for (int k = 0; k < a.K; ++k) dst[k] = (loc >= 0 && loc < a.n) ? a.X[size_t(loc) * a.ldx + k] : 0.0;
    }
}

// One Lloyd pass for RG restarts x one cell range per warp.  Threads = cells for the distance search (centres are
// broadcast from shared memory), lanes = dimensions for the ordered accumulation of the new sums.
// mode 0: part[wid][r][j][0..K) += x, [K] += 1;  mode 1: wss[wid][r] += min distance;  mode 2: labels of restart rsel
constexpr int KMW = 4;   // warps per CTA
__global__ void __launch_bounds__(KMW * 32) k_km_pass(KmArgs a, int RG, int cells_per_warp, int mode, int rsel,
                                                      int32_t* labels) {
    extern __shared__ __align__(16) double sm[];
    const int K = a.K, M0 = a.M0, KS = K + 1;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int r0 = (mode == 2) ? rsel : blockIdx.y * RG;
    const int nr = (mode == 2) ? 1 : min(RG, a.R - r0);
    if (mode == 0) {      // a restart whose centres did not move in the last update is converged: its sums stay valid
        bool live = false;
        for (int rr = 0; rr < nr; ++rr) live |= a.moved[r0 + rr] != 0;
        if (!live) return;
    }
    double* cen = sm;                                   // [RG][M0][K]
    double* xt = cen + size_t(RG) * M0 * K + size_t(warp) * 32 * KS;          // per warp [32][K+1] cell tile
    double* acc = sm + size_t(RG) * M0 * K + size_t(KMW) * 32 * KS + size_t(warp) * RG * M0 * KS;   // per warp [RG][M0][K+1]
    for (int e = threadIdx.x; e < nr * M0 * K; e += blockDim.x) cen[e] = a.cent[size_t(r0) * M0 * K + e];
    for (int e = lane; e < RG * M0 * KS; e += 32) acc[e] = 0.0;
    __syncthreads();
    const int wid = blockIdx.x * KMW + warp;
    const int i0 = wid * cells_per_warp, i1 = min(a.n, i0 + cells_per_warp);
    double wss[8];
#pragma unroll
    for (int q = 0; q < 8; ++q) wss[q] = 0.0;
    for (int ib = i0; ib < i1; ib += 32) {
        const int nb = min(32, i1 - ib);
        __syncwarp();
        for (int e = lane; e < nb * K; e += 32) {
            const int c = e / K, k = e - c * K;
            xt[c * KS + k] = a.X[size_t(ib + c) * a.ldx + k];
        }
        __syncwarp();
#pragma unroll 1
        for (int rr = 0; rr < nr; ++rr) {
            int best = 0;
            double bd = 1.0e308;
            if (lane < nb) {
                const double* x = xt + lane * KS;
                int j = 0;
                for (; j + 1 < M0; j += 2) {                       // two centres at a time: independent FMA chains
                    const double* c0 = cen + (size_t(rr) * M0 + j) * K;
                    const double* c1 = c0 + K;
                    double d0 = 0.0, d1 = 0.0;
                    for (int k = 0; k < K; ++k) {
                        const double xk = x[k];
                        const double t0 = xk - c0[k], t1 = xk - c1[k];
                        d0 = fma(t0, t0, d0);
                        d1 = fma(t1, t1, d1);
                    }
                    if (d0 < bd) { bd = d0; best = j; }
                    if (d1 < bd) { bd = d1; best = j + 1; }
                }
                if (j < M0) {
                    const double* c0 = cen + (size_t(rr) * M0 + j) * K;
                    double d0 = 0.0;
                    for (int k = 0; k < K; ++k) {
                        const double t0 = x[k] - c0[k];
                        d0 = fma(t0, t0, d0);
                    }
                    if (d0 < bd) { bd = d0; best = j; }
                }
            }
            if (mode == 0) {
                double* ac = acc + size_t(rr) * M0 * KS;
                for (int c = 0; c < nb; ++c) {                 // cells in order => fixed summation order
                    const int lab = __shfl_sync(0xffffffffu, best, c);
                    if (lane < K) ac[lab * KS + lane] += xt[c * KS + lane];
                    else if (lane == 31) ac[lab * KS + K] += 1.0;
                    if (K == 32 && lane == 0) ac[lab * KS + K] += 1.0;
                }
            } else if (mode == 1) {
                double s = 0.0;
                for (int c = 0; c < nb; ++c) s += __shfl_sync(0xffffffffu, bd, c);
                if (rr < 8) wss[rr] += s;
            } else if (lane < nb) {
                labels[ib + lane] = best + 1;
            }
        }
    }
    __syncwarp();
    if (mode == 0) {
        for (int e = lane; e < nr * M0 * KS; e += 32) a.part[(size_t(wid) * a.R + r0) * M0 * KS + e] = acc[e];
    } else if (mode == 1 && lane == 0) {
        for (int rr = 0; rr < nr; ++rr) a.wss_part[size_t(wid) * a.R + r0 + rr] = wss[rr];
    }
}

// centres <- sums / counts (an empty cluster keeps its centre); moved[r] = any centre of restart r changed;
// *nlive counts the restarts still moving (read by the host every few iterations to stop early)
__global__ void k_km_update(KmArgs a, const double* sums, int* nlive) {
    const int r = blockIdx.x;
    __shared__ int changed;
    if (threadIdx.x == 0) changed = 0;
    __syncthreads();
    if (a.moved[r]) {
        int ch = 0;
        for (int e = threadIdx.x; e < a.M0 * a.K; e += blockDim.x) {
            const int j = e / a.K, k = e - j * a.K;
            const size_t rj = size_t(r) * a.M0 + j;
            const double cnt = sums[rj * (a.K + 1) + a.K];
            if (cnt > 0.0) {
                const double v = sums[rj * (a.K + 1) + k] / cnt;
                double* c = a.cent + rj * a.K + k;
                ch |= (v != *c);
                *c = v;
            }
        }
        if (ch) atomicOr(&changed, 1);
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        a.moved[r] = changed;
        if (changed) atomicAdd(nlive, 1);
    }
}

// cellmat (n x K col-major, host layout) from P [n][ldx]
__global__ void k_cellmat_out(const double* P, double* out, int n, int K, int ldx) {
    const size_t e = blockIdx.x * size_t(blockDim.x) + threadIdx.x;
    if (e >= size_t(n) * K) return;
    const int k = int(e / n), i = int(e - size_t(k) * n);
    out[e] = P[size_t(i) * ldx + k];
}

}  // namespace

void launch_extract_cols(const double* Y, double* V, int n, int ldG, int G, int g0, int ncol, int VS, cudaStream_t st) {
    const size_t tot = size_t(n) * VS;
    k_extract_cols<<<unsigned((tot + 255) / 256), 256, 0, st>>>(Y, V, n, ldG, G, g0, ncol, VS);
}
void launch_scatter_block(const double* blk, double* C, int rows, int ldc, int NVP, int g0, int ncol, cudaStream_t st) {
    k_scatter_block<<<(rows * ncol + 255) / 256, 256, 0, st>>>(blk, C, rows, ldc, NVP, g0, ncol);
}
void launch_sd_from_ss(const double* ss, double* sd, double* zero, int G, int ldG, double nm1, cudaStream_t st) {
    k_sd_from_ss<<<(ldG + 127) / 128, 128, 0, st>>>(ss, sd, zero, G, ldG, nm1);
}
void launch_subspace(const double* C, double* Q, double* Z, double* A, double* dvals, int G, int ldG, int KP, int K, int QS,
                     int iters, unsigned long long seed, cudaStream_t st) {
    k_sub_init<<<(ldG * KP + 255) / 256, 256, 0, st>>>(Q, G, ldG, KP, seed);
    for (int it = 0; it < iters; ++it) {
        k_sub_mult<<<(ldG + 7) / 8, 256, 0, st>>>(C, Q, Z, G, ldG, KP);
        k_sub_orth<<<1, 1024, 0, st>>>(Z, Q, G, KP);
    }
    k_sub_mult<<<(ldG + 7) / 8, 256, 0, st>>>(C, Q, Z, G, ldG, KP);
    k_sub_ritz<<<1, 1024, 0, st>>>(Q, Z, A, dvals, G, ldG, KP, K, QS);
}

int km_restart_group(int M0, int K) {
    const size_t per = size_t(M0) * K * 8 + size_t(KMW) * M0 * (K + 1) * 8;
    int rg = int((40 * 1024 - size_t(KMW) * 32 * (K + 1) * 8) / per);   // ~40 KB per CTA => 5 CTAs (20 warps) per SM
    return rg < 1 ? 1 : (rg > 8 ? 8 : rg);
}
void launch_km_init(const KmArgs& a, unsigned long long seed, long long n_total, long long cell_offset, cudaStream_t st) {
    k_km_init<<<(a.R + 63) / 64, 64, 0, st>>>(a, seed, n_total, cell_offset);
}
void launch_km_pass(const KmArgs& a, int mode, int rsel, int32_t* labels, cudaStream_t st) {
    const int RG = km_restart_group(a.M0, a.K);
    const size_t smem = sizeof(double) * (size_t(RG) * a.M0 * a.K + size_t(KMW) * 32 * (a.K + 1) +
                                          size_t(KMW) * RG * a.M0 * (a.K + 1));
    ensure_dyn_smem_k(k_km_pass, smem);
    const int nwarps = a.nsplit * KMW;
    const int cpw = ((a.n + nwarps - 1) / nwarps + 31) / 32 * 32;
    dim3 grid(a.nsplit, mode == 2 ? 1 : (a.R + RG - 1) / RG);
    k_km_pass<<<grid, KMW * 32, smem, st>>>(a, RG, cpw, mode, rsel, labels);
}
void launch_km_update(const KmArgs& a, const double* sums, int* nlive, cudaStream_t st) {
    k_km_update<<<a.R, 128, 0, st>>>(a, sums, nlive);
}
void launch_cellmat_out(const double* P, double* out, int n, int K, int ldx, cudaStream_t st) {
    const size_t tot = size_t(n) * K;
    k_cellmat_out<<<unsigned((tot + 255) / 256), 256, 0, st>>>(P, out, n, K, ldx);
}

}  // namespace sb
