// Prologue / epilogue / bookkeeping kernels: mask build (Y0 <= cutoff, R/EM_impute.R:173), clip and
// gene means (R/EM_impute.R:82-90), centring, layout transposes, do_impute accumulators
// (R/do_impute.R:139-148,156-159) and the optional n x n consensus matrix (R/do_impute.R:96).
#include "ptx.cuh"
#include "rng.cuh"
#include "state.h"

namespace sb {
namespace {

__global__ void k_build_mask(const double* Y0, uint32_t* mask, int ncell, int cell0, int n_alloc, int G, int ldG,
                             double cutoff, unsigned long long* nnz) {
    const int nwords = ldG / 32;
    const int lane = threadIdx.x & 31;
    const long long nw = (long long)ncell * nwords;
    unsigned long long cnt = 0;
    for (long long w = (blockIdx.x * (long long)blockDim.x + threadIdx.x) >> 5; w < nw;
         w += ((long long)gridDim.x * blockDim.x) >> 5) {
        const int cell = int(w / nwords), word = int(w - (long long)cell * nwords);
        const int g = word * 32 + lane;
        const bool bit = (g < G) && (Y0[size_t(cell) * G + g] <= cutoff);
        const unsigned b = __ballot_sync(0xffffffffu, bit);
        if (lane == 0) {
            mask[size_t(word) * n_alloc + cell0 + cell] = b;
            cnt += __popc(b);
        }
    }
    if (lane == 0 && cnt) atomicAdd(nnz, cnt);
}

__global__ void k_clip(double* Y, int n, int G, int ldG, double lower, double upper) {
    const size_t total = size_t(n) * ldG;
    for (size_t i = blockIdx.x * size_t(blockDim.x) + threadIdx.x; i < total; i += size_t(gridDim.x) * blockDim.x) {
        if (int(i % ldG) < G) {
            double v = Y[i];
            v = v > upper ? upper : v;
            v = v < lower ? lower : v;
            Y[i] = v;
        }
    }
}

__global__ void k_rowsum(const double* Y, double* part, int n, int ldG, int cps) {
    const int g = blockIdx.x * blockDim.x + threadIdx.x;
    const int sp = blockIdx.y;
    if (g >= ldG) return;
    const int cbeg = sp * cps, cend = min(n, cbeg + cps);
    double s = 0.0;
    for (int i = cbeg; i < cend; ++i) s += Y[size_t(i) * ldG + g];
    part[size_t(sp) * ldG + g] = s;
}

__global__ void k_center(double* Y, const double* gmean, const double* gsd, int n, int ldG) {
    const size_t total = size_t(n) * ldG;
    for (size_t i = blockIdx.x * size_t(blockDim.x) + threadIdx.x; i < total; i += size_t(gridDim.x) * blockDim.x) {
        const int g = int(i % ldG);
        Y[i] = (Y[i] - gmean[g]) / gsd[g];
    }
}

__global__ void k_transpose(const double* in, double* out, int rows, int cols, int ld_in, int ld_out) {
    __shared__ double tile[32][33];
    const int bx = blockIdx.x * 32, by = blockIdx.y * 32;
    for (int r = threadIdx.y; r < 32; r += blockDim.y) {
        const int i = by + r, j = bx + threadIdx.x;
        if (i < rows && j < cols) tile[r][threadIdx.x] = in[size_t(i) * ld_in + j];
    }
    __syncthreads();
    for (int r = threadIdx.y; r < 32; r += blockDim.y) {
        const int j = bx + r, i = by + threadIdx.x;
        if (i < rows && j < cols) out[size_t(j) * ld_out + i] = tile[threadIdx.x][r];
    }
}

__global__ void k_fill(double* p, size_t count, double v) {
    for (size_t i = blockIdx.x * size_t(blockDim.x) + threadIdx.x; i < count; i += size_t(gridDim.x) * blockDim.x) p[i] = v;
}
__global__ void k_scale(double* p, size_t count, double s) {
    for (size_t i = blockIdx.x * size_t(blockDim.x) + threadIdx.x; i < count; i += size_t(gridDim.x) * blockDim.x) p[i] *= s;
}

// A[r][c] /= d[c]   (row-major [rows][cols]); columns with d = 0 are left untouched
__global__ void k_div_cols(double* A, const double* d, size_t rows, int cols) {
    const size_t total = rows * cols;
    for (size_t i = blockIdx.x * size_t(blockDim.x) + threadIdx.x; i < total; i += size_t(gridDim.x) * blockDim.x) {
        const double dv = d[i % cols];
        if (dv != 0.0) A[i] /= dv;
    }
}

__global__ void k_uncenter(const double* Y, double* out, const double* gmean, const double* gsd, int n, int G, int ldG) {
    const size_t total = size_t(n) * G;
    for (size_t i = blockIdx.x * size_t(blockDim.x) + threadIdx.x; i < total; i += size_t(gridDim.x) * blockDim.x) {
        const size_t c = i / G;
        const int g = int(i - c * G);
        out[i] = Y[c * ldG + g] * gsd[g] + gmean[g];
    }
}

// Slot table of the compact MI accumulators.  One thread per 8-cell group: off[g][w] = zero entries of the group in the
// chunks before w, tot[g] = zero entries of the group.
__global__ void k_mi_group_counts(const uint32_t* mask, int n_alloc, int nwords, int ngroups, uint32_t* off, unsigned long long* tot) {
    const int g = blockIdx.x * blockDim.x + threadIdx.x;
    if (g >= ngroups) return;
    uint32_t run = 0;
    for (int w = 0; w < nwords; ++w) {
        off[size_t(g) * nwords + w] = run;
        const uint4* p = reinterpret_cast<const uint4*>(mask + size_t(w) * n_alloc + size_t(g) * 8);
        const uint4 a = p[0], b = p[1];
        run += __popc(a.x) + __popc(a.y) + __popc(a.z) + __popc(a.w) + __popc(b.x) + __popc(b.y) + __popc(b.z) + __popc(b.w);
    }
    tot[g] = run;
}
// exclusive scan of tot[0..ngroups) in place (one CTA: contiguous segment per thread, then a scan of the segment sums)
__global__ void __launch_bounds__(1024) k_mi_scan(unsigned long long* tot, int ngroups) {
    __shared__ unsigned long long seg[1024];
    const int per = (ngroups + 1023) / 1024, b = threadIdx.x * per, e = min(ngroups, b + per);
    unsigned long long s = 0;
    for (int i = b; i < e; ++i) s += tot[i];
    seg[threadIdx.x] = s;
    __syncthreads();
    if (threadIdx.x == 0) {
        unsigned long long run = 0;
        for (int i = 0; i < 1024; ++i) {
            const unsigned long long v = seg[i];
            seg[i] = run;
            run += v;
        }
    }
    __syncthreads();
    unsigned long long run = seg[threadIdx.x];
    for (int i = b; i < e; ++i) {
        const unsigned long long v = tot[i];
        tot[i] = run;
        run += v;
    }
}

__global__ void k_mi_finalize(const double* Y, const uint32_t* mask, const double* rec1, const double* rec2,
                              const double* gmean, const double* gsd, double* impt, double* impt_var, int n, int cell0,
                              int n_alloc, int G, int ldG, int mcmc, const unsigned long long* gbase, const uint32_t* off) {
    const size_t total = size_t(n) * G;
    const int nwords = ldG / 32;
    for (size_t i = blockIdx.x * size_t(blockDim.x) + threadIdx.x; i < total; i += size_t(gridDim.x) * blockDim.x) {
        const size_t c = i / G;
        const int g = int(i - c * G);
        const size_t o = c * ldG + g;
        const size_t cell = size_t(cell0) + c;
        const int w = g >> 5, bit = g & 31;
        const uint32_t* mw = mask + size_t(w) * n_alloc;
        const uint32_t word = mw[cell];
        double m1, v;
        if ((word >> bit) & 1u) {
            // slot = group base + block offset + entries of the block's earlier rows + earlier genes of this row
            const size_t grp = cell >> 3;
            size_t slot = gbase[grp] + off[grp * nwords + w] + __popc(word & ((1u << bit) - 1u));
            for (size_t r = grp << 3; r < cell; ++r) slot += __popc(mw[r]);
            const double2 q = reinterpret_cast<const double2*>(rec1)[slot];      // {sum v, sum v^2} (rec2 unused)
            m1 = q.x / mcmc;
            v = q.y / mcmc - m1 * m1;
        } else {
            m1 = Y[o];     // constant over sweeps: mean = Y, variance = 0
            v = 0.0;
        }
        if (impt) impt[i] = m1 * gsd[g] + gmean[g];
        if (impt_var) impt_var[i] = v;
    }
}

__global__ void k_mi_record_factors(Dims d, const double* W, const double* z, const double* Mm, double* rEF, double* rF2,
                                    double* rVF) {
    const int K0 = d.K0, M0 = d.M0, n = d.n;
    const size_t total = size_t(n) * K0;
    for (size_t e = blockIdx.x * size_t(blockDim.x) + threadIdx.x; e < total; e += size_t(gridDim.x) * blockDim.x) {
        const size_t i = e / K0;
        const int k = int(e - i * K0);
        double ef = 0.0, f2 = 0.0, vf = 0.0;
        for (int m = 0; m < M0; ++m) {
            const double zz = z[i * M0 + m];
            const double w = W[(size_t(m) * n + i) * K0 + k];
            ef = fma(zz, w, ef);
            f2 = fma(zz * w, w, f2);
            vf = fma(zz, Mm[(size_t(m) * K0 + k) * K0 + k], vf);
        }
        rEF[e] += ef;
        rF2[e] += f2;
        rVF[e] += vf;
    }
}

__global__ void k_consensus(const double* z, double* cons, int n, int M0) {
    const int i = blockIdx.y * 16 + threadIdx.y, j = blockIdx.x * 16 + threadIdx.x;
    if (i >= n || j >= n) return;
    double s = 0.0;
    for (int m = 0; m < M0; ++m) s = fma(z[size_t(i) * M0 + m], z[size_t(j) * M0 + m], s);
    cons[size_t(j) * n + i] += s;   // symmetric; column-major n x n
}

// Sharded consensus (R/do_impute.R:96): this shard's row block of z z', rows = local cells, columns = ALL cells.
// cons is column-major [n_total][n_local]: cons[j * nl + i] += <z_loc[i], z_all[j]>.
__global__ void k_consensus_rows(const double* zloc, const double* zall, double* cons, int nl, int ntot, int M0) {
    const int i = blockIdx.y * 16 + threadIdx.y, j = blockIdx.x * 16 + threadIdx.x;
    if (i >= nl || j >= ntot) return;
    double s = 0.0;
    for (int m = 0; m < M0; ++m) s = fma(zloc[size_t(i) * M0 + m], zall[size_t(j) * M0 + m], s);
    cons[size_t(j) * nl + i] += s;
}

// Yimp0 chunk (R/SIMPLE.R:421-426): out[i][g] = (sum_m mu[g,m] z[i,m] + sum_k beta[g,k] Wbar[i,k]) gsd[g] + gmean[g],
// Wbar_i = sum_m z_im Ef_m[i,] (first K0 columns of the M-step operand row V_i).
__global__ void k_yimp0(Dims d, const double* mu, const double* beta, const double* z, const double* V, const double* gmean,
                        const double* gsd, double* out, int cell0, int ncell) {
    extern __shared__ double sh[];          // [8][M0 + K0]
    const int M0 = d.M0, K0 = d.K0, W = M0 + K0;
    const int c0 = blockIdx.y * 8;
    for (int idx = threadIdx.x; idx < 8 * W; idx += blockDim.x) {
        const int r = idx / W, c = idx - r * W;
        const int cell = cell0 + c0 + r;
        double v = 0.0;
        if (c0 + r < ncell) v = (c < M0) ? z[size_t(cell) * M0 + c] : V[size_t(cell) * d.VS + (c - M0)];
        sh[idx] = v;
    }
    __syncthreads();
    const int g = blockIdx.x * blockDim.x + threadIdx.x;
    if (g >= d.G) return;
    double acc[8] = {0, 0, 0, 0, 0, 0, 0, 0};
    for (int m = 0; m < M0; ++m) {
        const double a = mu[size_t(g) * M0 + m];
#pragma unroll
        for (int r = 0; r < 8; ++r) acc[r] = fma(a, sh[r * W + m], acc[r]);
    }
    for (int k = 0; k < K0; ++k) {
        const double b = beta[size_t(g) * K0 + k];
#pragma unroll
        for (int r = 0; r < 8; ++r) acc[r] = fma(b, sh[r * W + M0 + k], acc[r]);
    }
    const double sd = gsd[g], mean = gmean[g];
#pragma unroll
    for (int r = 0; r < 8; ++r)
        if (c0 + r < ncell) out[size_t(c0 + r) * d.G + g] = acc[r] * sd + mean;
}

inline int gs_grid(size_t total) {
    size_t g = (total + 255) / 256;
    if (g > 148 * 16) g = 148 * 16;
    if (g < 1) g = 1;
    return int(g);
}

// lq-gene fit: regression operand Vg = [z_m Ef_m | z | sqrt z] and the sweep operand Fh = Ef_{m_i}[i,]
// with m_i drawn from z[i,] (R/SIMPLE.R:376-382; the factor is NOT sampled here, :386).
__global__ void k_lq_cell(LqCellArgs a) {
    const Dims& d = a.d;
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= d.n) return;
    const int K0 = d.K0, M0 = d.M0, ncg = K0 * M0 + 2 * M0;
    double* vg = a.Vg + size_t(i) * ncg;
    const double* zi = a.z + size_t(i) * M0;
    int im = 0;
    if (a.draw_membership) {
        double u0, u1;
        cell_uniform_pair(a.seed, a.sweep, (unsigned long long)(d.cell_offset + i), 0u, u0, u1);
        double cum = 0.0;
        for (int m = 0; m < M0; ++m) {
            cum += zi[m];
            im += (u0 >= cum) ? 1 : 0;
        }
        im = min(im, M0 - 1);
    }
    a.im[i] = im;
    double* fh = a.Fh + size_t(i) * d.FS;
    for (int m = 0; m < M0; ++m) {
        const double zm = zi[m];
        const double* w = a.W + (size_t(m) * d.n + i) * K0;
        for (int k = 0; k < K0; ++k) {
            vg[m * K0 + k] = zm * w[k];
            if (m == im) fh[k] = w[k];
        }
        vg[K0 * M0 + m] = zm;
        vg[K0 * M0 + M0 + m] = sqrt(zm);
    }
    for (int c = K0; c < d.FS; ++c) fh[c] = 0.0;
}

// number of non-finite entries of a buffer (input validation)
__global__ void k_count_nonfinite(const double* p, size_t count, unsigned long long* out) {
    unsigned long long c = 0;
    for (size_t i = blockIdx.x * size_t(blockDim.x) + threadIdx.x; i < count; i += size_t(gridDim.x) * blockDim.x)
        c += isfinite(p[i]) ? 0ull : 1ull;
    c = __reduce_add_sync(0xffffffffu, (unsigned)c);
    if ((threadIdx.x & 31) == 0 && c) atomicAdd(out, c);
}

// out[m] = sum_i z[i][m] in a fixed order (one CTA), out[M0 .. 2 M0] = 0
__global__ void __launch_bounds__(256) k_colsum(const double* z, double* out, int n, int M0) {
    __shared__ double red[256];
    for (int m = 0; m < M0; ++m) {
        double s = 0.0;
        for (int i = threadIdx.x; i < n; i += 256) s += z[size_t(i) * M0 + m];
        red[threadIdx.x] = s;
        __syncthreads();
        for (int o = 128; o > 0; o >>= 1) {
            if (threadIdx.x < o) red[threadIdx.x] += red[threadIdx.x + o];
            __syncthreads();
        }
        if (threadIdx.x == 0) out[m] = red[0];
        __syncthreads();
    }
    if (threadIdx.x <= M0) out[M0 + threadIdx.x] = 0.0;
}

}  // namespace

void launch_build_mask(const double* Y0chunk, uint32_t* mask, int ncell, int cell0, int n_alloc, int G, int ldG,
                       double cutoff, unsigned long long* nnz, cudaStream_t st) {
    k_build_mask<<<gs_grid(size_t(ncell) * ldG), 256, 0, st>>>(Y0chunk, mask, ncell, cell0, n_alloc, G, ldG, cutoff, nnz);
}
void launch_clip(double* Y, int n, int G, int ldG, double lower, double upper, cudaStream_t st) {
    k_clip<<<gs_grid(size_t(n) * ldG), 256, 0, st>>>(Y, n, G, ldG, lower, upper);
}
void launch_rowsum(const double* Y, double* part, int n, int ldG, int nsplit, cudaStream_t st) {
    const int cps = (n + nsplit - 1) / nsplit;
    dim3 grid((ldG + 127) / 128, nsplit);
    k_rowsum<<<grid, 128, 0, st>>>(Y, part, n, ldG, cps);
}
void launch_center(double* Y, const double* gmean, const double* gsd, int n, int ldG, double, int, cudaStream_t st) {
    k_center<<<gs_grid(size_t(n) * ldG), 256, 0, st>>>(Y, gmean, gsd, n, ldG);
}
void launch_transpose(const double* in, double* out, int rows, int cols, int ld_in, int ld_out, cudaStream_t st) {
    dim3 grid((cols + 31) / 32, (rows + 31) / 32), block(32, 8);
    k_transpose<<<grid, block, 0, st>>>(in, out, rows, cols, ld_in, ld_out);
}
void launch_fill(double* p, size_t count, double v, cudaStream_t st) { k_fill<<<gs_grid(count), 256, 0, st>>>(p, count, v); }
void launch_div_cols(double* A, const double* d, size_t rows, int cols, cudaStream_t st) {
    k_div_cols<<<gs_grid(rows * cols), 256, 0, st>>>(A, d, rows, cols);
}
void launch_scale(double* p, size_t count, double s, cudaStream_t st) { k_scale<<<gs_grid(count), 256, 0, st>>>(p, count, s); }
void launch_uncenter(const double* Y, double* out, const double* gmean, const double* gsd, int n, int G, int ldG,
                     cudaStream_t st) {
    k_uncenter<<<gs_grid(size_t(n) * G), 256, 0, st>>>(Y, out, gmean, gsd, n, G, ldG);
}
void launch_mi_finalize(const double* Y, const uint32_t* mask, const double* rec1, const double* rec2, const double* gmean,
                        const double* gsd, double* impt, double* impt_var, int n, int cell0, int n_alloc, int G, int ldG,
                        int mcmc, const unsigned long long* gbase, const uint32_t* off, cudaStream_t st) {
    k_mi_finalize<<<gs_grid(size_t(n) * G), 256, 0, st>>>(Y, mask, rec1, rec2, gmean, gsd, impt, impt_var, n, cell0, n_alloc, G,
                                                          ldG, mcmc, gbase, off);
}
void launch_mi_index(const uint32_t* mask, int n, int n_alloc, int nwords, unsigned long long* gbase, uint32_t* off,
                     cudaStream_t st) {
    const int ngroups = (n + 7) / 8;
    k_mi_group_counts<<<(ngroups + 127) / 128, 128, 0, st>>>(mask, n_alloc, nwords, ngroups, off, gbase);
    k_mi_scan<<<1, 1024, 0, st>>>(gbase, ngroups);
}
void launch_mi_record_factors(const Dims& d, const double* W, const double* z, const double* Mm, double* rEF, double* rF2,
                              double* rVF, cudaStream_t st) {
    k_mi_record_factors<<<gs_grid(size_t(d.n) * d.K0), 256, 0, st>>>(d, W, z, Mm, rEF, rF2, rVF);
}
void launch_yimp0(const Dims& d, const double* mu, const double* beta, const double* z, const double* V, const double* gmean,
                  const double* gsd, double* out, int cell0, int ncell, cudaStream_t st) {
    dim3 grid((d.G + 127) / 128, (ncell + 7) / 8);
    k_yimp0<<<grid, 128, sizeof(double) * 8 * (d.M0 + d.K0), st>>>(d, mu, beta, z, V, gmean, gsd, out, cell0, ncell);
}
void launch_consensus(const double* z, double* cons, int n, int M0, cudaStream_t st) {
    dim3 grid((n + 15) / 16, (n + 15) / 16), block(16, 16);
    k_consensus<<<grid, block, 0, st>>>(z, cons, n, M0);
}

void launch_consensus_rows(const double* zloc, const double* zall, double* cons, int nl, int ntot, int M0, cudaStream_t st) {
    dim3 grid((ntot + 15) / 16, (nl + 15) / 16), block(16, 16);
    k_consensus_rows<<<grid, block, 0, st>>>(zloc, zall, cons, nl, ntot, M0);
}

void launch_count_nonfinite(const double* p, size_t count, unsigned long long* out, cudaStream_t st) {
    k_count_nonfinite<<<gs_grid(count), 256, 0, st>>>(p, count, out);
}
void launch_lq_cell(const LqCellArgs& a, cudaStream_t st) { k_lq_cell<<<(a.d.n + 127) / 128, 128, 0, st>>>(a); }
void launch_colsum(const double* z, double* out, int n, int M0, cudaStream_t st) { k_colsum<<<1, 256, 0, st>>>(z, out, n, M0); }

}  // namespace sb
