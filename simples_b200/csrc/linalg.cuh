// Small dense linear algebra in shared memory, executed by ONE WARP (K <= 32):
// Cholesky, SPD inverse, log-determinant, cyclic Jacobi eigen-decomposition / symmetric sqrt.
// Replaces base-R solve()/det()/chol() (R/EM_impute.R:124-125,311) and the per-cell eigen() inside
// mixtools::rmvnorm (R/EM_impute.R:171) -- the square root is hoisted to once per cluster.
#pragma once
#include <cuda_runtime.h>

namespace sb {

// In-place lower Cholesky of the K x K matrix A (row-major, stride K).  Upper part left untouched.
__device__ inline void warp_cholesky(double* A, int K, int lane) {
    for (int j = 0; j < K; ++j) {
        __syncwarp();
        double d = A[j * K + j];
        for (int p = 0; p < j; ++p) d -= A[j * K + p] * A[j * K + p];
        d = sqrt(d);
        __syncwarp();
        if (lane == 0) A[j * K + j] = d;
        const int i = j + 1 + lane;
        if (i < K) {
            double s = A[i * K + j];
            for (int p = 0; p < j; ++p) s -= A[i * K + p] * A[j * K + p];
            A[i * K + j] = s / d;
        }
    }
    __syncwarp();
}

// Given lower Cholesky L (in A), write Linv (lower) into Li.  Lane c computes column c.
__device__ inline void warp_tri_inverse(const double* L, double* Li, int K, int lane) {
    if (lane < K) {
        const int c = lane;
        for (int i = 0; i < K; ++i) {
            if (i < c) { Li[i * K + c] = 0.0; continue; }
            double s = (i == c) ? 1.0 : 0.0;
            for (int p = c; p < i; ++p) s -= L[i * K + p] * Li[p * K + c];
            Li[i * K + c] = s / L[i * K + i];
        }
    }
    __syncwarp();
}

// M = Li^T Li (symmetric), Li lower triangular.
__device__ inline void warp_ltl(const double* Li, double* M, int K, int lane) {
    for (int e = lane; e < K * K; e += 32) {
        const int i = e / K, j = e - i * K;
        const int p0 = i > j ? i : j;
        double s = 0.0;
        for (int p = p0; p < K; ++p) s += Li[p * K + i] * Li[p * K + j];
        M[e] = s;
    }
    __syncwarp();
}

// Solve (L L^T) x = b in place (b in smem), by lane 0 (K small).
__device__ inline void warp_chol_solve(const double* L, double* b, int K, int lane) {
    if (lane == 0) {
        for (int i = 0; i < K; ++i) {
            double s = b[i];
            for (int p = 0; p < i; ++p) s -= L[i * K + p] * b[p];
            b[i] = s / L[i * K + i];
        }
        for (int i = K - 1; i >= 0; --i) {
            double s = b[i];
            for (int p = i + 1; p < K; ++p) s -= L[p * K + i] * b[p];
            b[i] = s / L[i * K + i];
        }
    }
    __syncwarp();
}

// Symmetric square root of SPD matrix A (K x K, smem) -> R.  A is destroyed, V is K x K scratch, W is scratch for
// 2 * 16 doubles (rotation cosines / sines of one round).
// Jacobi eigen-decomposition with the round-robin ("chess tournament") PARALLEL ordering: a sweep is K'-1 rounds
// (K' = K rounded up to even) of K'/2 disjoint pivot pairs; the rotations of a round commute in exact arithmetic up to
// the usual Jacobi coupling, so their parameters are computed together by K'/2 lanes and the column / row updates of
// all pairs are spread over the warp.  A round costs three warp barriers instead of three per rotation (the cyclic
// ordering made cluster_prep the longest replicated kernel of a strongly scaled run: profiles/r02_notes.md).
// Vwarm (optional, K x K orthogonal, global or shared): eigenvectors of a NEARBY matrix (the previous EM iteration's);
// A is first rotated into that basis, where it is almost diagonal, and the Jacobi sweeps continue from there (2 sweeps
// instead of ~7).  On return V holds the eigenvectors of the input A either way.
__device__ inline void warp_sym_sqrt(double* A, double* V, double* R, double* W, int K, int lane,
                                     const double* Vwarm = nullptr) {
    if (Vwarm) {
        for (int e = lane; e < K * K; e += 32) V[e] = Vwarm[e];
        __syncwarp();
        for (int e = lane; e < K * K; e += 32) {          // R = A V
            const int i = e / K, j = e - i * K;
            double s = 0.0;
            for (int p = 0; p < K; ++p) s = fma(A[i * K + p], V[p * K + j], s);
            R[e] = s;
        }
        __syncwarp();
        for (int e = lane; e < K * K; e += 32) {          // A = V' R, symmetrised
            const int i = e / K, j = e - i * K;
            if (i > j) continue;
            double s = 0.0;
            for (int p = 0; p < K; ++p) s = fma(V[p * K + i], R[p * K + j], s);
            A[i * K + j] = s;
            A[j * K + i] = s;
        }
    } else {
        for (int e = lane; e < K * K; e += 32) V[e] = ((e / K) == (e % K)) ? 1.0 : 0.0;
    }
    __syncwarp();
    const int Ke = (K + 1) & ~1, H = Ke >> 1;          // H <= 16 pairs per round
    double* Wc = W;
    double* Ws = W + 16;
    for (int sweep = 0; sweep < 30; ++sweep) {
        unsigned any = 0u;
        for (int r = 0; r < Ke - 1; ++r) {
            // pair of slot i in round r
            auto pair_of = [&](int i, int& p, int& q) {
                int a, b;
                if (i == 0) {
                    a = Ke - 1;
                    b = r;
                } else {                       // (r + i) mod (Ke - 1), (r - i) mod (Ke - 1) without divisions
                    a = r + i;
                    if (a >= Ke - 1) a -= Ke - 1;
                    b = r - i;
                    if (b < 0) b += Ke - 1;
                }
                p = a < b ? a : b;
                q = a < b ? b : a;
            };
            bool rot = false;
            if (lane < H) {
                int p, q;
                pair_of(lane, p, q);
                double c = 1.0, sn = 0.0;
                if (q < K) {
                    const double apq = A[p * K + q], app = A[p * K + p], aqq = A[q * K + q];
                    // converged pair: the rotation would not change the diagonal in double precision
                    if (!(fabs(apq) <= 4e-16 * sqrt(fabs(app * aqq)) || fabs(apq) <= 1e-300)) {
                        const double theta = (aqq - app) / (2.0 * apq);
                        const double tt = (theta >= 0.0 ? 1.0 : -1.0) / (fabs(theta) + sqrt(theta * theta + 1.0));
                        c = 1.0 / sqrt(tt * tt + 1.0);
                        sn = tt * c;
                        rot = true;
                    }
                }
                Wc[lane] = c;
                Ws[lane] = sn;
            }
            const unsigned rb = __ballot_sync(0xffffffffu, rot);
            any |= rb;
            if (rb == 0u) continue;            // warp-uniform
            __syncwarp();
            for (int e = lane; e < 32 * H; e += 32) {     // columns: A <- A J, V <- V J  (lane = row i, one pair per step)
                const int i = lane, sl = e >> 5;
                if (i >= K) continue;
                int p, q;
                pair_of(sl, p, q);
                const double c = Wc[sl], sn = Ws[sl];
                if (q < K && sn != 0.0) {
                    const double aip = A[i * K + p], aiq = A[i * K + q];
                    A[i * K + p] = c * aip - sn * aiq;
                    A[i * K + q] = sn * aip + c * aiq;
                    const double vip = V[i * K + p], viq = V[i * K + q];
                    V[i * K + p] = c * vip - sn * viq;
                    V[i * K + q] = sn * vip + c * viq;
                }
            }
            __syncwarp();
            for (int e = lane; e < 32 * H; e += 32) {     // rows: A <- J' A  (lane = column j)
                const int j = lane, sl = e >> 5;
                if (j >= K) continue;
                int p, q;
                pair_of(sl, p, q);
                const double c = Wc[sl], sn = Ws[sl];
                if (q < K && sn != 0.0) {
                    const double apj = A[p * K + j], aqj = A[q * K + j];
                    A[p * K + j] = c * apj - sn * aqj;
                    A[q * K + j] = sn * apj + c * aqj;
                }
            }
            __syncwarp();
        }
        if (any == 0u) break;
    }
    __syncwarp();
    for (int e = lane; e < K * K; e += 32) {
        const int i = e / K, j = e - i * K;
        double s = 0.0;
        for (int p = 0; p < K; ++p) s += V[i * K + p] * sqrt(fmax(A[p * K + p], 0.0)) * V[j * K + p];
        R[e] = s;
    }
    __syncwarp();
}

}  // namespace sb
