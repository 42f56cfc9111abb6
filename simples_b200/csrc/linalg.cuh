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

// Symmetric square root of SPD matrix A (K x K, smem) -> R.  A is destroyed, V is K x K scratch.
__device__ inline void warp_sym_sqrt(double* A, double* V, double* R, int K, int lane) {
    for (int e = lane; e < K * K; e += 32) V[e] = ((e / K) == (e % K)) ? 1.0 : 0.0;
    __syncwarp();
    for (int sweep = 0; sweep < 30; ++sweep) {
        int nrot = 0;
        for (int p = 0; p < K - 1; ++p) {
            for (int q = p + 1; q < K; ++q) {
                const double apq = A[p * K + q];
                const double app = A[p * K + p], aqq = A[q * K + q];
                __syncwarp();
                // converged pair: the rotation would not change the diagonal in double precision
                if (fabs(apq) <= 4e-16 * sqrt(fabs(app * aqq)) || fabs(apq) <= 1e-300) continue;
                ++nrot;
                const double theta = (aqq - app) / (2.0 * apq);
                const double tt = (theta >= 0.0 ? 1.0 : -1.0) / (fabs(theta) + sqrt(theta * theta + 1.0));
                const double c = 1.0 / sqrt(tt * tt + 1.0), s = tt * c;
                if (lane < K) {   // columns p, q
                    const int i = lane;
                    const double aip = A[i * K + p], aiq = A[i * K + q];
                    A[i * K + p] = c * aip - s * aiq;
                    A[i * K + q] = s * aip + c * aiq;
                    const double vip = V[i * K + p], viq = V[i * K + q];
                    V[i * K + p] = c * vip - s * viq;
                    V[i * K + q] = s * vip + c * viq;
                }
                __syncwarp();
                if (lane < K) {   // rows p, q
                    const int j = lane;
                    const double apj = A[p * K + j], aqj = A[q * K + j];
                    A[p * K + j] = c * apj - s * aqj;
                    A[q * K + j] = s * apj + c * aqj;
                }
                __syncwarp();
            }
        }
        if (nrot == 0) break;
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
