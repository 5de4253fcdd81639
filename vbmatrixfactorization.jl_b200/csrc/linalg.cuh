// Small dense linear algebra on shared memory, cooperative over a thread group (a warp or a whole CTA):
//   * SPD inverse by diagonal equilibration + Cholesky (replaces Julia `inv` = LU getrf/getri on the H x H posterior
//     precision matrices: src/vbmf.jl:96,110,133,145; src/vbmf_sparse.jl:188,259,265; src/vbmf_dual.jl:228,297,303).
//     ARD pruning drives prior precisions to ~1e10 next to O(1) entries (seen in the reference's own log), so the
//     matrix is symmetrically scaled to unit diagonal before factorising.
//   * largest eigenvalue of a symmetric PSD matrix by parallel-order cyclic Jacobi (for Julia 0.5 `norm(::Matrix)`
//     = spectral norm in `delta`, src/util.jl:27-29).
#pragma once
#include "common.cuh"

namespace vb {

struct BlockGroup {
    __device__ __forceinline__ int tid() const { return threadIdx.x; }
    __device__ __forceinline__ int size() const { return blockDim.x; }
    __device__ __forceinline__ void sync() const { __syncthreads(); }
};
struct WarpGroup {
    __device__ __forceinline__ int tid() const { return threadIdx.x & 31; }
    __device__ __forceinline__ int size() const { return 32; }
    __device__ __forceinline__ void sync() const { __syncwarp(); }
};

// In-place inverse of the SPD matrix A (H x H, row stride ld, shared memory).  `vec` = 2*H doubles of scratch.
// Returns false (for every thread of the group) when a pivot is not positive / not finite; A is then garbage.
template <class G>
__device__ bool spd_inverse(const G& g, double* A, int ld, int H, double* vec) {
    const int t = g.tid(), n = g.size();
    double* scal = vec;
    double* col = vec + H;
    bool ok = true;
    // 1. equilibrate to unit diagonal
    for (int i = t; i < H; i += n) scal[i] = 1.0 / sqrt(A[i * ld + i]);
    g.sync();
    for (int e = t; e < H * H; e += n) {
        const int i = e / H, j = e - i * H;
        A[i * ld + j] *= scal[i] * scal[j];
    }
    g.sync();
    // 2. Cholesky, lower triangle, right-looking
    for (int k = 0; k < H; ++k) {
        const double pkk = A[k * ld + k];
        if (!(pkk > 0.0) || !(pkk < 1e300)) ok = false;
        const double rk = sqrt(pkk), ik = 1.0 / rk;
        g.sync();
        for (int i = k + t; i < H; i += n) A[i * ld + k] = (i == k) ? rk : A[i * ld + k] * ik;
        g.sync();
        const int cnt = H - k - 1;
        for (int e = t; e < cnt * cnt; e += n) {
            const int i = k + 1 + e / cnt, j = k + 1 + e % cnt;
            if (i >= j) A[i * ld + j] -= A[i * ld + k] * A[j * ld + k];
        }
        g.sync();
    }
    // 3. X = L^{-1} in place (lower), columns from last to first
    for (int j = H - 1; j >= 0; --j) {
        for (int k = j + t; k < H; k += n) col[k] = A[k * ld + j];
        g.sync();
        const double xjj = 1.0 / col[j];
        for (int i = j + t; i < H; i += n) {
            if (i == j) { A[j * ld + j] = xjj; continue; }
            double s = 0.0;
            for (int k = j + 1; k <= i; ++k) s = fma(A[i * ld + k], col[k], s);   // X[i][k] (already inverted) * L[k][j]
            A[i * ld + j] = -s * xjj;
        }
        g.sync();
    }
    // 4. inv = X' X : strict upper part into the upper triangle, diagonal into col[]
    for (int e = t; e < H * H; e += n) {
        const int i = e / H, j = e - i * H;
        if (i <= j) {
            double s = 0.0;
            for (int k = j; k < H; ++k) s = fma(A[k * ld + i], A[k * ld + j], s);
            if (i == j) col[i] = s; else A[i * ld + j] = s;
        }
    }
    g.sync();
    // 5. mirror + undo the scaling
    for (int e = t; e < H * H; e += n) {
        const int i = e / H, j = e - i * H;
        if (i == j) A[i * ld + i] = col[i] * scal[i] * scal[i];
        else if (i > j) A[i * ld + j] = A[j * ld + i] * scal[i] * scal[j];
    }
    g.sync();
    for (int e = t; e < H * H; e += n) {
        const int i = e / H, j = e - i * H;
        if (i < j) A[i * ld + j] = A[j * ld + i];
    }
    g.sync();
    return ok;
}

// Largest eigenvalue of the symmetric matrix A (n x n, full storage, row stride ld, shared memory; destroyed) by
// Householder tridiagonalisation followed by parallel multisection on the Sturm sequence: orthogonal similarity +
// bisection, so lambda_max carries an absolute error of a few ulp of ||A|| (what LAPACK's SVD gives the reference for
// norm(::Matrix)).  work = 4*n doubles.  Whole CTA (blockDim.x >= 64, multiple of 32); result valid in every thread.
__device__ inline double sym_lambda_max(double* A, int ld, int n, double* work) {
    const int t = threadIdx.x, nt = blockDim.x, lane = t & 31, warp = t >> 5;
    double* dv = work;            // diagonal of T
    double* ev = work + n;        // sub-diagonal of T (ev[k] couples k and k+1)
    double* v = work + 2 * n;     // Householder vector
    double* p = work + 3 * n;     // p / q vector
    __shared__ double s_sc[4];    // beta, alpha, K, flag
    __shared__ double s_lo, s_hi;
    __shared__ int s_first;
    if (n <= 0) return 0.0;
    // NaN / Inf anywhere -> NaN (the reference's norm would propagate it)
    int bad = 0;
    for (int e = t; e < n * n; e += nt) { const double x = A[(e / n) * ld + (e % n)]; if (!(fabs(x) <= 1.79e308)) bad = 1; }
    if (__syncthreads_or(bad)) return nan("");
    if (n == 1) return A[0];
    for (int k = 0; k < n - 2; ++k) {
        const int m = n - k - 1;                  // trailing block is rows/cols k+1 .. n-1
        double* Asub = A + (k + 1) * ld + (k + 1);
        if (warp == 0) {                          // x = A[k+1:, k] ; v = x - alpha*e1 ; beta = 2/(v'v)
            double s = 0.0;
            for (int i = lane; i < m; i += 32) { const double x = A[(k + 1 + i) * ld + k]; v[i] = x; s = fma(x, x, s); }
            s = warp_sum(s);
            __syncwarp();
            const double x0 = v[0];
            const double tail = s - x0 * x0;      // sum of squares below the first entry
            if (lane == 0) {
                if (!(tail > 0.0)) { s_sc[0] = 0.0; s_sc[1] = x0; }          // already tridiagonal in this column
                else {
                    const double alpha = (x0 >= 0.0) ? -sqrt(s) : sqrt(s);
                    const double v0 = x0 - alpha;
                    v[0] = v0;
                    s_sc[0] = 2.0 / (tail + v0 * v0);
                    s_sc[1] = alpha;
                }
            }
        }
        __syncthreads();
        const double beta = s_sc[0];
        if (beta != 0.0) {
            if (t < m) {                          // p = beta * Asub * v   (column access: A is symmetric)
                double s = 0.0;
                for (int j = 0; j < m; ++j) s = fma(Asub[j * ld + t], v[j], s);
                p[t] = beta * s;
            }
            __syncthreads();
            if (warp == 0) {                      // K = beta/2 * p'v ; q = p - K v
                double s = 0.0;
                for (int i = lane; i < m; i += 32) s = fma(p[i], v[i], s);
                s = warp_sum(s);
                const double K = 0.5 * beta * s;
                for (int i = lane; i < m; i += 32) p[i] -= K * v[i];
            }
            __syncthreads();
            for (int e = t; e < m * m; e += nt) { // Asub -= v q' + q v'
                const int i = e / m, j = e - i * m;
                Asub[i * ld + j] -= v[i] * p[j] + p[i] * v[j];
            }
        }
        if (t == 0) { dv[k] = A[k * ld + k]; ev[k] = s_sc[1]; }
        __syncthreads();
    }
    if (t == 0) {
        dv[n - 2] = A[(n - 2) * ld + (n - 2)];
        dv[n - 1] = A[(n - 1) * ld + (n - 1)];
        ev[n - 2] = A[(n - 1) * ld + (n - 2)];
        double lo = 1e308, hi = -1e308;           // Gershgorin interval
        for (int i = 0; i < n; ++i) {
            const double r = (i > 0 ? fabs(ev[i - 1]) : 0.0) + (i < n - 1 ? fabs(ev[i]) : 0.0);
            lo = fmin(lo, dv[i] - r); hi = fmax(hi, dv[i] + r);
        }
        const double pad = 4e-16 * fmax(fabs(lo), fabs(hi)) + 1e-300;
        s_lo = lo - pad; s_hi = hi + pad;
    }
    __syncthreads();
    // multisection: count(x) = #eigenvalues < x ; lambda_max is where count steps from n-1 to n
    const double tiny = 1e-300;
    for (int iter = 0; iter < 16; ++iter) {
        const double lo = s_lo, hi = s_hi;
        if (!(hi - lo > 4.5e-16 * fmax(fabs(lo), fabs(hi)))) break;
        const double x = lo + (hi - lo) * ((double)(t + 1) / (double)(nt + 1));
        double q = dv[0] - x;
        int cnt = q < 0.0;
        for (int i = 1; i < n; ++i) {
            if (q == 0.0) q = tiny;
            q = dv[i] - x - ev[i - 1] * ev[i - 1] / q;
            cnt += q < 0.0;
        }
        if (t == 0) s_first = nt;
        __syncthreads();
        if (cnt == n) atomicMin(&s_first, t);     // smallest sample point that lies above every eigenvalue
        __syncthreads();
        const int f = s_first;
        __syncthreads();
        if (t == 0) {
            const double w = (hi - lo) / (double)(nt + 1);
            s_hi = (f < nt) ? lo + w * (double)(f + 1) : hi;
            s_lo = (f > 0) ? lo + w * (double)f : lo;
        }
        __syncthreads();
    }
    return 0.5 * (s_lo + s_hi);
}

}  // namespace vb
