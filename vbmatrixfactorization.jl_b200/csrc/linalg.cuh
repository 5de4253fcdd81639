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

// Largest eigenvalue of the symmetric matrix A (n x n, n even, row stride ld, shared memory; destroyed).
// rot = 4*(n/2) doubles of scratch, red = 32 doubles.  Whole CTA cooperates.  Result valid in every thread.
__device__ inline double jacobi_lambda_max(double* A, int ld, int n, double* rot, double* red) {
    const int t = threadIdx.x, nt = blockDim.x;
    const int half = n / 2;
    __shared__ double s_off, s_diag;
    if (n == 0) return 0.0;
    for (int sweep = 0; sweep < 30; ++sweep) {
        double off = 0.0, dg = 0.0;
        for (int e = t; e < n * n; e += nt) {
            const int i = e / n, j = e - i * n;
            const double v = A[i * ld + j];
            if (i == j) dg += v * v; else off += v * v;
        }
        off = block_sum(off, red);
        if (t == 0) s_off = off;
        dg = block_sum(dg, red);
        if (t == 0) s_diag = dg;
        __syncthreads();
        if (!(s_off > 1e-32 * s_diag)) break;      // also leaves on NaN
        for (int round = 0; round < n - 1; ++round) {
            if (t < half) {
                int p, q;
                if (t == 0) { p = n - 1; q = round; }
                else { p = (round + t) % (n - 1); q = (round - t + (n - 1)) % (n - 1); }
                const double app = A[p * ld + p], aqq = A[q * ld + q], apq = A[p * ld + q];
                double c = 1.0, s = 0.0;
                if (apq != 0.0) {
                    const double tau = (aqq - app) / (2.0 * apq);
                    const double tt = (tau >= 0.0 ? 1.0 : -1.0) / (fabs(tau) + sqrt(1.0 + tau * tau));
                    c = 1.0 / sqrt(1.0 + tt * tt);
                    s = tt * c;
                    if (!isfinite(tau)) { c = 1.0; s = 0.0; }
                }
                rot[4 * t] = c; rot[4 * t + 1] = s; rot[4 * t + 2] = (double)p; rot[4 * t + 3] = (double)q;
            }
            __syncthreads();
            for (int e = t; e < half * n; e += nt) {     // columns p, q of every row
                const int i = e / n, r = e - i * n;
                const double c = rot[4 * i], s = rot[4 * i + 1];
                const int p = (int)rot[4 * i + 2], q = (int)rot[4 * i + 3];
                const double arp = A[r * ld + p], arq = A[r * ld + q];
                A[r * ld + p] = c * arp - s * arq;
                A[r * ld + q] = s * arp + c * arq;
            }
            __syncthreads();
            for (int e = t; e < half * n; e += nt) {     // rows p, q of every column
                const int i = e / n, r = e - i * n;
                const double c = rot[4 * i], s = rot[4 * i + 1];
                const int p = (int)rot[4 * i + 2], q = (int)rot[4 * i + 3];
                const double apr = A[p * ld + r], aqr = A[q * ld + r];
                A[p * ld + r] = c * apr - s * aqr;
                A[q * ld + r] = s * apr + c * aqr;
            }
            __syncthreads();
        }
    }
    double mx = -1e308;
    for (int i = t; i < n; i += nt) mx = fmax(mx, A[i * ld + i]);
    // NaN-propagating max reduce
    bool bad = false;
    for (int i = t; i < n; i += nt) if (A[i * ld + i] != A[i * ld + i]) bad = true;
    __syncthreads();
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) mx = fmax(mx, __shfl_xor_sync(0xffffffffu, mx, o));
    const int any_bad = __syncthreads_or(bad ? 1 : 0);
    if ((t & 31) == 0) red[t >> 5] = mx;
    __syncthreads();
    double r = -1e308;
    for (int w = 0; w < (nt + 31) / 32; ++w) r = fmax(r, red[w]);
    __syncthreads();
    return any_bad ? nan("") : r;
}

}  // namespace vb
