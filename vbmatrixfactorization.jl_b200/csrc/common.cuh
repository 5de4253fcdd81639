// Shared device/host helpers for the vbmf_b200 CUDA library (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <cuda.h>
#include <cstdint>
#include <cstdio>
#include <cmath>

namespace vb {

// ---------------------------------------------------------------------------------------------------------
// Device-resident control block + scalars.  Every hot-path kernel starts with `if (!sc->active) return;`
// so the host can enqueue iterations ahead of the device-side convergence test (no per-iteration host sync).
// ---------------------------------------------------------------------------------------------------------
struct Scalars {
    // loop control (vbmf! / vbmf_sparse! / vbmf_dual! while-loop, src/vbmf.jl:193, src/vbmf_sparse.jl:368)
    int active;       // 1 while (i <= niter && d > eps)
    int iter;         // iterations completed
    int niter;
    int norm_mode;    // 0 spectral (Julia 0.5 norm(::Matrix)), 1 frobenius
    double eps;
    double d;         // last delta
    double normBold;  // norm(old) carried between iterations
    double normD, normBnew;   // norm(BHat - old), norm(BHat) of the current iteration (norms_kernel)
    // model scalars
    double sigma2;                     // dense: noise variance
    double sigmaHat, eta, zeta;        // sparse/dual: noise precision posterior
    double eta0, zeta0;
    double alpha0p, beta0p;            // sparse: hyper-prior (alpha0, beta0)
    double alpha;                      // sparse: alpha = alpha0 + 1/2
    double gamma0, delta0, gamma;
    double alpha00, beta00, alpha01, beta01;   // dual / trial: learned hyper-priors of groups 0, 1
    double alpha02, beta02;                    // trial: third group (rows >= M0 of the columns >= H0)
    double alpha_g0, alpha_g1, alpha_g2;       // posterior shapes alpha_g = alpha0g + 1/2
    double trYTY;                      // global sum(Y.^2)
    double meanSigmaVec;               // mean(sigmaVecHat), diag_var
    double trBQ;                       // sum(BHat .* (Y*AHat))
    double lb;                         // last lower bound
    int    chol_fail;                  // sticky: bit 0: a pivot was <= 0 / NaN in an SPD inverse; bit 2: a peer-exchange barrier timed out
    int    pad_;
};

#define VB_CUDA_OK(call)                                                                       \
    do {                                                                                       \
        cudaError_t e__ = (call);                                                              \
        if (e__ != cudaSuccess) {                                                              \
            (void)cudaGetLastError();                                                          \
            vb::set_error("%s failed: %s (%s:%d)", #call, cudaGetErrorString(e__), __FILE__, __LINE__); \
            return -1;                                                                         \
        }                                                                                      \
    } while (0)

void set_error(const char* fmt, ...);
void count_launch();

// after every kernel launch: count it (bench.py reports the number) and surface launch-configuration errors
#define VB_LAUNCH_OK()                                                                         \
    do {                                                                                       \
        vb::count_launch();                                                                    \
        cudaError_t e__ = cudaGetLastError();                                                  \
        if (e__ != cudaSuccess) {                                                              \
            vb::set_error("kernel launch failed: %s (%s:%d)", cudaGetErrorString(e__), __FILE__, __LINE__); \
            return -1;                                                                         \
        }                                                                                      \
    } while (0)

// ---------------------------------------------------------------------------------------------------------
// small device helpers
// ---------------------------------------------------------------------------------------------------------
__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

// deterministic block sum (fixed tree); result valid in thread 0. `red` must hold >= 32 doubles.
__device__ __forceinline__ double block_sum(double v, double* red) {
    v = warp_sum(v);
    int w = threadIdx.x >> 5, l = threadIdx.x & 31;
    __syncthreads();
    if (l == 0) red[w] = v;
    __syncthreads();
    if (w == 0) {
        int nw = (blockDim.x + 31) >> 5;
        v = (l < nw) ? red[l] : 0.0;
        v = warp_sum(v);
    }
    return v;
}

// digamma for x > 0: upward recurrence to x >= 10, then the asymptotic series (|err| ~ 1e-16 relative)
__device__ __host__ inline double digamma_pos(double x) {
    double r = 0.0;
    while (x < 10.0) { r -= 1.0 / x; x += 1.0; }
    double xi = 1.0 / x, x2 = xi * xi;
    double s = x2 * (1.0 / 12 - x2 * (1.0 / 120 - x2 * (1.0 / 252 - x2 * (1.0 / 240 - x2 * (1.0 / 132 - x2 * (691.0 / 32760 - x2 * (1.0 / 12)))))));
    return r + log(x) - 0.5 * xi - s;
}

__device__ __host__ inline double trigamma_pos(double x) {
    double r = 0.0;
    while (x < 10.0) { r += 1.0 / (x * x); x += 1.0; }
    double xi = 1.0 / x, x2 = xi * xi;
    return r + xi * (1.0 + xi * (0.5 + xi * (1.0 / 6 - x2 * (1.0 / 30 - x2 * (1.0 / 42 - x2 * (1.0 / 30 - x2 * (5.0 / 66)))))));
}

}  // namespace vb
