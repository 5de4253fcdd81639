// Standalone race hunt for K1: compile with -DVARIANT=n
#include <cstdarg>
#include "../../vbmatrixfactorization.jl_b200/csrc/gemm_dmma.cu"
#include <vector>
#include <cstdio>
namespace vb { void set_error(const char* fmt, ...) { va_list ap; va_start(ap, fmt); vprintf(fmt, ap); va_end(ap); printf("\n"); } void count_launch() {} }
#include <cstdarg>
__global__ void fill(double* x, size_t n, unsigned seed) {
    for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
        unsigned long long z = (i + 1) * 0x9E3779B97F4A7C15ull + seed; z ^= z >> 31; z *= 0xBF58476D1CE4E5B9ull; z ^= z >> 29;
        x[i] = (double)(z >> 11) * (1.0 / 9007199254740992.0) - 0.5;
    }
}
__global__ void cmp(const double* a, const double* b, size_t n, int H, unsigned long long* nbad, double* maxerr, int* firstbad) {
    for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
        double e = fabs(a[i] - b[i]);
        if (e > 1e-9) { unsigned long long k = atomicAdd(nbad, 1ull); if (k < 16) firstbad[k] = (int)(i / H); }
    }
}
int main() {
    using namespace vb;
    int cases[4][3] = {{3000, 50000, 64}, {10000, 100000, 32}, {20000, 100000, 64}, {20000, 25000, 64}};
    for (auto& c : cases) {
        int L = c[0], M = c[1], H = c[2];
        double *Y, *B, *P, *R; unsigned long long* nbad; double* maxerr; int* fb;
        cudaMalloc(&Y, (size_t)L * M * 8); cudaMalloc(&B, (size_t)L * H * 8); cudaMalloc(&P, (size_t)M * H * 8); cudaMalloc(&R, (size_t)M * H * 8);
        cudaMalloc(&nbad, 8); cudaMalloc(&maxerr, 8); cudaMalloc(&fb, 64);
        fill<<<1024, 256>>>(Y, (size_t)L * M, 1); fill<<<64, 256>>>(B, (size_t)L * H, 2);
        launch_gemm_ytb_simt(0, Y, L, B, L, R, M, L, H, H, nullptr);
        CUtensorMap tmY, tmB;
        make_tmap_2d(&tmY, Y, L, M, (uint64_t)L * 8, 16, 128);
        make_tmap_2d(&tmB, B, L, H, (uint64_t)L * 8, 16, gemm_geometry(H).bn);
        int S1 = 1, kbs = 1; plan_splitk_ytb(L, M, H, 148, &S1, &kbs);
        double* Pp; cudaMalloc(&Pp, (size_t)S1 * M * H * 8);
        for (int rep = 0; rep < 2; ++rep) {
            cudaMemset(P, 0, (size_t)M * H * 8); cudaMemset(nbad, 0, 8);
            launch_gemm_ytb(0, &tmY, &tmB, P, M, L, H, H, 1, (L + 15) / 16, 0, nullptr, 148);
            cmp<<<1024, 256>>>(P, R, (size_t)M * H, H, nbad, maxerr, fb);
            unsigned long long h; int hfb[16];
            cudaMemcpy(&h, nbad, 8, cudaMemcpyDeviceToHost); cudaMemcpy(hfb, fb, 64, cudaMemcpyDeviceToHost);
            printf("variant %d  L=%d M=%d H=%d rep %d: bad entries %llu", VARIANT, L, M, H, rep, h);
            if (h) { printf("  rows:"); for (int i = 0; i < (h < 8 ? (int)h : 8); ++i) printf(" %d(t%d,r%d)", hfb[i], hfb[i] / 128, hfb[i] % 128); }
            printf("  %s\n", cudaGetErrorString(cudaDeviceSynchronize()));
        }
        {
            cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
            cudaEventRecord(e0);
            for (int rep = 0; rep < 5; ++rep) launch_gemm_ytb(0, &tmY, &tmB, Pp, M, L, H, H, S1, kbs, (size_t)M * H, nullptr, 148);
            cudaEventRecord(e1); cudaEventSynchronize(e1);
            float ms; cudaEventElapsedTime(&ms, e0, e1); ms /= 5;
            printf("wm64 %d wm32 %d S %d  L=%d M=%d H=%d  K1 %.3f ms  %.2f TFLOP/s\n", K1_WM64, K1_WM32, S1, L, M, H, ms, 2.0 * L * M * H / ms * 1e-9);
        }
        cudaFree(Y); cudaFree(B); cudaFree(P); cudaFree(R);
    }
    return 0;
}
