// FP64 issue-rate microbenchmark for B200 (sm_100a): DMMA (mma.sync m8n8k4 f64) vs DFMA.
// Establishes the FP64 roofline denominator that MEASURED_PEAKS.json does not carry.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o fp64_peak fp64_peak.cu
#include <cstdio>
#include <cstdlib>
#include <cuda_runtime.h>

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %d\n", cudaGetErrorString(e), __LINE__); exit(1);} } while (0)

template <int NACC>
__global__ void dmma_kernel(double* out, int iters, double a0, double b0) {
    double c[NACC][2];
#pragma unroll
    for (int i = 0; i < NACC; ++i) { c[i][0] = 0.0; c[i][1] = 0.0; }
    double a = a0 + threadIdx.x * 1e-9, b = b0;
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < NACC; ++i) {
            asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n"
                         : "+d"(c[i][0]), "+d"(c[i][1]) : "d"(a), "d"(b));
        }
    }
    double s = 0;
#pragma unroll
    for (int i = 0; i < NACC; ++i) s += c[i][0] + c[i][1];
    if (s == 123.456) out[0] = s;
}

template <int NACC>
__global__ void dfma_kernel(double* out, int iters, double a0, double b0) {
    double c[NACC];
#pragma unroll
    for (int i = 0; i < NACC; ++i) c[i] = i;
    double a = a0 + threadIdx.x * 1e-9, b = b0;
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < NACC; ++i) c[i] = fma(a, c[i], b);
    }
    double s = 0;
#pragma unroll
    for (int i = 0; i < NACC; ++i) s += c[i];
    if (s == 123.456) out[0] = s;
}

// mixed: per loop NACC dmma + NF dfma
template <int NACC, int NF>
__global__ void mixed_kernel(double* out, int iters, double a0, double b0) {
    double c[NACC][2]; double f[NF];
#pragma unroll
    for (int i = 0; i < NACC; ++i) { c[i][0] = 0.0; c[i][1] = 0.0; }
#pragma unroll
    for (int i = 0; i < NF; ++i) f[i] = i;
    double a = a0 + threadIdx.x * 1e-9, b = b0;
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < NACC; ++i) {
            asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n"
                         : "+d"(c[i][0]), "+d"(c[i][1]) : "d"(a), "d"(b));
            if (i < NF) f[i] = fma(a, f[i], b);
        }
    }
    double s = 0;
#pragma unroll
    for (int i = 0; i < NACC; ++i) s += c[i][0] + c[i][1];
#pragma unroll
    for (int i = 0; i < NF; ++i) s += f[i];
    if (s == 123.456) out[0] = s;
}

template <typename F>
float time_ms(F launch, int reps) {
    cudaEvent_t e0, e1; CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
    launch(); CK(cudaDeviceSynchronize());
    CK(cudaEventRecord(e0));
    for (int r = 0; r < reps; ++r) launch();
    CK(cudaEventRecord(e1)); CK(cudaEventSynchronize(e1));
    float ms; CK(cudaEventElapsedTime(&ms, e0, e1));
    return ms / reps;
}

int main(int argc, char** argv) {
    int dev = 0; CK(cudaSetDevice(dev));
    cudaDeviceProp p; CK(cudaGetDeviceProperties(&p, dev));
    int sms = p.multiProcessorCount;
    int clk_khz = 0; cudaDeviceGetAttribute(&clk_khz, cudaDevAttrClockRate, dev);
    printf("{\"device\": \"%s\", \"sms\": %d, \"clock_khz_attr\": %d}\n", p.name, sms, clk_khz);
    double* out; CK(cudaMalloc(&out, 8));
    const int iters = 20000;
    // DMMA: sweep warps per SM and accumulators
    int warps_list[] = {4, 8, 16, 32};
    for (int wi = 0; wi < 4; ++wi) {
        int warps = warps_list[wi];
        int threads = warps * 32 > 1024 ? 1024 : warps * 32;
        int blocks = sms * (warps * 32 / threads);
#define RUN_DMMA(N) { float ms = time_ms([&]{ dmma_kernel<N><<<blocks, threads>>>(out, iters, 1.0, 1.0); }, 3); \
        double flops = 2.0 * 256 * (double)N * iters * warps * sms; \
        printf("{\"bench\": \"dmma\", \"warps_per_sm\": %d, \"nacc\": %d, \"ms\": %.4f, \"tflops\": %.3f}\n", warps, N, ms, flops / ms * 1e-9); }
        RUN_DMMA(1) RUN_DMMA(2) RUN_DMMA(4) RUN_DMMA(8) RUN_DMMA(16) RUN_DMMA(32)
#define RUN_DFMA(N) { float ms = time_ms([&]{ dfma_kernel<N><<<blocks, threads>>>(out, iters, 1.0, 1.0); }, 3); \
        double flops = 2.0 * 32 * (double)N * iters * warps * sms; \
        printf("{\"bench\": \"dfma\", \"warps_per_sm\": %d, \"nacc\": %d, \"ms\": %.4f, \"tflops\": %.3f}\n", warps, N, ms, flops / ms * 1e-9); }
        RUN_DFMA(1) RUN_DFMA(4) RUN_DFMA(8) RUN_DFMA(16)
        { float ms = time_ms([&]{ mixed_kernel<16, 8><<<blocks, threads>>>(out, iters, 1.0, 1.0); }, 3);
          double flops = 2.0 * (256 * 16 + 32 * 8) * (double)iters * warps * sms;
          printf("{\"bench\": \"mixed16d8f\", \"warps_per_sm\": %d, \"ms\": %.4f, \"tflops\": %.3f}\n", warps, ms, flops / ms * 1e-9); }
        { float ms = time_ms([&]{ mixed_kernel<16, 16><<<blocks, threads>>>(out, iters, 1.0, 1.0); }, 3);
          double flops = 2.0 * (256 * 16 + 32 * 16) * (double)iters * warps * sms;
          printf("{\"bench\": \"mixed16d16f\", \"warps_per_sm\": %d, \"ms\": %.4f, \"tflops\": %.3f}\n", warps, ms, flops / ms * 1e-9); }
    }
    // sustained DMMA for ~3 s at 8 warps/SM, 16 acc: report per-0.25 s window
    {
        int warps = 8, threads = 256, blocks = sms;
        const int it2 = 200000;
        for (int w = 0; w < 12; ++w) {
            float ms = time_ms([&]{ dmma_kernel<16><<<blocks, threads>>>(out, it2, 1.0, 1.0); }, 2);
            double flops = 2.0 * 256 * 16.0 * it2 * warps * sms;
            printf("{\"bench\": \"dmma_sustained\", \"window\": %d, \"ms\": %.3f, \"tflops\": %.3f}\n", w, ms, flops / ms * 1e-9);
        }
    }
    return 0;
}
