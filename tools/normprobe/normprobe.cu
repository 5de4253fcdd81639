// Timing probe for sym_lambda_max (phases via clock64): nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o normprobe normprobe.cu
#include <cstdarg>
#include <cstdio>
#include <vector>
#include <cmath>
#include "../../vbmatrixfactorization.jl_b200/csrc/linalg.cuh"
namespace vb { void set_error(const char*, ...) {} void count_launch() {} }
__global__ void probe(const double* G, int H, double* out, long long* clk) {
    extern __shared__ double sm[];
    const int ld = H + 1, t = threadIdx.x;
    double* Mx = sm; double* vec = sm + H * ld;
    long long c0 = clock64();
    for (int e = t; e < H * H; e += blockDim.x) { const int i = e / H, j = e - i * H; Mx[i * ld + j] = 0.5 * (G[i * H + j] + G[j * H + i]); }
    __syncthreads();
    long long c1 = clock64();
    double v = vb::sym_lambda_max(Mx, ld, H, vec);
    long long c2 = clock64();
    if (t == 0) { out[0] = v; clk[0] = c1 - c0; clk[1] = c2 - c1; }
}
int main() {
    for (int H : {32, 64, 128}) {
        std::vector<double> X(2000 * H), G(H * H, 0.0);
        unsigned s = 1; for (auto& x : X) { s = s * 1664525u + 1013904223u; x = (double)(s >> 8) / (1 << 24) - 0.5; }
        for (int i = 0; i < H; ++i) for (int j = 0; j < H; ++j) { double a = 0; for (int l = 0; l < 2000; ++l) a += X[l * H + i] * X[l * H + j]; G[i * H + j] = a; }
        double *dG, *dout; long long* dclk;
        cudaMalloc(&dG, H * H * 8); cudaMalloc(&dout, 8); cudaMalloc(&dclk, 16);
        cudaMemcpy(dG, G.data(), H * H * 8, cudaMemcpyHostToDevice);
        size_t smem = ((H + 2) * (H + 2) + 4 * (H + 2) + 8) * 8;
        cudaFuncSetAttribute(probe, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        for (int rep = 0; rep < 3; ++rep) probe<<<1, 512, smem>>>(dG, H, dout, dclk);
        cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
        cudaEventRecord(e0); for (int rep = 0; rep < 20; ++rep) probe<<<1, 512, smem>>>(dG, H, dout, dclk); cudaEventRecord(e1); cudaEventSynchronize(e1);
        float ms; cudaEventElapsedTime(&ms, e0, e1);
        double out; long long clk[2];
        cudaMemcpy(&out, dout, 8, cudaMemcpyDeviceToHost); cudaMemcpy(clk, dclk, 16, cudaMemcpyDeviceToHost);
        printf("H=%d lambda_max=%.15e load=%lld cyc  sym_lambda_max=%lld cyc  kernel=%.1f us  (%s)\n", H, out, clk[0], clk[1], ms / 20 * 1e3, cudaGetErrorString(cudaGetLastError()));
    }
    return 0;
}
