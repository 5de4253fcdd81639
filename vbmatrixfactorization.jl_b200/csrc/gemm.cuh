// Launcher interface of the two FP64 tensor-core (DMMA) contractions of the VB loop.
#pragma once
#include "common.cuh"

namespace vb {

// Encode a rank-2 FP64 tensor map (dim0 contiguous) with 128-byte swizzle and zero OOB fill.
int make_tmap_2d(CUtensorMap* out, const void* base, uint64_t dim0, uint64_t dim1, uint64_t stride1_bytes,
                 uint32_t box0, uint32_t box1);

struct GemmGeometry {
    int bn;          // N tile (32 / 64 / 128), smallest >= H
    int ctas_per_sm;
    int stages;
};
GemmGeometry gemm_geometry(int H);
// opt the contraction kernels into their dynamic shared memory on the CURRENT device (once per context)
int gemm_init_device();

// K1: P[m][h] = sum_l Y[l, m] * B[l, h]      (Y' * BHat, src/vbmf.jl:98, src/vbmf_sparse.jl:195,232)
// tmY : dims {L, M}, box {16, 128};  tmB : dims {L, H} (column h contiguous in l), box {16, bn}
// P is row-major [M][ldP].
// With S > 1 the kernel writes S slabs P + s*slab_stride (split over L, kb_per_split 16-row blocks each) that the caller
// sums in fixed order (deterministic split-K); S == 1 writes P directly.
int launch_gemm_ytb(cudaStream_t st, const CUtensorMap* tmY, const CUtensorMap* tmB, double* P,
                    int M, int L, int H, int ldP, int S, int kb_per_split, size_t slab_stride, const Scalars* sc, int num_sms);
void plan_splitk_ytb(int L, int M, int H, int num_sms, int* S, int* kb_per_split);

// K2: Qpart[s][h][l] = sum_{m in chunk s} Y[l, m] * A[m][h]   (Y * AHat, src/vbmf.jl:112, src/vbmf_sparse.jl:266)
// tmY : dims {L, M}, box {16, 16};  tmA : dims {H, M} (row m contiguous in h), box {16, 16}
// Qpart is S slabs of column-major [H][ldQ]; the caller reduces the slabs in fixed order (deterministic split-K).
int launch_gemm_ya(cudaStream_t st, const CUtensorMap* tmY, const CUtensorMap* tmA, double* Qpart,
                   int L, int M, int H, int ldQ, int kchunk, int S, const Scalars* sc, int num_sms);

// K2 with the stream-K decomposition (gemm_dmma.cu): Qpart holds smax partial slabs [H][ldQ]; launch_reduce_q_sk sums the partials
// of every row tile in fixed CTA order into Q.  plan_streamk: unit length kq, units per tile nk, grid, worst-case partials per tile.
int plan_streamk(int L, int M, int H, int num_sms, int* kq, int* nk, int* grid, int* smax);
int launch_gemm_ya_sk(cudaStream_t st, const CUtensorMap* tmY, const CUtensorMap* tmA, double* Qpart, double* Q, int L, int M, int H,
                      int ldQ, int kq, int nk, int grid, const Scalars* sc);
int launch_reduce_q_sk(cudaStream_t st, const double* Qpart, double* Q, int L, int M, int H, int ldQ, int nk, int grid, const Scalars* sc);

// split-K plan for K2: number of slabs and chunk length (multiple of 16)
void plan_splitk(int L, int M, int H, int num_sms, int* S, int* kchunk);

// Plain SIMT FP64 versions (debug / cross-check only; selected with VBMF_B200_GEMM=simt).
int launch_gemm_ytb_simt(cudaStream_t st, const double* Y, int ldY, const double* B, int ldB, double* P,
                         int M, int L, int H, int ldP, const Scalars* sc);
int launch_gemm_ya_simt(cudaStream_t st, const double* Y, int ldY, const double* A, double* Qpart,
                        int L, int M, int H, int ldQ, int kchunk, int S, const Scalars* sc);

}  // namespace vb
