// K11: `vbls!` for many small independent problems, one CTA per problem, one launch for the whole batch.
//
// Reference call pattern: examples/mil_util.jl:179-203 (vbls!: niter x { updateA!, updateCA!, updateSigma! } with BHat fixed,
// then updateYHat!) as used by classify(..., class_alg = "dual"/"vbls") (:453-535) on every test bag and every class model:
// thousands of L x M_p problems with L ~ 40, M_p = 5..40, H ~ 20.  With BHat fixed, B'B, B'Y and sum(Y.^2) are loop
// invariants, so the whole state of a problem lives in shared memory and registers for all niter iterations.
// Formulas: src/vbmf_sparse.jl:176-247 (updateA!, both covariance modes incl. Q2/Q3), :284-288 (updateCA!), :317-322
// (updateSigma!, homoscedastic); src/vbmf_dual.jl:322-351 for the two-group ARD update.
#include "kernels.cuh"
#include "linalg.cuh"
#include <algorithm>

namespace vb {

struct BatchDesc {
    int nprob, L, H, H0, kind, niter, full_cov, Mmax;
    const int* moff;        // [nprob+1] column offsets into the packed per-column arrays
    const double* Y;        // [L][Mtot] column-major, problems concatenated
    const double* B;        // [nprob][L*H] column-major L x H
    const double* SigmaB;   // [nprob][H*H]
    double* A;              // [Mtot][H]   in: unused (vbls overwrites), out: AHat rows
    double* CA;             // [Mtot*H]    in/out
    double* beta;           // [Mtot*H]    out
    double* sdiag;          // [Mtot*H]    out
    double* SigmaA;         // [nprob][H*H] out
    double* blocks;         // [Mtot][H][H] out or nullptr
    double* YHat;           // [L][Mtot] out or nullptr
    double* scal;           // [nprob][16]: 0 sigmaHat 1 eta 2 zeta 3 zeta0 4 trYTY 5 alpha (sparse) | M0 (trial) 6 beta0 (sparse)
                            //              7 8 = prior (alpha, beta) of ARD group 0, 9 10 = group 1, 14 15 = group 2 (trial)
                            //              out: 0 sigmaHat 2 zeta 11 12 6 = alpha of groups 0 1 2, 13 fail
                            //              dense: 0 sigma2 (in/out), 13 fail
};

template <int HP>
__global__ void __launch_bounds__(128, 3) batched_vbls_kernel(BatchDesc bd) {
    extern __shared__ __align__(16) double sm[];
    __shared__ double red[32];
    __shared__ double s_sig, s_fail;
    const int p = blockIdx.x;
    const int L = bd.L, H = bd.H, H0 = bd.H0;
    const int m0 = bd.moff[p], M = bd.moff[p + 1] - m0;
    const int t = threadIdx.x, nt = blockDim.x, lane = t & 31, warp = t >> 5, nw = nt >> 5;
    // shared layout
    double* Bs = sm;                         // [L][H]
    double* V = Bs + L * H;                  // [M][H]   V[m][h] = (B'Y)[h][m]
    double* As = V + bd.Mmax * H;            // [M][H]
    double* CAs = As + bd.Mmax * H;          // [M][H]
    double* Ss = CAs + bd.Mmax * H;          // [M][H]   diag of Sigma
    double* BtB = Ss + bd.Mmax * H;          // [H][H]
    double* G0 = BtB + H * H;                // [H][H]   B'B + L*SigmaB
    double* Gm = G0 + H * H;                 // [H][H]   sigmaHat*G0 (full_cov)
    double* SA = Gm + H * H;                 // [H][H]
    double* AtA = SA + H * H;                // [H][H]
    double* wacc = AtA + H * H;              // [nw][H*H] per-warp sums of Sigma_m
    double* colb = wacc + nw * H * H;        // [nw][64]
    colb += (colb - sm) & 1;                 // 16-byte aligned: the register inverse reads it as double2 (odd H would misalign it)
    double* pvb = colb + nw * 64;            // [nw][32]
    double* scal = bd.scal + (size_t)p * 16;
    const double* Y = bd.Y + (size_t)m0 * L;
    const double* Bg = bd.B + (size_t)p * L * H;
    const double* SBg = bd.SigmaB + (size_t)p * H * H;

    for (int e = t; e < L * H; e += nt) { const int h = e / L, l = e - h * L; Bs[l * H + h] = Bg[e]; }
    for (int e = t; e < M * H; e += nt) CAs[e] = bd.CA[(size_t)m0 * H + e];
    if (t == 0) { s_sig = scal[0]; s_fail = 0.0; }
    __syncthreads();
    for (int e = t; e < H * H; e += nt) {
        const int a = e / H, b = e - a * H;
        double s = 0.0;
        for (int l = 0; l < L; ++l) s = fma(Bs[l * H + a], Bs[l * H + b], s);
        BtB[e] = s;
        G0[e] = s + (double)L * SBg[e];
    }
    for (int e = t; e < M * H; e += nt) {
        const int m = e / H, h = e - m * H;
        double s = 0.0;
        for (int l = 0; l < L; ++l) s = fma(Y[(size_t)m * L + l], Bs[l * H + h], s);
        V[e] = s;
    }
    __syncthreads();
    // ARD groups (src/vbmf_dual.jl:322-351, src/vbmf_trial.jl:357-400): 0 = columns h < H0, 1 = h >= H0 and row m < M0,
    // 2 = h >= H0 and m >= M0 (trial only; dual has M0 = M, sparse H0 = H)
    const bool grouped = bd.kind == KIND_DUAL || bd.kind == KIND_TRIAL;
    const int M0p = bd.kind == KIND_TRIAL ? (int)scal[5] : M;
    const double eta = scal[1], zeta0 = scal[3], trYTY = scal[4];
    const double al0 = grouped ? scal[7] + 0.5 : scal[5], al1 = grouped ? scal[9] + 0.5 : scal[5], al2 = scal[14] + 0.5;
    const double be0 = grouped ? scal[8] : scal[6], be1 = grouped ? scal[10] : scal[6], be2 = scal[15];
    double zeta = scal[2];

    for (int it = 0; it < bd.niter; ++it) {
        const double sh = s_sig;
        // ---- updateA!
        if (bd.full_cov) {
            for (int e = t; e < H * H; e += nt) Gm[e] = sh * G0[e];
            __syncthreads();
            const bool live = lane < H;
            double* wa = wacc + warp * H * H;           // this warp's running sum of Sigma_m (lane owns column `lane`)
            if (live) for (int q = 0; q < H; ++q) wa[q * H + lane] = 0.0;
            double* col = colb + warp * 64;
            double* pv = pvb + warp * 32;
            bool all_ok = true;
            for (int m = warp; m < M; m += nw) {
                const double ca = live ? CAs[m * H + lane] : 1.0;
                double a[HP];
#pragma unroll
                for (int q = 0; q < HP; ++q) a[q] = ((q < H && live) ? Gm[q * H + lane] : 0.0) + ((q == lane) ? ca : 0.0);
                pv[lane] = live ? V[m * H + lane] : 0.0;
                const bool ok = warp_spd_inverse_reg<HP>(a, (live ? Gm[lane * H + lane] : 0.0) + ca, lane, col);
                all_ok = all_ok && ok;
                double s = 0.0, dg = 0.0;
#pragma unroll
                for (int q = 0; q < HP; ++q) {
                    s = fma(sh * a[q], pv[q], s);
                    if (q == lane) dg = a[q];
                }
                if (live) {
                    As[m * H + lane] = ok ? s : nan("");
                    Ss[m * H + lane] = ok ? dg : nan("");
                    if (bd.blocks != nullptr && it == bd.niter - 1) {
#pragma unroll
                        for (int q = 0; q < HP; ++q) if (q < H) bd.blocks[(size_t)(m0 + m) * H * H + q * H + lane] = a[q];
                    }
                }
                if (live) {
#pragma unroll
                    for (int q = 0; q < HP; ++q) if (q < H) wa[q * H + lane] += a[q];
                }
                __syncwarp();
            }
            if (!all_ok && lane == 0) s_fail = 1.0;
            __syncthreads();
            for (int e = t; e < H * H; e += nt) {
                double s = 0.0;
                for (int w = 0; w < nw; ++w) s += wacc[w * H * H + e];
                SA[e] = s;
            }
        } else {
            // diagonal path incl. Q2: element j (0-based) reads d[j] for j < H, else d[(j - H) / (M - 1)]
            for (int e = t; e < M * H; e += nt) {
                const int src = (e < H) ? e : (e - H) / max(M - 1, 1);
                const double dsrc = sh * BtB[src * H + src] + (double)L * SBg[src * H + src];
                const double s = 1.0 / (dsrc + CAs[e]);
                Ss[e] = s;
                As[e] = (sh * s) * V[e];
            }
            __syncthreads();
            for (int e = t; e < H * H; e += nt) {
                const int a = e / H, b = e - a * H;
                double s = 0.0;
                if (a == b) for (int m = 0; m < M; ++m) s += Ss[m * H + a];
                SA[e] = s;
            }
        }
        __syncthreads();
        // ---- updateCA!
        for (int e = t; e < M * H; e += nt) {
            const int m = e / H, h = e - m * H;
            const int g = (grouped && h >= H0) ? (m < M0p ? 1 : 2) : 0;
            const double a = As[e];
            const double beta = (g == 0 ? be0 : g == 1 ? be1 : be2) + 0.5 * (a * a + Ss[e]);
            CAs[e] = (g == 0 ? al0 : g == 1 ? al1 : al2) / beta;
            if (it == bd.niter - 1) bd.beta[(size_t)m0 * H + e] = beta;
        }
        // ---- updateSigma!
        for (int e = t; e < H * H; e += nt) {
            const int a = e / H, b = e - a * H;
            double s = 0.0;
            for (int m = 0; m < M; ++m) s = fma(As[m * H + a], As[m * H + b], s);
            AtA[e] = s;
        }
        double tr = 0.0;
        for (int e = t; e < M * H; e += nt) tr = fma(V[e], As[e], tr);
        tr = block_sum(tr, red);            // leading __syncthreads also publishes AtA
        double tt = 0.0;
        for (int e = t; e < H * H; e += nt) tt = fma(AtA[e] + SA[e], G0[e], tt);
        __syncthreads();
        if (t == 0) red[0] = tr;
        __syncthreads();
        tr = red[0];
        tt = block_sum(tt, red);
        if (t == 0) {
            zeta = zeta0 + 0.5 * trYTY - tr + 0.5 * tt;
            s_sig = eta / zeta;
        }
        __syncthreads();
    }
    // ---- outputs
    for (int e = t; e < M * H; e += nt) {
        bd.A[(size_t)m0 * H + e] = As[e];
        bd.CA[(size_t)m0 * H + e] = CAs[e];
        bd.sdiag[(size_t)m0 * H + e] = Ss[e];
    }
    for (int e = t; e < H * H; e += nt) bd.SigmaA[(size_t)p * H * H + e] = SA[e];
    if (t == 0) { scal[0] = s_sig; scal[2] = zeta; scal[11] = al0; scal[12] = al1; scal[6] = al2; scal[13] = s_fail; }
    if (bd.YHat != nullptr) {   // updateYHat!  src/vbmf_sparse.jl:275
        for (int e = t; e < L * M; e += nt) {
            const int m = e / L, l = e - m * L;
            double s = 0.0;
            for (int h = 0; h < H; ++h) s = fma(Bs[l * H + h], As[m * H + h], s);
            bd.YHat[(size_t)(m0 + m) * L + l] = s;
        }
    }
}

size_t batched_smem_bytes(int L, int H, int Mmax, int nwarps) {
    return (size_t)(L * H + 4 * Mmax * H + 5 * H * H + nwarps * (H * H + 96)) * sizeof(double) + 64;
}

int k_batched_vbls(cudaStream_t st, const BatchDesc& bd) {
    if (bd.nprob <= 0) return 0;
    if (bd.H > 32) { set_error("batched vbls supports H <= 32 (got %d)", bd.H); return -1; }
    const size_t smem = batched_smem_bytes(bd.L, bd.H, bd.Mmax, 4);
    if (smem > 200 * 1024) { set_error("problem too large for the one-CTA-per-problem path (%zu bytes of shared memory); use the solver API", smem); return -1; }
#define BLAUNCH(HPV)                                                                                              \
    {                                                                                                             \
        VB_CUDA_OK(cudaFuncSetAttribute(batched_vbls_kernel<HPV>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem)); \
        batched_vbls_kernel<HPV><<<bd.nprob, 128, smem, st>>>(bd);                                                \
    }
    if (bd.H <= 8) BLAUNCH(8) else if (bd.H <= 16) BLAUNCH(16) else if (bd.H <= 24) BLAUNCH(24) else BLAUNCH(32)
#undef BLAUNCH
    VB_LAUNCH_OK();
    return 0;
}


// Dense `vbmf_parameters` flavour of vbls! (examples/mil_util.jl:182-185, the class_alg = "vbls" pattern :470-478):
//   updateA!      SigmaA = sigma2*inv(B'B + L*SigmaB + sigma2*invCA);  A = ((Y'B)*SigmaA)/sigma2        src/vbmf.jl:95-102
//   updateCA!     CA[h,h] = ||A[:,h]||^2/M + SigmaA[h,h];  invCA = inv(CA)                               src/vbmf.jl:129-134
//   updateSigma2! sigma2 = (sum(Y.^2) - 2*tr(Y'B A') + tr((A'A + M*SigmaA)(B'B + L*SigmaB)))/(L*M)       src/vbmf.jl:153-157
// One H x H inverse per iteration (warp 0, register-resident Gauss-Jordan); B'B, V = B'Y and sum(Y.^2) are loop invariants.
struct BatchDenseDesc {
    int nprob, L, H, niter, Mmax;
    const int* moff;
    const double* Y;        // [L][Mtot]
    const double* B;        // [nprob][L*H]
    const double* SigmaB;   // [nprob][H*H]
    double* A;              // [Mtot][H] out
    double* SigmaA;         // [nprob][H*H] out
    double* icA;            // [nprob][H] in: diag(invCA), out: diag(invCA)
    double* cA;             // [nprob][H] out: diag(CA)
    double* YHat;           // [L][Mtot] out or nullptr
    double* scal;           // [nprob][16]
};

template <int HP>
__global__ void __launch_bounds__(128) batched_vbls_dense_kernel(BatchDenseDesc bd) {
    extern __shared__ __align__(16) double sm[];
    __shared__ double red[32];
    __shared__ __align__(16) double s_col[64];
    __shared__ double s_sig, s_fail;
    const int p = blockIdx.x;
    const int L = bd.L, H = bd.H;
    const int m0 = bd.moff[p], M = bd.moff[p + 1] - m0;
    const int t = threadIdx.x, nt = blockDim.x, lane = t & 31, warp = t >> 5;
    double* Bs = sm;                         // [L][H]
    double* V = Bs + L * H;                  // [M][H]
    double* As = V + bd.Mmax * H;            // [M][H]
    double* G0 = As + bd.Mmax * H;           // [H][H]  B'B + L*SigmaB
    double* SA = G0 + H * H;                 // [H][H]
    double* AtA = SA + H * H;                // [H][H]
    double* ic = AtA + H * H;                // [H]     diag(invCA)
    double* cd = ic + H;                     // [H]     diag(CA)
    double* scal = bd.scal + (size_t)p * 16;
    const double* Y = bd.Y + (size_t)m0 * L;
    const double* Bg = bd.B + (size_t)p * L * H;
    const double* SBg = bd.SigmaB + (size_t)p * H * H;

    for (int e = t; e < L * H; e += nt) { const int h = e / L, l = e - h * L; Bs[l * H + h] = Bg[e]; }
    for (int e = t; e < H; e += nt) { ic[e] = bd.icA[(size_t)p * H + e]; cd[e] = 0.0; }
    if (t == 0) { s_sig = scal[0]; s_fail = 0.0; }
    double y2 = 0.0;
    for (int e = t; e < L * M; e += nt) { const double y = Y[e]; y2 = fma(y, y, y2); }
    y2 = block_sum(y2, red);                 // includes the barrier that publishes Bs
    __syncthreads();
    if (t == 0) red[0] = y2;
    __syncthreads();
    const double trYTY = red[0];
    __syncthreads();
    for (int e = t; e < H * H; e += nt) {
        const int a = e / H, b = e - a * H;
        double s = 0.0;
        for (int l = 0; l < L; ++l) s = fma(Bs[l * H + a], Bs[l * H + b], s);
        G0[e] = s + (double)L * SBg[e];
    }
    for (int e = t; e < M * H; e += nt) {
        const int m = e / H, h = e - m * H;
        double s = 0.0;
        for (int l = 0; l < L; ++l) s = fma(Y[(size_t)m * L + l], Bs[l * H + h], s);
        V[e] = s;
    }
    __syncthreads();
    for (int it = 0; it < bd.niter; ++it) {
        const double s2 = s_sig;
        // ---- updateA!: SigmaA
        if (warp == 0) {
            const bool live = lane < H;
            double a[HP];
#pragma unroll
            for (int q = 0; q < HP; ++q)
                a[q] = ((q < H && live) ? G0[q * H + lane] : 0.0) + ((q == lane) ? (live ? s2 * ic[lane] : 1.0) : 0.0);
            const double dg = live ? G0[lane * H + lane] + s2 * ic[lane] : 1.0;
            const bool ok = warp_spd_inverse_reg<HP>(a, dg, lane, s_col);
            if (live) {
#pragma unroll
                for (int q = 0; q < HP; ++q) if (q < H) SA[q * H + lane] = ok ? s2 * a[q] : nan("");
            }
            if (!ok && lane == 0) s_fail = 1.0;
        }
        __syncthreads();
        // ---- A = ((Y'B) * SigmaA) / sigma2
        for (int e = t; e < M * H; e += nt) {
            const int m = e / H, h = e - m * H;
            double s = 0.0;
            for (int k = 0; k < H; ++k) s = fma(V[m * H + k], SA[k * H + h], s);
            As[e] = s / s2;
        }
        __syncthreads();
        // ---- updateCA! and the Gram of A
        for (int e = t; e < H * H; e += nt) {
            const int a = e / H, b = e - a * H;
            double s = 0.0;
            for (int m = 0; m < M; ++m) s = fma(As[m * H + a], As[m * H + b], s);
            AtA[e] = s;
            if (a == b) { const double c = s / (double)M + SA[e]; cd[a] = c; ic[a] = 1.0 / c; }
        }
        // ---- updateSigma2!
        double tr = 0.0;
        for (int e = t; e < M * H; e += nt) tr = fma(V[e], As[e], tr);
        tr = block_sum(tr, red);            // leading barrier also publishes AtA
        __syncthreads();
        if (t == 0) red[0] = tr;
        __syncthreads();
        tr = red[0];
        __syncthreads();
        double tt = 0.0;
        for (int e = t; e < H * H; e += nt) tt = fma(AtA[e] + (double)M * SA[e], G0[e], tt);
        tt = block_sum(tt, red);
        if (t == 0) s_sig = (trYTY - 2.0 * tr + tt) / ((double)L * (double)M);
        __syncthreads();
    }
    for (int e = t; e < M * H; e += nt) bd.A[(size_t)m0 * H + e] = As[e];
    for (int e = t; e < H * H; e += nt) bd.SigmaA[(size_t)p * H * H + e] = SA[e];
    for (int e = t; e < H; e += nt) { bd.icA[(size_t)p * H + e] = ic[e]; bd.cA[(size_t)p * H + e] = cd[e]; }
    if (t == 0) { scal[0] = s_sig; scal[13] = s_fail; }
    if (bd.YHat != nullptr) {
        for (int e = t; e < L * M; e += nt) {
            const int m = e / L, l = e - m * L;
            double s = 0.0;
            for (int h = 0; h < H; ++h) s = fma(Bs[l * H + h], As[m * H + h], s);
            bd.YHat[(size_t)(m0 + m) * L + l] = s;
        }
    }
}

int k_batched_vbls_dense(cudaStream_t st, const BatchDenseDesc& bd) {
    if (bd.nprob <= 0) return 0;
    if (bd.H > 32) { set_error("batched vbls supports H <= 32 (got %d)", bd.H); return -1; }
    const size_t smem = (size_t)(bd.L * bd.H + 2 * bd.Mmax * bd.H + 3 * bd.H * bd.H + 2 * bd.H) * sizeof(double) + 64;
    if (smem > 200 * 1024) { set_error("problem too large for the one-CTA-per-problem path (%zu bytes of shared memory); use the solver API", smem); return -1; }
#define BDLAUNCH(HPV)                                                                                             \
    {                                                                                                             \
        VB_CUDA_OK(cudaFuncSetAttribute(batched_vbls_dense_kernel<HPV>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem)); \
        batched_vbls_dense_kernel<HPV><<<bd.nprob, 128, smem, st>>>(bd);                                          \
    }
    if (bd.H <= 8) BDLAUNCH(8) else if (bd.H <= 16) BDLAUNCH(16) else if (bd.H <= 24) BDLAUNCH(24) else BDLAUNCH(32)
#undef BDLAUNCH
    VB_LAUNCH_OK();
    return 0;
}

}  // namespace vb
