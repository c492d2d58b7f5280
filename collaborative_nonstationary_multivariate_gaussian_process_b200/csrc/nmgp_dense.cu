// Large dense FP64 kernels of the SIM_code (exact/Kronecker) line:
//   nmgp_gemm_nt      C = alpha A B^T + beta C on the FP64 tensor cores (DMMA m8n8k4), 128x128 CTA tiles
//   nmgp_potrf_big    blocked right-looking Cholesky (lower) of a T x T matrix: diagonal block in one CTA,
//                     panel TRSM one thread per row, trailing SYRK through the DMMA GEMM (lower tiles only)
//   nmgp_potrs_vec    solve L L^T x = b for one right-hand side (blocked substitution)
//   nmgp_eigh_small   cyclic Jacobi eigen-decomposition of a small symmetric matrix (the D x D output covariance)
//   helpers           A = alpha K + sigma2 I, Kronecker products
// They replace torch.symeig / torch.inverse / torch.logdet / torch.mm at
// code/SIM_code/Utility/kronecker_operation.py:36-85 and distributions.py:26-113 (SURVEY.md 7.2 "Kronecker without
// eigen of K": sigma2 I + B (x) K = (V (x) I) blkdiag_m(sigma2 I + lambda_m K) (V^T (x) I)).
#include <stdlib.h>

#include "common.cuh"

__device__ __forceinline__ void dmma884d(double& d0, double& d1, double a, double b) {
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n"
                 : "+d"(d0), "+d"(d1)
                 : "d"(a), "d"(b));
}
__device__ __forceinline__ void cpd8(double* smem_dst, const double* gsrc) {
    unsigned sa = (unsigned)__cvta_generic_to_shared(smem_dst);
    asm volatile("cp.async.ca.shared.global [%0], [%1], 8;\n" ::"r"(sa), "l"(gsrc));
}
__device__ __forceinline__ void cpd16(double* smem_dst, const double* gsrc) {
    unsigned sa = (unsigned)__cvta_generic_to_shared(smem_dst);
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;\n" ::"r"(sa), "l"(gsrc));
}
__device__ __forceinline__ void cpd_commit() { asm volatile("cp.async.commit_group;\n" ::); }
__device__ __forceinline__ void cpd_wait0() { asm volatile("cp.async.wait_group 0;\n" ::); }

// ------------------------------------------------------------------------------------------------------------
// C[M,N] = alpha * A[M,K] * B[N,K]^T + beta * C      (row-major, leading dimensions lda/ldb/ldc)
// CTA tile 128 x 128, 8 warps as 4 (m) x 2 (n): warp tile 32 x 64 = 4 x 8 DMMA blocks, K staged 32 at a time in a
// cp.async double buffer.  lower_only: skip tiles strictly above the diagonal (SYRK-style trailing update).
#define GT_M 128
#define GT_N 128
#define GT_K 32
#define GT_LD 36          // 36 % 8 == 4: conflict-free fragment loads (see pad4mod8 in nmgp_quadform_mma.cu)
#define GT_THREADS 256

template <int NWN>
__global__ void __launch_bounds__(128 * NWN, NWN == 2 ? 1 : 2)
k_gemm_nt(const double* __restrict__ A, const double* __restrict__ Bm, double* __restrict__ C, long long M,
          long long N, long long K, long long lda, long long ldb, long long ldc, double alpha, double beta,
          int lower_only) {
    extern __shared__ __align__(16) double sm[];
    double* As = sm;                              // [2][GT_M][GT_LD]
    constexpr int TN = 64 * NWN, NTHREADS = 128 * NWN;
    double* Bs = As + 2 * GT_M * GT_LD;           // [2][TN][GT_LD]
    const long long m0 = (long long)blockIdx.y * GT_M, n0 = (long long)blockIdx.x * TN;
    if (lower_only && n0 > m0 + GT_M - 1) return;
    const int tid = threadIdx.x, lane = tid & 31, w = tid >> 5, g = lane >> 2, t = lane & 3;
    const int wm = w / NWN, wn = w % NWN;         // warp position: rows 32*wm, cols 64*wn
    const bool vec_ok = ((lda & 1) == 0) && ((ldb & 1) == 0) && ((((size_t)A) & 15) == 0) && ((((size_t)Bm) & 15) == 0);
    if (beta != 0.0) {   // pull the C tile towards L2 while the K loop runs
        for (int e = tid; e < GT_M * (TN / 16); e += NTHREADS) {
            const long long r = m0 + e / (TN / 16), c = n0 + (e % (TN / 16)) * 16;
            if (r < M && c < N) asm volatile("prefetch.global.L2 [%0];" ::"l"(&C[r * ldc + c]));
        }
    }

    double acc[4][8][2];
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 8; ++j) acc[i][j][0] = acc[i][j][1] = 0.0;

    auto stage = [&](long long k0, int buf) {
        double* Ad = As + buf * GT_M * GT_LD;
        double* Bd = Bs + buf * TN * GT_LD;
        // 128 (A) / TN (B) rows x 32 k each: 16 double2 per row
        for (int e = tid; e < GT_M * (GT_K / 2); e += NTHREADS) {
            int r = e / (GT_K / 2), c2 = (e - r * (GT_K / 2)) * 2;
            long long gr = m0 + r, gk = k0 + c2;
            double* d = &Ad[r * GT_LD + c2];
            if (gr < M && gk + 1 < K && vec_ok) cpd16(d, &A[gr * lda + gk]);
            else {
                d[0] = (gr < M && gk < K) ? A[gr * lda + gk] : 0.0;
                d[1] = (gr < M && gk + 1 < K) ? A[gr * lda + gk + 1] : 0.0;
            }
        }
        for (int e = tid; e < TN * (GT_K / 2); e += NTHREADS) {
            int r = e / (GT_K / 2), c2 = (e - r * (GT_K / 2)) * 2;
            long long gr = n0 + r, gk = k0 + c2;
            double* d = &Bd[r * GT_LD + c2];
            if (gr < N && gk + 1 < K && vec_ok) cpd16(d, &Bm[gr * ldb + gk]);
            else {
                d[0] = (gr < N && gk < K) ? Bm[gr * ldb + gk] : 0.0;
                d[1] = (gr < N && gk + 1 < K) ? Bm[gr * ldb + gk + 1] : 0.0;
            }
        }
    };
    const long long nk = (K + GT_K - 1) / GT_K;
    stage(0, 0);
    cpd_commit();
    for (long long kt = 0; kt < nk; ++kt) {
        const int buf = (int)(kt & 1);
        cpd_wait0();
        __syncthreads();
        if (kt + 1 < nk) stage((kt + 1) * GT_K, buf ^ 1);
        cpd_commit();
        const double* Ad = As + buf * GT_M * GT_LD + (32 * wm) * GT_LD;
        const double* Bd = Bs + buf * TN * GT_LD + (64 * wn) * GT_LD;
#pragma unroll
        for (int ks = 0; ks < GT_K / 4; ++ks) {
            double af[4], bf[8];
#pragma unroll
            for (int i = 0; i < 4; ++i) af[i] = Ad[(8 * i + g) * GT_LD + 4 * ks + t];      // A[row][k]
#pragma unroll
            for (int j = 0; j < 8; ++j) bf[j] = Bd[(8 * j + g) * GT_LD + 4 * ks + t];      // B^T[k][col] = B[col][k]
#pragma unroll
            for (int i = 0; i < 4; ++i)
#pragma unroll
                for (int j = 0; j < 8; ++j) dmma884d(acc[i][j][0], acc[i][j][1], af[i], bf[j]);
        }
    }
    cpd_wait0();
    // epilogue: all of the thread's C values are fetched as independent 16-byte loads before the first store, so the
    // read-modify-write costs one memory round trip instead of one per element (the tile was L2-prefetched at entry)
    const bool c_vec = ((ldc & 1) == 0) && ((((size_t)C) & 15) == 0);
#pragma unroll
    for (int ih = 0; ih < 4; ih += 2) {
        double2 cv[2][8];
        if (beta != 0.0) {
#pragma unroll
            for (int i = 0; i < 2; ++i) {
                const long long r = m0 + 32 * wm + 8 * (ih + i) + g;
#pragma unroll
                for (int j = 0; j < 8; ++j) {
                    const long long c = n0 + 64 * wn + 8 * j + 2 * t;
                    cv[i][j] = make_double2(0.0, 0.0);
                    if (r < M) {
                        if (c_vec && c + 1 < N) cv[i][j] = *reinterpret_cast<const double2*>(&C[r * ldc + c]);
                        else {
                            if (c < N) cv[i][j].x = C[r * ldc + c];
                            if (c + 1 < N) cv[i][j].y = C[r * ldc + c + 1];
                        }
                    }
                }
            }
        }
#pragma unroll
        for (int i = 0; i < 2; ++i) {
            const long long r = m0 + 32 * wm + 8 * (ih + i) + g;
            if (r >= M) continue;
#pragma unroll
            for (int j = 0; j < 8; ++j) {
                const long long c = n0 + 64 * wn + 8 * j + 2 * t;
                double2 v = make_double2(alpha * acc[ih + i][j][0], alpha * acc[ih + i][j][1]);
                if (beta != 0.0) {
                    v.x = fma(beta, cv[i][j].x, v.x);
                    v.y = fma(beta, cv[i][j].y, v.y);
                }
                if (c_vec && c + 1 < N) *reinterpret_cast<double2*>(&C[r * ldc + c]) = v;
                else {
                    if (c < N) C[r * ldc + c] = v.x;
                    if (c + 1 < N) C[r * ldc + c + 1] = v.y;
                }
            }
        }
    }
}
// TMA-fed variant of the 128 x 64 tile kernel (nmgp_gemm_tma.cu): 0 launched, 1 not applicable (alignment / driver)
int nmgp_gemm_nt_tma(const double* A, const double* Bm, double* C, long long M, long long N, long long K, long long lda,
                     long long ldb, long long ldc, double alpha, double beta, int lower_only, cudaStream_t st);
static int g_gemm_cp_async_only = 0;
NMGP_API void nmgp_gemm_concurrent_mode(int on) {
    static const bool keep_tma = [] { const char* e = getenv("NMGP_TMA_CONCURRENT"); return e && e[0] == '1'; }();   // probe
    g_gemm_cp_async_only = (on && !keep_tma) ? 1 : 0;
}
static int g_gemm_narrow = -1;   // 1: 128x64 CTA tiles, two CTAs per SM (epilogue of one overlaps the MMAs of the other)
static int gemm_nt_launch(const double* A, const double* Bm, double* C, long long M, long long N, long long K,
                          long long lda, long long ldb, long long ldc, double alpha, double beta, int mode,
                          cudaStream_t st) {
    // mode bit 0: lower tiles only (SYRK); bit 1: 128-wide tiles (one CTA per row block when N <= 128: C may alias A)
    if (M <= 0 || N <= 0) return 0;
    const int lower_only = mode & 1;
    if (g_gemm_narrow < 0) {
        const char* e = getenv("NMGP_GEMM_TILE");
        g_gemm_narrow = (e && e[0] == 'w') ? 0 : 1;
    }
    const bool narrow = !(mode & 2) && (g_gemm_narrow || N <= 64 || K <= 256);
    if (narrow) {
        // operand tiles by the TMA unit when the descriptors can be built (16-byte aligned rows), else cp.async staging.
        // g_gemm_cp_async_only (nmgp_gemm_concurrent_mode): set by the eigen-block pipeline while it keeps several
        // factorisations in flight on different streams: with four of them in flight at T = 12288 the TMA-fed kernel gave
        // run-to-run different factors (relative 2e-7, profiles/microbench/kron_determinism.py) while the cp.async kernel
        // and every single-factorisation run stayed bit-reproducible; until that is understood concurrent work does not
        // use it.
        if (!g_gemm_cp_async_only) {
            const int rt = nmgp_gemm_nt_tma(A, Bm, C, M, N, K, lda, ldb, ldc, alpha, beta, lower_only, st);
            if (rt <= 0) return rt;
        }
        size_t smem = sizeof(double) * 2 * (GT_M + 64) * GT_LD;
        if (int r = nmgp_opt_in_smem(k_gemm_nt<1>, smem, "nmgp_gemm_nt")) return r;
        dim3 grid((unsigned)((N + 63) / 64), (unsigned)((M + GT_M - 1) / GT_M));
        k_gemm_nt<1><<<NMGP_L(grid), 128, smem, st>>>(A, Bm, C, M, N, K, lda, ldb, ldc, alpha, beta, lower_only);
    } else {
        size_t smem = sizeof(double) * 2 * (GT_M + GT_N) * GT_LD;
        if (int r = nmgp_opt_in_smem(k_gemm_nt<2>, smem, "nmgp_gemm_nt")) return r;
        dim3 grid((unsigned)((N + GT_N - 1) / GT_N), (unsigned)((M + GT_M - 1) / GT_M));
        k_gemm_nt<2><<<NMGP_L(grid), 256, smem, st>>>(A, Bm, C, M, N, K, lda, ldb, ldc, alpha, beta, lower_only);
    }
    return nmgp_launch_status("nmgp_gemm_nt");
}
NMGP_API int nmgp_gemm_nt(const double* A, const double* Bm, double* C, long long M, long long N, long long K,
                          long long lda, long long ldb, long long ldc, double alpha, double beta, cudaStream_t st) {
    NMGP_REQUIRE(M >= 0 && N >= 0 && K >= 0 && lda >= K && ldb >= K && ldc >= N, "nmgp_gemm_nt");
    return gemm_nt_launch(A, Bm, C, M, N, K, lda, ldb, ldc, alpha, beta, 0, st);
}

// ------------------------------------------------------------------------------------------------------------
// Blocked Cholesky.  PB = panel width.
#define PB 128

// Diagonal block: in-place lower Cholesky of the nb x nb block at A (nb <= 128, leading dimension lda) AND its explicit
// inverse Linv (128 x 128 row-major, identity-padded beyond nb), one CTA of 16 warps with the block resident in shared
// memory.  16-column steps: warp 0 factorises the 16 x 16 diagonal sub-block in registers (one row per lane, column
// broadcasts by shuffle) and inverts it; all warps form the sub-panel X = A D^-T from that inverse and apply the rank-16
// trailing update with DMMA 8x8x4 tiles.  The inverse turns the panel solve of the rows below into a plain DMMA GEMM
// (X = A21 Linv^T), which removes the latency-bound substitution kernels from the factorisation's critical path.
#define DI_N 128
#define DI_LD 136        // % 16 == 8: conflict-free 16-byte C-fragment accesses of the trailing tiles
#define DI_XLD 20        // % 8 == 4: conflict-free A/B fragment loads of the sub-panel
#define DI_THREADS 512
#define DI_SMEM (sizeof(double) * (DI_N * DI_LD + DI_N * DI_XLD + 8 * 16 * 17))
__global__ void __launch_bounds__(DI_THREADS)
k_potrf_diag_inv(double* __restrict__ A, long long lda, int nb, double* Linv, int* __restrict__ info, int blockno) {
    extern __shared__ __align__(16) double sm[];
    double* Ls = sm;                          // [128][DI_LD]
    double* Xs = Ls + DI_N * DI_LD;           // [128][DI_XLD] sub-panel (DMMA operands); scratch for the inverse
    double* Di = Xs + DI_N * DI_XLD;          // [8][16][17]   inverses of the diagonal 16 x 16 sub-blocks
    const int tid = threadIdx.x, lane = tid & 31, w = tid >> 5, g = lane >> 2, t = lane & 3;
#pragma unroll 8
    for (int e = tid; e < DI_N * DI_N; e += DI_THREADS) {
        const int a = e >> 7, b = e & 127;
        Ls[a * DI_LD + b] = (a < nb && b < nb) ? A[(long long)a * lda + b] : (a == b ? 1.0 : 0.0);
    }
    __syncthreads();
    for (int bk = 0; bk < 8; ++bk) {
        const int kb = 16 * bk;
        if (w == 0) {
            const int i = lane & 15;
            double r[16], rd[16];
#pragma unroll
            for (int c = 0; c < 16; ++c) r[c] = Ls[(kb + i) * DI_LD + kb + c];
#pragma unroll
            for (int k = 0; k < 16; ++k) {
                const double dkk = __shfl_sync(0xffffffffu, r[k], k);
                if (!(dkk > 0.0) && lane == 0) atomicCAS(info, 0, blockno + kb + k + 1);   // first failing pivot wins
                const double rinv = rsqrt(dkk);
                rd[k] = rinv;
                const double lik = (i == k) ? dkk * rinv : r[k] * rinv;
                r[k] = lik;
#pragma unroll
                for (int j = k + 1; j < 16; ++j) {
                    const double v = __shfl_sync(0xffffffffu, lik, j);
                    r[j] = fma(-lik, v, r[j]);
                }
            }
            if (lane < 16) {
#pragma unroll
                for (int c = 0; c < 16; ++c) Ls[(kb + i) * DI_LD + kb + c] = (c <= i) ? r[c] : 0.0;
            }
            __syncwarp();
            // column i of D^-1 by forward substitution (entries above the diagonal come out as exact zeros)
            double x[16];
#pragma unroll
            for (int a = 0; a < 16; ++a) {
                double s0 = (a == i) ? 1.0 : 0.0, s1 = 0.0;
#pragma unroll
                for (int k = 0; k < a; ++k) {
                    const double lv = Ls[(kb + a) * DI_LD + kb + k];
                    if (k & 1) s1 = fma(-lv, x[k], s1);
                    else s0 = fma(-lv, x[k], s0);
                }
                x[a] = (s0 + s1) * rd[a];
            }
            if (lane < 16) {
#pragma unroll
                for (int a = 0; a < 16; ++a) Di[(kb + a) * 17 + i] = x[a];
            }
        }
        __syncthreads();
        const int nrem = DI_N - kb - 16;
        // sub-panel X = A[:, kb:kb+16] Dinv^T on DMMA: one warp per 8 rows (both 8-column halves), written back in place
        // and to Xs (operand layout of the trailing update)
        for (int ti = w; ti < (nrem >> 3); ti += DI_THREADS / 32) {
            double* arow = &Ls[(kb + 16 + 8 * ti + g) * DI_LD + kb];
            double xa[4][2], xb[4][2];
#pragma unroll
            for (int ks = 0; ks < 4; ++ks) {
                const double av = arow[4 * ks + t];
                xa[ks][0] = xa[ks][1] = xb[ks][0] = xb[ks][1] = 0.0;
                dmma884d(xa[ks][0], xa[ks][1], av, Di[(kb + g) * 17 + 4 * ks + t]);
                dmma884d(xb[ks][0], xb[ks][1], av, Di[(kb + 8 + g) * 17 + 4 * ks + t]);
            }
            const double a0 = (xa[0][0] + xa[1][0]) + (xa[2][0] + xa[3][0]), a1 = (xa[0][1] + xa[1][1]) + (xa[2][1] + xa[3][1]);
            const double b0 = (xb[0][0] + xb[1][0]) + (xb[2][0] + xb[3][0]), b1 = (xb[0][1] + xb[1][1]) + (xb[2][1] + xb[3][1]);
            __syncwarp();
            arow[2 * t] = a0; arow[2 * t + 1] = a1; arow[8 + 2 * t] = b0; arow[8 + 2 * t + 1] = b1;
            double* xrow = &Xs[(8 * ti + g) * DI_XLD];
            xrow[2 * t] = a0; xrow[2 * t + 1] = a1; xrow[8 + 2 * t] = b0; xrow[8 + 2 * t + 1] = b1;
        }
        __syncthreads();
        // trailing update with DMMA: lower 8 x 8 tiles (ti >= tj) of the nrem x nrem block, one accumulator per k-step
        const int nt = nrem >> 3, ntiles = nt * (nt + 1) / 2;
        for (int tile = w; tile < ntiles; tile += DI_THREADS / 32) {
            int ti = (int)((sqrtf(8.f * (float)tile + 1.f) - 1.f) * 0.5f);
            while ((ti + 1) * (ti + 2) / 2 <= tile) ++ti;
            while (ti * (ti + 1) / 2 > tile) --ti;
            const int tj = tile - ti * (ti + 1) / 2;
            double c[4][2];
#pragma unroll
            for (int ks = 0; ks < 4; ++ks) {
                c[ks][0] = c[ks][1] = 0.0;
                dmma884d(c[ks][0], c[ks][1], Xs[(8 * ti + g) * DI_XLD + 4 * ks + t], Xs[(8 * tj + g) * DI_XLD + 4 * ks + t]);
            }
            double2* p = reinterpret_cast<double2*>(&Ls[(kb + 16 + 8 * ti + g) * DI_LD + kb + 16 + 8 * tj + 2 * t]);
            double2 v = *p;
            v.x -= (c[0][0] + c[1][0]) + (c[2][0] + c[3][0]);
            v.y -= (c[0][1] + c[1][1]) + (c[2][1] + c[3][1]);
            *p = v;
        }
        __syncthreads();
    }
#pragma unroll 8
    for (int e = tid; e < DI_N * DI_N; e += DI_THREADS) {
        const int a = e >> 7, b = e & 127;
        if (a < nb && b < nb) A[(long long)a * lda + b] = (b <= a) ? Ls[a * DI_LD + b] : 0.0;
        // inverse: diagonal 16 x 16 sub-blocks from Di, zeros above them
        if ((b >> 4) >= (a >> 4)) Linv[e] = ((b >> 4) == (a >> 4)) ? Di[a * 17 + (b & 15)] : 0.0;
    }
    __syncthreads();
    // sub-blocks below the diagonal:  Linv[bi][bj] = -Dinv[bi] * sum_{k=bj}^{bi-1} L[bi][k] Linv[k][bj]  as DMMA block
    // products.  The two 8-column halves of a block column bj are independent, so each of the 16 warps owns one
    // (bj, half) task and walks down bi with warp-level synchronisation only; the task table balances the DMMA count
    // per scheduler (448 each).  Finished blocks are kept in the unused upper block (bj, bi) of Ls.
    {
        const int sp = w & 3, slot = w >> 2, tj = sp & 1;
        const int bj = (((sp >> 1) == 0 ? 0x6530 : 0x7421) >> (4 * slot)) & 15;
        double* Tw = Xs + w * (16 * 9);       // warp-private 16 x 8 scratch
        for (int bi = bj + 1; bi < 8; ++bi) {
            const double* ar0 = &Ls[(16 * bi + g) * DI_LD];
            const double* ar1 = ar0 + 8 * DI_LD;
            double c[2][4][2];
#pragma unroll
            for (int ks = 0; ks < 4; ++ks) {
                const double bv = Di[(16 * bj + 4 * ks + t) * 17 + 8 * tj + g];
                c[0][ks][0] = c[0][ks][1] = c[1][ks][0] = c[1][ks][1] = 0.0;
                dmma884d(c[0][ks][0], c[0][ks][1], ar0[16 * bj + 4 * ks + t], bv);
                dmma884d(c[1][ks][0], c[1][ks][1], ar1[16 * bj + 4 * ks + t], bv);
            }
            for (int k = bj + 1; k < bi; ++k) {
#pragma unroll
                for (int ks = 0; ks < 4; ++ks) {
                    const double bv = Ls[(16 * bj + 4 * ks + t) * DI_LD + 16 * k + 8 * tj + g];
                    dmma884d(c[0][ks][0], c[0][ks][1], ar0[16 * k + 4 * ks + t], bv);
                    dmma884d(c[1][ks][0], c[1][ks][1], ar1[16 * k + 4 * ks + t], bv);
                }
            }
#pragma unroll
            for (int ti = 0; ti < 2; ++ti) {
                Tw[(8 * ti + g) * 9 + 2 * t] = (c[ti][0][0] + c[ti][1][0]) + (c[ti][2][0] + c[ti][3][0]);
                Tw[(8 * ti + g) * 9 + 2 * t + 1] = (c[ti][0][1] + c[ti][1][1]) + (c[ti][2][1] + c[ti][3][1]);
            }
            __syncwarp();
#pragma unroll
            for (int ks = 0; ks < 4; ++ks) {
                const double bv = Tw[(4 * ks + t) * 9 + g];
                c[0][ks][0] = c[0][ks][1] = c[1][ks][0] = c[1][ks][1] = 0.0;
                dmma884d(c[0][ks][0], c[0][ks][1], Di[(16 * bi + g) * 17 + 4 * ks + t], bv);
                dmma884d(c[1][ks][0], c[1][ks][1], Di[(16 * bi + 8 + g) * 17 + 4 * ks + t], bv);
            }
#pragma unroll
            for (int ti = 0; ti < 2; ++ti) {
                const double v0 = -((c[ti][0][0] + c[ti][1][0]) + (c[ti][2][0] + c[ti][3][0]));
                const double v1 = -((c[ti][0][1] + c[ti][1][1]) + (c[ti][2][1] + c[ti][3][1]));
                const int rr = 8 * ti + g, cc = 8 * tj + 2 * t;
                Ls[(16 * bj + rr) * DI_LD + 16 * bi + cc] = v0;
                Ls[(16 * bj + rr) * DI_LD + 16 * bi + cc + 1] = v1;
                Linv[(16 * bi + rr) * DI_N + 16 * bj + cc] = v0;
                Linv[(16 * bi + rr) * DI_N + 16 * bj + cc + 1] = v1;
            }
            __syncwarp();
        }
    }
    __syncthreads();
}
// Batched variant of the factorisation half of k_potrf_diag_inv for the Q x Q matrices of the DSVI step with
// 64 < Q <= 128 (one CTA of 16 warps per matrix, 16-column steps, DMMA sub-panel and trailing update):
// C = chol(A + jitter I) (strict upper triangle zeroed), hld = sum log diag C, *info = 1 + index of a non-PD matrix.
// Replaces the one-thread-per-row kernel of nmgp_small.cu there (Q = 100: ~130 us of serial dot products per launch).
__global__ void __launch_bounds__(DI_THREADS)
k_potrf_batched_blocked(const double* __restrict__ A, double jitter, double* __restrict__ C, double* __restrict__ hld,
                        int* __restrict__ info, int nb) {
    const double* Ain = A + (size_t)blockIdx.x * nb * nb;
    double* Cout = C + (size_t)blockIdx.x * nb * nb;
    extern __shared__ __align__(16) double sm[];
    double* Ls = sm;                          // [128][DI_LD]
    double* Xs = Ls + DI_N * DI_LD;           // [128][DI_XLD] sub-panel (DMMA operands); scratch for the inverse
    double* Di = Xs + DI_N * DI_XLD;          // [8][16][17]   inverses of the diagonal 16 x 16 sub-blocks
    const int tid = threadIdx.x, lane = tid & 31, w = tid >> 5, g = lane >> 2, t = lane & 3;
#pragma unroll 8
    for (int e = tid; e < DI_N * DI_N; e += DI_THREADS) {
        const int a = e >> 7, b = e & 127;
        Ls[a * DI_LD + b] = (a < nb && b < nb) ? Ain[(long long)a * nb + b] + (a == b ? jitter : 0.0) : (a == b ? 1.0 : 0.0);
    }
    __syncthreads();
    for (int bk = 0; bk < 8; ++bk) {
        const int kb = 16 * bk;
        if (w == 0) {
            const int i = lane & 15;
            double r[16], rd[16];
#pragma unroll
            for (int c = 0; c < 16; ++c) r[c] = Ls[(kb + i) * DI_LD + kb + c];
#pragma unroll
            for (int k = 0; k < 16; ++k) {
                const double dkk = __shfl_sync(0xffffffffu, r[k], k);
                if (!(dkk > 0.0) && lane == 0 && kb + k < nb) atomicMax(info, (int)blockIdx.x + 1);
                const double rinv = rsqrt(dkk);
                rd[k] = rinv;
                const double lik = (i == k) ? dkk * rinv : r[k] * rinv;
                r[k] = lik;
#pragma unroll
                for (int j = k + 1; j < 16; ++j) {
                    const double v = __shfl_sync(0xffffffffu, lik, j);
                    r[j] = fma(-lik, v, r[j]);
                }
            }
            if (lane < 16) {
#pragma unroll
                for (int c = 0; c < 16; ++c) Ls[(kb + i) * DI_LD + kb + c] = (c <= i) ? r[c] : 0.0;
            }
            __syncwarp();
            // column i of D^-1 by forward substitution (entries above the diagonal come out as exact zeros)
            double x[16];
#pragma unroll
            for (int a = 0; a < 16; ++a) {
                double s0 = (a == i) ? 1.0 : 0.0, s1 = 0.0;
#pragma unroll
                for (int k = 0; k < a; ++k) {
                    const double lv = Ls[(kb + a) * DI_LD + kb + k];
                    if (k & 1) s1 = fma(-lv, x[k], s1);
                    else s0 = fma(-lv, x[k], s0);
                }
                x[a] = (s0 + s1) * rd[a];
            }
            if (lane < 16) {
#pragma unroll
                for (int a = 0; a < 16; ++a) Di[(kb + a) * 17 + i] = x[a];
            }
        }
        __syncthreads();
        const int nrem = DI_N - kb - 16;
        // sub-panel X = A[:, kb:kb+16] Dinv^T on DMMA: one warp per 8 rows (both 8-column halves), written back in place
        // and to Xs (operand layout of the trailing update)
        for (int ti = w; ti < (nrem >> 3); ti += DI_THREADS / 32) {
            double* arow = &Ls[(kb + 16 + 8 * ti + g) * DI_LD + kb];
            double xa[4][2], xb[4][2];
#pragma unroll
            for (int ks = 0; ks < 4; ++ks) {
                const double av = arow[4 * ks + t];
                xa[ks][0] = xa[ks][1] = xb[ks][0] = xb[ks][1] = 0.0;
                dmma884d(xa[ks][0], xa[ks][1], av, Di[(kb + g) * 17 + 4 * ks + t]);
                dmma884d(xb[ks][0], xb[ks][1], av, Di[(kb + 8 + g) * 17 + 4 * ks + t]);
            }
            const double a0 = (xa[0][0] + xa[1][0]) + (xa[2][0] + xa[3][0]), a1 = (xa[0][1] + xa[1][1]) + (xa[2][1] + xa[3][1]);
            const double b0 = (xb[0][0] + xb[1][0]) + (xb[2][0] + xb[3][0]), b1 = (xb[0][1] + xb[1][1]) + (xb[2][1] + xb[3][1]);
            __syncwarp();
            arow[2 * t] = a0; arow[2 * t + 1] = a1; arow[8 + 2 * t] = b0; arow[8 + 2 * t + 1] = b1;
            double* xrow = &Xs[(8 * ti + g) * DI_XLD];
            xrow[2 * t] = a0; xrow[2 * t + 1] = a1; xrow[8 + 2 * t] = b0; xrow[8 + 2 * t + 1] = b1;
        }
        __syncthreads();
        // trailing update with DMMA: lower 8 x 8 tiles (ti >= tj) of the nrem x nrem block, one accumulator per k-step
        const int nt = nrem >> 3, ntiles = nt * (nt + 1) / 2;
        for (int tile = w; tile < ntiles; tile += DI_THREADS / 32) {
            int ti = (int)((sqrtf(8.f * (float)tile + 1.f) - 1.f) * 0.5f);
            while ((ti + 1) * (ti + 2) / 2 <= tile) ++ti;
            while (ti * (ti + 1) / 2 > tile) --ti;
            const int tj = tile - ti * (ti + 1) / 2;
            double c[4][2];
#pragma unroll
            for (int ks = 0; ks < 4; ++ks) {
                c[ks][0] = c[ks][1] = 0.0;
                dmma884d(c[ks][0], c[ks][1], Xs[(8 * ti + g) * DI_XLD + 4 * ks + t], Xs[(8 * tj + g) * DI_XLD + 4 * ks + t]);
            }
            double2* p = reinterpret_cast<double2*>(&Ls[(kb + 16 + 8 * ti + g) * DI_LD + kb + 16 + 8 * tj + 2 * t]);
            double2 v = *p;
            v.x -= (c[0][0] + c[1][0]) + (c[2][0] + c[3][0]);
            v.y -= (c[0][1] + c[1][1]) + (c[2][1] + c[3][1]);
            *p = v;
        }
        __syncthreads();
    }
    for (int e = tid; e < nb * nb; e += DI_THREADS) {
        const int a = e / nb, b = e - a * nb;
        Cout[e] = (b <= a) ? Ls[a * DI_LD + b] : 0.0;
    }
    double lg = (tid < nb) ? log(Ls[tid * DI_LD + tid]) : 0.0;
    lg = block_sum(lg);
    if (tid == 0) hld[blockIdx.x] = lg;
}
int nmgp_potrf_batched_blocked(const double* A, double jitter, double* C, double* hld, int* info, int nb_mat, int Q,
                               cudaStream_t st) {
    if (Q <= 64 || Q > DI_N) return 1;
    if (int r = nmgp_opt_in_smem(k_potrf_batched_blocked, DI_SMEM, "nmgp_potrf_batched(blocked)")) return r;
    k_potrf_batched_blocked<<<NMGP_L(nb_mat), DI_THREADS, DI_SMEM, st>>>(A, jitter, C, hld, info, Q);
    return nmgp_launch_status("nmgp_potrf_batched(blocked)");
}
__global__ void k_zero_upper(double* __restrict__ A, long long T, long long lda) {
    long long c = (long long)blockIdx.x * blockDim.x + threadIdx.x, r = blockIdx.y;
    if (c < T && c > r) A[r * lda + c] = 0.0;
}
__global__ void k_logdiag_sum(const double* __restrict__ A, long long T, long long lda, double* __restrict__ out) {
    double s = 0.0;
    for (long long i = threadIdx.x; i < T; i += blockDim.x) s += log(A[i * lda + i]);
    s = block_sum(s);
    if (threadIdx.x == 0) out[0] = s;
}
// factor one panel of width nb with `rest` rows below it, left-looking over 128-wide column blocks:
//   block column c0 (rows c0.. of the panel and all rows below) -= X[:, :c0] L[c0:c0+h, :c0]^T      (DMMA GEMM)
//   diagonal block factorised and inverted by k_potrf_diag_inv
//   rows below it:  X = A Linv^T                                                                    (DMMA GEMM, in place)
// linv: 128 x 128 scratch owned by the stream the panel runs on.
static int factor_panel(double* Akk, long long lda, int nb, long long rest, double* linv, int* info, int pivot_base,
                        cudaStream_t st) {
    for (int c0 = 0; c0 < nb; c0 += DI_N) {
        const int h = nb - c0 > DI_N ? DI_N : nb - c0;
        double* Dcc = Akk + (long long)c0 * lda + c0;
        const long long rows = (nb - c0) + rest;            // rows from the diagonal block down
        if (c0 > 0)
            if (int r = gemm_nt_launch(Akk + (long long)c0 * lda, Akk + (long long)c0 * lda, Dcc, rows, h, c0, lda, lda, lda,
                                       -1.0, 1.0, 0, st))
                return r;
        k_potrf_diag_inv<<<NMGP_L(1), DI_THREADS, DI_SMEM, st>>>(Dcc, lda, h, linv, info, pivot_base + c0);
        const long long below = rows - h;
        if (below > 0) {
            // in place: the 128-wide tile variant gives every row block to one CTA, which reads all of its K columns
            // before its epilogue writes them
            double* A21 = Dcc + (long long)h * lda;
            if (int r = gemm_nt_launch(A21, linv, A21, below, h, h, lda, DI_N, lda, 1.0, 0.0, 2, st)) return r;
        }
    }
    return 0;
}
// top level with look-ahead: the next panel is factorised on a helper stream while the main stream applies the bulk
// of the trailing update, so the latency-bound panel work hides behind the DMMA SYRK.
// Library-owned scratch of the blocked Cholesky, one set per (device, slot): a highest-priority helper stream, two
// events and two 128 x 128 inverse blocks (main stream / helper stream).  Independent factorisations that run
// concurrently (the eigen-blocks of a Kronecker log-density, dealt round-robin to a few streams so that the panel phase
// of one overlaps the trailing updates of the others) must use different slots.
#define POTRF_SLOTS 8
#define POTRF_MAXDEV 16
struct PotrfCtx {
    cudaStream_t helper = nullptr;
    cudaEvent_t ev_upd = nullptr, ev_pan = nullptr;
    double* linv = nullptr;
};
static PotrfCtx g_potrf_ctx[POTRF_MAXDEV][POTRF_SLOTS];
static PotrfCtx* potrf_ctx(int slot) {
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= POTRF_MAXDEV || slot < 0 || slot >= POTRF_SLOTS) {
        nmgp_set_error("nmgp_potrf_big: bad device / slot (%d, %d)", dev, slot);
        return nullptr;
    }
    PotrfCtx* c = &g_potrf_ctx[dev][slot];
    if (!c->linv) {
        // highest priority: the panel's CTAs must get SM slots ahead of the queued tiles of the trailing update
        int prio_lo = 0, prio_hi = 0;
        cudaDeviceGetStreamPriorityRange(&prio_lo, &prio_hi);
        if (cudaMalloc(&c->linv, sizeof(double) * 2 * DI_N * DI_N) != cudaSuccess ||
            cudaStreamCreateWithPriority(&c->helper, cudaStreamNonBlocking, prio_hi) != cudaSuccess ||
            cudaEventCreateWithFlags(&c->ev_upd, cudaEventDisableTiming) != cudaSuccess ||
            cudaEventCreateWithFlags(&c->ev_pan, cudaEventDisableTiming) != cudaSuccess) {
            c->linv = nullptr;
            nmgp_set_error("nmgp_potrf_big: cannot create the look-ahead stream / 256 KB panel scratch");
            return nullptr;
        }
    }
    return c;
}
static int potrf_lookahead(PotrfCtx* ctx, double* A, long long T, long long lda, int pb, int* info, cudaStream_t s0) {
    cudaStream_t s1 = ctx->helper;
    cudaEvent_t g_ev_upd = ctx->ev_upd, g_ev_pan = ctx->ev_pan;
    double* g_linv = ctx->linv;
    {   // panel 0 on the main stream
        const int nb = (int)min((long long)pb, T);
        if (int r = factor_panel(A, lda, nb, T - nb, g_linv, info, 0, s0)) return r;
    }
    for (long long k = 0; k < T; k += pb) {
        const int nb = (int)min((long long)pb, T - k);
        const long long rest = T - k - nb;
        if (rest <= 0) break;
        double* A21 = A + (k + nb) * lda + k;            // rest x nb   (factorised panel k, rows below the diagonal block)
        double* A22 = A + (k + nb) * lda + (k + nb);     // rest x rest (trailing matrix)
        const int nb2 = (int)min((long long)pb, rest);   // width of the next panel
        // 1. bring the next panel's columns up to date (main stream)
        if (int r = gemm_nt_launch(A21, A21, A22, rest, nb2, nb, lda, lda, lda, -1.0, 1.0, 0, s0)) return r;
        cudaEventRecord(g_ev_upd, s0);
        // 2. factorise the next panel on the helper stream
        cudaStreamWaitEvent(s1, g_ev_upd, 0);
        if (int r = factor_panel(A22, lda, nb2, rest - nb2, g_linv + DI_N * DI_N, info, (int)(k + nb), s1)) return r;
        cudaEventRecord(g_ev_pan, s1);
        // 3. rest of the trailing update (columns beyond the next panel, lower tiles only) on the main stream
        const long long rest2 = rest - nb2;
        if (rest2 > 0) {
            double* B21 = A21 + (long long)nb2 * lda;
            if (int r = gemm_nt_launch(B21, B21, A22 + (long long)nb2 * lda + nb2, rest2, rest2, nb, lda, lda, lda, -1.0, 1.0,
                                       1, s0))
                return r;
        }
        // 4. the next iteration (and anything after us on the main stream) needs the factorised panel
        cudaStreamWaitEvent(s0, g_ev_pan, 0);
    }
    return 0;
}
// In-place lower Cholesky of A (T x T, leading dimension lda); strict upper triangle zeroed; hld = sum log diag(L);
// *info = 1 + index of the first non-positive pivot (0 if none).  Reference sites: torch.logdet / torch.inverse
// at distributions.py:109-110 and logpos.py:352-353 (dense path), and the per-eigen-block factorisations of the
// Kronecker path.
NMGP_API int nmgp_potrf_big_slot(double* A, long long T, long long lda, double* hld, int* info, int slot, int panel,
                                 cudaStream_t st) {
    NMGP_REQUIRE(T > 0 && lda >= T && T < 2147483647LL, "nmgp_potrf_big");
    if (int r = nmgp_opt_in_smem(k_potrf_diag_inv, DI_SMEM, "nmgp_potrf_big")) return r;
    PotrfCtx* ctx = potrf_ctx(slot);
    if (!ctx) return -4;
    if (T > 1024) {
        int pb = T >= 12288 ? 512 : (T >= 6144 ? 256 : PB);   // measured best on B200 (profiles/README.md)
        if (panel >= 128) pb = (panel / 128) * 128;            // caller's choice (concurrent blocks prefer wider panels)
        if (const char* e = getenv("NMGP_POTRF_PB")) pb = atoi(e) >= 128 ? (atoi(e) / 128) * 128 : pb;   // tuning knob
        if (int r = potrf_lookahead(ctx, A, T, lda, pb, info, st)) return r;
    } else {
        if (int r = factor_panel(A, lda, (int)T, 0, ctx->linv, info, 0, st)) return r;
    }
    dim3 gz((unsigned)((T + 255) / 256), (unsigned)min(T, 65535LL));
    if (T <= 65535) k_zero_upper<<<NMGP_L(gz), 256, 0, st>>>(A, T, lda);
    if (hld) k_logdiag_sum<<<NMGP_L(1), 1024, 0, st>>>(A, T, lda, hld);
    return nmgp_launch_status("nmgp_potrf_big");
}
NMGP_API int nmgp_potrf_big(double* A, long long T, long long lda, double* hld, int* info, cudaStream_t st) {
    return nmgp_potrf_big_slot(A, T, lda, hld, info, 0, 0, st);
}

// Inverse of the lower-triangular nb x nb block at L (nb <= 128) into out (row-major, leading dimension ldo; entries
// above the diagonal are written as zeros): one CTA, thread c solves L x = e_c by forward substitution from a shared-
// memory copy.  Building block of the blocked triangular inverse used by the Cholesky-based log-density adjoint.
__global__ void __launch_bounds__(PB)
k_tri_inv_block(const double* __restrict__ L, long long lda, int nb, double* out, long long ldo, double scale) {
    extern __shared__ double sm[];
    constexpr int LD = PB + 1;
    double* Ls = sm;                        // [PB][PB+1]; the inverse is built in `out` itself (column c by thread c:
                                            // coalesced across the threads, each thread only re-reads its own column)
    const int tid = threadIdx.x;
    for (int e = tid; e < PB * PB; e += blockDim.x) {
        const int a = e / PB, b = e - a * PB;
        Ls[a * LD + b] = (a < nb && b < nb && b <= a) ? L[(long long)a * lda + b] : (a == b ? 1.0 : 0.0);
    }
    __syncthreads();
    const int c = tid;
    if (c < nb) {
        for (int a = 0; a < nb; ++a) {
            double v = 0.0;
            if (a >= c) {
                double s0 = (a == c) ? 1.0 : 0.0, s1 = 0.0;
                int k = c;
                for (; k + 1 < a; k += 2) {
                    s0 = fma(-Ls[a * LD + k], out[(long long)k * ldo + c], s0);
                    s1 = fma(-Ls[a * LD + k + 1], out[(long long)(k + 1) * ldo + c], s1);
                }
                if (k < a) s0 = fma(-Ls[a * LD + k], out[(long long)k * ldo + c], s0);
                v = (s0 + s1) / Ls[a * LD + a];
            }
            out[(long long)a * ldo + c] = v;
        }
        if (scale != 1.0)
            for (int a = c; a < nb; ++a) out[(long long)a * ldo + c] *= scale;
    }
}
NMGP_API int nmgp_tri_inv_block(const double* L, long long lda, int nb, double* out, long long ldo, double scale,
                                cudaStream_t st) {
    NMGP_REQUIRE(nb > 0 && nb <= PB && lda >= nb && ldo >= nb, "nmgp_tri_inv_block");
    const size_t smem = sizeof(double) * PB * (PB + 1);
    if (int r = nmgp_opt_in_smem(k_tri_inv_block, smem, "nmgp_tri_inv_block")) return r;
    k_tri_inv_block<<<NMGP_L(1), PB, smem, st>>>(L, lda, nb, out, ldo, scale);
    return nmgp_launch_status("nmgp_tri_inv_block");
}

// ------------------------------------------------------------------------------------------------------------
// x <- (L L^T)^-1 x for one vector: blocked forward then backward substitution (PB-wide diagonal solves in one CTA,
// the remaining update as a memory-bound matrix-vector product).
// One CTA (4 warps): 32-wide chunks are solved by warp 0 with the running right-hand side in registers (one row per
// lane, solved entries broadcast by shuffle: no block-wide barrier per pivot), the other rows of the block are then
// updated by all threads.  Indices >= nb are padded with an identity so every chunk is full.
__global__ void __launch_bounds__(PB)
k_trsv_diag(const double* __restrict__ L, long long lda, double* __restrict__ x, int nb, int transposed) {
    extern __shared__ double sm[];
    constexpr int LD = PB + 1;
    double* Ls = sm;            // [PB][PB+1]
    double* xs = Ls + PB * LD;  // [PB]
    const int tid = threadIdx.x, lane = tid & 31, w = tid >> 5;
    for (int e = tid; e < PB * PB; e += blockDim.x) {
        const int a = e / PB, b = e - a * PB;
        Ls[a * LD + b] = (a < nb && b < nb) ? L[(long long)a * lda + b] : (a == b ? 1.0 : 0.0);
    }
    xs[tid] = tid < nb ? x[tid] : 0.0;
    __syncthreads();
    const int nchunk = (nb + 31) / 32;
    for (int ci = 0; ci < nchunk; ++ci) {
        const int c0 = 32 * (transposed ? nchunk - 1 - ci : ci);
        if (w == 0) {
            double r = xs[c0 + lane];
            const double rinv = 1.0 / Ls[(c0 + lane) * LD + c0 + lane];
            if (!transposed) {
#pragma unroll 8
                for (int c = 0; c < 32; ++c) {
                    const double xv = __shfl_sync(0xffffffffu, r * rinv, c);
                    if (lane == c) r = xv;
                    else if (lane > c) r = fma(-Ls[(c0 + lane) * LD + c0 + c], xv, r);
                }
            } else {
#pragma unroll 8
                for (int c = 31; c >= 0; --c) {
                    const double xv = __shfl_sync(0xffffffffu, r * rinv, c);
                    if (lane == c) r = xv;
                    else if (lane < c) r = fma(-Ls[(c0 + c) * LD + c0 + lane], xv, r);
                }
            }
            xs[c0 + lane] = r;
        }
        __syncthreads();
        // rows of the block still to be solved: below the chunk (forward) / above it (transposed)
        const bool mine = transposed ? (tid < c0) : (tid >= c0 + 32);
        if (mine) {
            double acc = xs[tid];
            if (!transposed) {
#pragma unroll 8
                for (int c = 0; c < 32; ++c) acc = fma(-Ls[tid * LD + c0 + c], xs[c0 + c], acc);
            } else {
#pragma unroll 8
                for (int c = 0; c < 32; ++c) acc = fma(-Ls[(c0 + c) * LD + tid], xs[c0 + c], acc);
            }
            xs[tid] = acc;
        }
        __syncthreads();
    }
    if (tid < nb) x[tid] = xs[tid];
}
// y[r] -= sum_c M[r,c] x[c] (not transposed: M is rows x nb) or y[c] -= sum_r M[r,c] x[r] (transposed: M is nb.. rows)
__global__ void k_gemv_sub(const double* __restrict__ Mx, long long lda, const double* __restrict__ x,
                           double* __restrict__ y, long long rows, int nb) {
    // one warp per row of M: y[row] -= M[row, 0:nb] . x
    const int lane = threadIdx.x & 31;
    long long row = (long long)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (row >= rows) return;
    double acc = 0.0;
    for (int c = lane; c < nb; c += 32) acc = fma(Mx[row * lda + c], x[c], acc);
    acc = warp_sum(acc);
    if (lane == 0) y[row] -= acc;
}
__global__ void k_gemv_t_sub(const double* __restrict__ Mx, long long lda, const double* __restrict__ x,
                             double* __restrict__ y, int rows, long long ncols) {
    // y[c] -= sum_{r < rows} M[r, c] x[r]   (one thread per column, coalesced across columns)
    const long long c = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= ncols) return;
    double acc = 0.0;
    for (int r = 0; r < rows; ++r) acc = fma(Mx[r * lda + c], x[r], acc);
    y[c] -= acc;
}
NMGP_API int nmgp_potrs_vec(const double* L, long long T, long long lda, double* x, cudaStream_t st) {
    NMGP_REQUIRE(T > 0 && lda >= T, "nmgp_potrs_vec");
    const size_t smem = sizeof(double) * (PB * (PB + 1) + PB);
    if (int r = nmgp_opt_in_smem(k_trsv_diag, smem, "nmgp_potrs_vec")) return r;
    for (long long k = 0; k < T; k += PB) {                       // L y = b
        const int nb = (int)min((long long)PB, T - k);
        k_trsv_diag<<<NMGP_L(1), PB, smem, st>>>(L + k * lda + k, lda, x + k, nb, 0);
        const long long rest = T - k - nb;
        if (rest > 0)
            k_gemv_sub<<<NMGP_L((unsigned)((rest + 7) / 8)), 256, 0, st>>>(L + (k + nb) * lda + k, lda, x + k, x + k + nb, rest, nb);
    }
    for (long long k = ((T - 1) / PB) * PB; k >= 0; k -= PB) {    // L^T x = y
        const int nb = (int)min((long long)PB, T - k);
        k_trsv_diag<<<NMGP_L(1), PB, smem, st>>>(L + k * lda + k, lda, x + k, nb, 1);
        if (k > 0)   // x[0:k] -= L[k:k+nb, 0:k]^T x[k:k+nb]
            k_gemv_t_sub<<<NMGP_L((unsigned)((k + 127) / 128)), 128, 0, st>>>(L + k * lda, lda, x + k, x, nb, k);
    }
    return nmgp_launch_status("nmgp_potrs_vec");
}

// ------------------------------------------------------------------------------------------------------------
// A = alpha K + sigma2 I
__global__ void k_scale_add_diag(const double* __restrict__ K, double* __restrict__ A, long long T, double alpha,
                                 double sigma2) {
    long long c = (long long)blockIdx.x * blockDim.x + threadIdx.x, r = blockIdx.y + (long long)blockIdx.z * 65535;
    if (c < T && r < T) A[r * T + c] = fma(alpha, K[r * T + c], (r == c) ? sigma2 : 0.0);
}
// same with alpha = alpha_dev[0] and sigma2 = sigma2_dev[0] read on the device (no host round trip for the eigenvalue)
__global__ void k_scale_add_diag_dev(const double* __restrict__ K, double* __restrict__ A, long long T,
                                     const double* __restrict__ alpha_dev, const double* __restrict__ sigma2_dev) {
    const double alpha = alpha_dev[0], sigma2 = sigma2_dev[0];
    long long c = (long long)blockIdx.x * blockDim.x + threadIdx.x, r = blockIdx.y + (long long)blockIdx.z * 65535;
    if (c < T && r < T) A[r * T + c] = fma(alpha, K[r * T + c], (r == c) ? sigma2 : 0.0);
}
NMGP_API int nmgp_scale_add_diag_dev(const double* K, double* A, long long T, const double* alpha_dev,
                                     const double* sigma2_dev, cudaStream_t st) {
    NMGP_REQUIRE(T > 0, "nmgp_scale_add_diag_dev");
    dim3 grid((unsigned)((T + 255) / 256), (unsigned)min(T, 65535LL), (unsigned)((T + 65534) / 65535));
    k_scale_add_diag_dev<<<NMGP_L(grid), 256, 0, st>>>(K, A, T, alpha_dev, sigma2_dev);
    return nmgp_launch_status("nmgp_scale_add_diag_dev");
}
// Augmented system of one eigen-block: rows 0..T-1 = alpha K + sigma2 I, row T = (r^T, c) with
// c = 1 + |r|^2 / sigma2 >= 1 + r^T A^-1 r (A >= sigma2 I), leading dimension lda >= T + 1.  The Cholesky factor of the
// (T+1) x (T+1) matrix carries (L^-1 r)^T in its last row, so the quadratic form r^T A^-1 r = |L^-1 r|^2 comes out of the
// blocked factorisation itself (tensor-core panel GEMMs) instead of a chain of ~2 T/128 latency-bound substitution
// launches per block.  Only the lower triangle is referenced by the factorisation; the upper part of the last column is
// left unset.
__global__ void k_build_augmented(const double* __restrict__ K, const double* __restrict__ r, double* __restrict__ A,
                                  long long T, long long lda, const double* __restrict__ alpha_dev,
                                  const double* __restrict__ sigma2_dev, const double* __restrict__ rnorm2_dev) {
    const double alpha = alpha_dev[0], sigma2 = sigma2_dev[0];
    long long c = (long long)blockIdx.x * blockDim.x + threadIdx.x, row = blockIdx.y + (long long)blockIdx.z * 65535;
    if (row > T || c > T) return;
    if (row < T) {
        if (c < T) A[row * lda + c] = fma(alpha, K[row * T + c], (row == c) ? sigma2 : 0.0);
    } else {
        A[row * lda + c] = (c < T) ? r[c] : 1.0 + rnorm2_dev[0] / sigma2;
    }
}
NMGP_API int nmgp_build_augmented(const double* K, const double* r, double* A, long long T, long long lda,
                                  const double* alpha_dev, const double* sigma2_dev, const double* rnorm2_dev,
                                  cudaStream_t st) {
    NMGP_REQUIRE(T > 0 && lda >= T + 1, "nmgp_build_augmented");
    dim3 grid((unsigned)((T + 1 + 255) / 256), (unsigned)min(T + 1, 65535LL), (unsigned)((T + 1 + 65534) / 65535));
    k_build_augmented<<<NMGP_L(grid), 256, 0, st>>>(K, r, A, T, lda, alpha_dev, sigma2_dev, rnorm2_dev);
    return nmgp_launch_status("nmgp_build_augmented");
}
// from the factor of the augmented system: quad = |row T, columns 0..T-1|^2, hld = hld_aug - log(A[T,T])
__global__ void k_augmented_results(const double* __restrict__ A, long long T, long long lda,
                                    const double* __restrict__ hld_aug, double* __restrict__ hld, double* __restrict__ quad) {
    double s = 0.0;
    const double* row = A + T * lda;
    for (long long i = threadIdx.x; i < T; i += blockDim.x) s = fma(row[i], row[i], s);
    s = block_sum(s);
    if (threadIdx.x == 0) {
        quad[0] = s;
        hld[0] = hld_aug[0] - log(row[T]);
    }
}
NMGP_API int nmgp_augmented_results(const double* A, long long T, long long lda, const double* hld_aug, double* hld,
                                    double* quad, cudaStream_t st) {
    NMGP_REQUIRE(T > 0 && lda >= T + 1, "nmgp_augmented_results");
    k_augmented_results<<<NMGP_L(1), 1024, 0, st>>>(A, T, lda, hld_aug, hld, quad);
    return nmgp_launch_status("nmgp_augmented_results");
}
NMGP_API int nmgp_scale_add_diag(const double* K, double* A, long long T, double alpha, double sigma2,
                                 cudaStream_t st) {
    NMGP_REQUIRE(T > 0, "nmgp_scale_add_diag");
    dim3 grid((unsigned)((T + 255) / 256), (unsigned)min(T, 65535LL), (unsigned)((T + 65534) / 65535));
    k_scale_add_diag<<<NMGP_L(grid), 256, 0, st>>>(K, A, T, alpha, sigma2);
    return nmgp_launch_status("nmgp_scale_add_diag");
}
// out[(i1*h2+i2), (j1*w2+j2)] = t1[i1,j1] * t2[i2,j2]     (kronecker_operation.py:5-22)
__global__ void k_kron(const double* __restrict__ t1, const double* __restrict__ t2, double* __restrict__ out, int h1,
                       int w1, long long h2, long long w2) {
    long long total = (long long)h1 * h2 * w1 * w2;
    long long gid = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (gid >= total) return;
    long long W = (long long)w1 * w2;
    long long r = gid / W, c = gid - r * W;
    long long i1 = r / h2, i2 = r - i1 * h2, j1 = c / w2, j2 = c - j1 * w2;
    out[gid] = t1[i1 * w1 + j1] * t2[i2 * w2 + j2];
}
NMGP_API int nmgp_kron_product(const double* t1, const double* t2, double* out, int h1, int w1, long long h2,
                               long long w2, cudaStream_t st) {
    long long total = (long long)h1 * h2 * w1 * w2;
    NMGP_REQUIRE(total >= 0 && total < (1LL << 40), "nmgp_kron_product");
    if (total == 0) return 0;
    k_kron<<<NMGP_L((unsigned)((total + 255) / 256)), 256, 0, st>>>(t1, t2, out, h1, w1, h2, w2);
    return nmgp_launch_status("nmgp_kron_product");
}

// ------------------------------------------------------------------------------------------------------------
// Cyclic Jacobi eigen-decomposition of a small symmetric n x n matrix (n <= 128).  Reads the UPPER triangle
// (torch.symeig's default, kronecker_operation.py:45).  Eigenvalues ascending in w, eigenvectors in the columns of V.
// Parallel (round-robin tournament) ordering: every step rotates n/2 disjoint index pairs at once -- all (c, s) from the
// current matrix, then S <- S J (columns) and S <- J^T S (rows) -- so a sweep costs n - 1 steps of three barriers.
// Two kernels: k_eigh_jacobi (one CTA) keeps only S, in shared memory, and LOGS every step's rotations; the
// eigenvector matrix U = J_1 J_2 ... never enters that serial chain (at n = 128 S and U do not fit shared memory
// together, and updating U in global memory was 90 % of the 37 ms the one-kernel version took).  k_eigh_replay then
// applies the logged rotations to the n rows of the identity independently -- one warp per row, the row in shared
// memory, the next step's parameters prefetched -- which is embarrassingly parallel.
#define EJ_THREADS 512
#define EJ_MAXSWEEP 30
struct EJMeta { int nsteps; int pad; };
__host__ __device__ inline long long ej_work_doubles(int n) {
    const long long m = (n + (n & 1)) / 2, steps = (long long)EJ_MAXSWEEP * (n + (n & 1) - 1);
    return 8 + 128 + steps * m * 3;      // meta | order (ints) | cs (double2 per rotation) | pq (int2 per rotation)
}
__global__ void __launch_bounds__(EJ_THREADS)
k_eigh_jacobi(const double* __restrict__ A, double* __restrict__ w, double* __restrict__ work, int n) {
    extern __shared__ double sm[];
    double* S = sm;              // [n][n]
    __shared__ double cs[64][2];
    __shared__ int pq[64][2];
    const int tid = threadIdx.x;
    const int ne = n + (n & 1);          // players of the tournament (a dummy index n when n is odd)
    const int m = ne / 2, N1 = ne - 1;
    EJMeta* meta = reinterpret_cast<EJMeta*>(work);
    int* order = reinterpret_cast<int*>(work + 8);
    double2* lcs = reinterpret_cast<double2*>(work + 8 + 128);
    int2* lpq = reinterpret_cast<int2*>(work + 8 + 128 + (size_t)EJ_MAXSWEEP * N1 * m * 2);
    for (int e = tid; e < n * n; e += blockDim.x) {
        int a = e / n, b = e - a * n;
        S[e] = (b >= a) ? A[a * n + b] : A[b * n + a];
    }
    __syncthreads();
    int step = 0;
    double prev_off = 1e300;
    for (int sweep = 0; sweep < EJ_MAXSWEEP; ++sweep) {
        double off = 0.0;
        for (int e = tid; e < n * n; e += blockDim.x) {
            int a = e / n, b = e - a * n;
            if (a != b) off = fma(S[e], S[e], off);
        }
        off = block_sum(off);
        double diag = 0.0;
        for (int a = tid; a < n; a += blockDim.x) diag = fma(S[a * n + a], S[a * n + a], diag);
        diag = block_sum(diag);
        // converged: quadratic convergence takes the ratio from ~1e-12 to the rounding floor in one sweep; the floor itself
        // (rotated entries are not exact zeros) sits around 1e-32, so a fixed 1e-34 never triggered and all EJ_MAXSWEEP
        // sweeps ran (measured: 24 ms at n = 128 for 10 useful sweeps)
        if (off <= 1e-30 * diag || (sweep >= 3 && off <= 1e-24 * diag && off >= 0.25 * prev_off)) break;
        prev_off = off;
        for (int r = 0; r < N1; ++r, ++step) {
            if (tid < m) {                // pair i of round r (circle method): (r, ne-1), ((r+i) mod N1, (r-i) mod N1)
                int a_ = tid == 0 ? r : (r + tid) % N1, b_ = tid == 0 ? ne - 1 : (r - tid + N1) % N1;
                int p = min(a_, b_), q = max(a_, b_);
                double c = 1.0, sn = 0.0;
                if (q < n) {
                    const double apq = S[p * n + q];
                    if (apq != 0.0) {
                        const double tau = (S[q * n + q] - S[p * n + p]) / (2.0 * apq);
                        const double tt = (tau >= 0.0 ? 1.0 : -1.0) / (fabs(tau) + sqrt(1.0 + tau * tau));
                        c = 1.0 / sqrt(1.0 + tt * tt);
                        sn = tt * c;
                    }
                } else {
                    q = p;                // pair with the dummy: identity
                }
                pq[tid][0] = p; pq[tid][1] = q;
                cs[tid][0] = c; cs[tid][1] = sn;
                lcs[(size_t)step * m + tid] = make_double2(c, sn);
                lpq[(size_t)step * m + tid] = make_int2(p, q);
            }
            __syncthreads();
            for (int i = tid & 63; i < m; i += 64) {                  // columns p, q of S: lane group = pair, rows strided
                const double c = cs[i][0], sn = cs[i][1];
                if (sn != 0.0) {
                    const int p = pq[i][0], q = pq[i][1];
                    for (int k = tid >> 6; k < n; k += EJ_THREADS / 64) {
                        const double skp = S[k * n + p], skq = S[k * n + q];
                        S[k * n + p] = c * skp - sn * skq;
                        S[k * n + q] = sn * skp + c * skq;
                    }
                }
            }
            __syncthreads();
            for (int i = tid >> 7; i < m; i += EJ_THREADS / 128) {    // rows p, q of S: consecutive threads = consecutive columns
                const double c = cs[i][0], sn = cs[i][1];
                if (sn != 0.0) {
                    const int p = pq[i][0], q = pq[i][1];
                    for (int k = tid & 127; k < n; k += 128) {
                        const double spk = S[p * n + k], sqk = S[q * n + k];
                        S[p * n + k] = c * spk - sn * sqk;
                        S[q * n + k] = sn * spk + c * sqk;
                    }
                }
            }
            __syncthreads();
        }
    }
    // sort ascending (n small): rank by counting
    if (tid < n) {
        double v = S[tid * n + tid];
        int rank = 0;
        for (int k = 0; k < n; ++k) {
            double u = S[k * n + k];
            if (u < v || (u == v && k < tid)) ++rank;
        }
        order[rank] = tid;
        w[rank] = v;
    }
    if (tid == 0) meta->nsteps = step;
}
// V[k][rank] = (e_k^T J_1 J_2 ... J_nsteps)[order[rank]]: one warp per row k
#define ER_WARPS 4
__global__ void __launch_bounds__(32 * ER_WARPS)
k_eigh_replay(const double* __restrict__ work, double* __restrict__ V, int n) {
    __shared__ double rows[ER_WARPS][128];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int k = blockIdx.x * ER_WARPS + warp;
    if (k >= n) return;
    const int ne = n + (n & 1), m = ne / 2, N1 = ne - 1;
    const EJMeta* meta = reinterpret_cast<const EJMeta*>(work);
    const int* order = reinterpret_cast<const int*>(work + 8);
    const double2* lcs = reinterpret_cast<const double2*>(work + 8 + 128);
    const int2* lpq = reinterpret_cast<const int2*>(work + 8 + 128 + (size_t)EJ_MAXSWEEP * N1 * m * 2);
    double* row = rows[warp];
    for (int c = lane; c < n; c += 32) row[c] = (c == k) ? 1.0 : 0.0;
    __syncwarp();
    const int nsteps = meta->nsteps;
    // rotations of one step touch disjoint index pairs: lanes take pairs lane, lane + 32 (m <= 64)
    double2 c0 = make_double2(1.0, 0.0), c1 = c0;
    int2 q0 = make_int2(0, 0), q1 = q0;
    if (nsteps > 0) {
        if (lane < m) { c0 = __ldg(&lcs[lane]); q0 = __ldg(&lpq[lane]); }
        if (lane + 32 < m) { c1 = __ldg(&lcs[lane + 32]); q1 = __ldg(&lpq[lane + 32]); }
    }
    for (int st = 0; st < nsteps; ++st) {
        const double2 a0 = c0, a1 = c1;
        const int2 p0 = q0, p1 = q1;
        if (st + 1 < nsteps) {                                        // prefetch the next step's parameters
            const size_t o = (size_t)(st + 1) * m;
            if (lane < m) { c0 = __ldg(&lcs[o + lane]); q0 = __ldg(&lpq[o + lane]); }
            if (lane + 32 < m) { c1 = __ldg(&lcs[o + lane + 32]); q1 = __ldg(&lpq[o + lane + 32]); }
        }
        if (lane < m && a0.y != 0.0) {
            const double up = row[p0.x], uq = row[p0.y];
            row[p0.x] = a0.x * up - a0.y * uq;
            row[p0.y] = a0.y * up + a0.x * uq;
        }
        if (lane + 32 < m && a1.y != 0.0) {
            const double up = row[p1.x], uq = row[p1.y];
            row[p1.x] = a1.x * up - a1.y * uq;
            row[p1.y] = a1.y * up + a1.x * uq;
        }
        __syncwarp();
    }
    for (int c = lane; c < n; c += 32) V[(size_t)k * n + c] = row[order[c]];
}
NMGP_API long long nmgp_eigh_small_work(int n) { return ej_work_doubles(n); }
NMGP_API int nmgp_eigh_small(const double* A, double* w, double* V, double* work /* nmgp_eigh_small_work(n) doubles */,
                             int n, cudaStream_t st) {
    NMGP_REQUIRE(n > 0 && n <= 128 && work != nullptr, "nmgp_eigh_small");
    size_t smem = sizeof(double) * n * n;
    if (int r = nmgp_opt_in_smem(k_eigh_jacobi, smem, "nmgp_eigh_small")) return r;
    k_eigh_jacobi<<<NMGP_L(1), EJ_THREADS, smem, st>>>(A, w, work, n);
    k_eigh_replay<<<NMGP_L((n + ER_WARPS - 1) / ER_WARPS), 32 * ER_WARPS, 0, st>>>(work, V, n);
    return nmgp_launch_status("nmgp_eigh_small");
}

// ------------------------------------------------------------------------------------------------------------
// small vector helpers of the Kronecker log-density (distributions.py:42-51)
// out = (a_scale * a_dev[0]) * x + b * y with the scalar a read on the device
__global__ void k_axpby_dev(const double* __restrict__ x, const double* __restrict__ y, double* __restrict__ out,
                            long long n, const double* __restrict__ a_dev, double a_scale, double b) {
    const double a = a_scale * a_dev[0];
    long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) out[i] = fma(a, x[i], b * y[i]);
}
NMGP_API int nmgp_axpby_dev(const double* x, const double* y, double* out, long long n, const double* a_dev,
                            double a_scale, double b, cudaStream_t st) {
    if (n <= 0) return 0;
    k_axpby_dev<<<NMGP_L((unsigned)((n + 255) / 256)), 256, 0, st>>>(x, y, out, n, a_dev, a_scale, b);
    return nmgp_launch_status("nmgp_axpby_dev");
}
__global__ void k_axpby(const double* __restrict__ x, const double* __restrict__ y, double* __restrict__ out,
                        long long n, double a, double b) {
    long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) out[i] = a * x[i] + b * y[i];
}
NMGP_API int nmgp_axpby(const double* x, const double* y, double* out, long long n, double a, double b,
                        cudaStream_t st) {
    if (n <= 0) return 0;
    k_axpby<<<NMGP_L((unsigned)((n + 255) / 256)), 256, 0, st>>>(x, y, out, n, a, b);
    return nmgp_launch_status("nmgp_axpby");
}
__global__ void k_dot(const double* __restrict__ x, const double* __restrict__ y, double* __restrict__ out,
                      long long n) {
    double s = 0.0;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x)
        s = fma(x[i], y[i], s);
    s = block_sum(s);
    if (threadIdx.x == 0) atomicAdd(out, s);
}
NMGP_API int nmgp_dot(const double* x, const double* y, double* out /* += */, long long n, cudaStream_t st) {
    if (n <= 0) return 0;
    long long blocks = (n + 255) / 256;
    if (blocks > 592) blocks = 592;
    k_dot<<<NMGP_L((unsigned)blocks), 256, 0, st>>>(x, y, out, n);
    return nmgp_launch_status("nmgp_dot");
}
// dist[i,j] = |x_i|^2 + |y_j|^2 - 2 x_i.y_j   (kernels.py:5-21)
__global__ void k_pairwise(const double* __restrict__ X1, const double* __restrict__ X2, double* __restrict__ out,
                           long long T1, long long T2, int dx) {
    long long j = (long long)blockIdx.x * blockDim.x + threadIdx.x, i = blockIdx.y + (long long)blockIdx.z * 65535;
    if (j >= T2 || i >= T1) return;
    double xn = 0.0, yn = 0.0, xy = 0.0;
    for (int k = 0; k < dx; ++k) {
        double a = X1[i * dx + k], b = X2[j * dx + k];
        xn = fma(a, a, xn);
        yn = fma(b, b, yn);
        xy = fma(a, b, xy);
    }
    out[i * T2 + j] = xn + yn - 2.0 * xy;
}
NMGP_API int nmgp_pairwise_dist(const double* X1, const double* X2, double* out, long long T1, long long T2, int dx,
                                cudaStream_t st) {
    NMGP_REQUIRE(T1 >= 0 && T2 >= 0 && dx > 0, "nmgp_pairwise_dist");
    if (T1 == 0 || T2 == 0) return 0;
    dim3 grid((unsigned)((T2 + 255) / 256), (unsigned)min(T1, 65535LL), (unsigned)((T1 + 65534) / 65535));
    k_pairwise<<<NMGP_L(grid), 256, 0, st>>>(X1, X2, out, T1, T2, dx);
    return nmgp_launch_status("nmgp_pairwise_dist");
}
