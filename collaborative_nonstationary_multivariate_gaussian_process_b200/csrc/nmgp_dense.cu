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
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        const long long r = m0 + 32 * wm + 8 * i + g;
        if (r >= M) continue;
#pragma unroll
        for (int j = 0; j < 8; ++j) {
#pragma unroll
            for (int e = 0; e < 2; ++e) {
                const long long c = n0 + 64 * wn + 8 * j + 2 * t + e;
                if (c < N) {
                    double* p = &C[r * ldc + c];
                    double v = alpha * acc[i][j][e];
                    if (beta != 0.0) v = fma(beta, *p, v);
                    *p = v;
                }
            }
        }
    }
}
static int g_gemm_narrow = -1;   // 1: 128x64 CTA tiles, two CTAs per SM (epilogue of one overlaps the MMAs of the other)
static int gemm_nt_launch(const double* A, const double* Bm, double* C, long long M, long long N, long long K,
                          long long lda, long long ldb, long long ldc, double alpha, double beta, int lower_only,
                          cudaStream_t st) {
    if (M <= 0 || N <= 0) return 0;
    if (g_gemm_narrow < 0) {
        const char* e = getenv("NMGP_GEMM_TILE");
        g_gemm_narrow = (e && e[0] == 'w') ? 0 : 1;
    }
    const bool narrow = g_gemm_narrow || N <= 64 || K <= 256;
    if (narrow) {
        size_t smem = sizeof(double) * 2 * (GT_M + 64) * GT_LD;
        if (int r = nmgp_opt_in_smem(k_gemm_nt<1>, smem, "nmgp_gemm_nt")) return r;
        dim3 grid((unsigned)((N + 63) / 64), (unsigned)((M + GT_M - 1) / GT_M));
        k_gemm_nt<1><<<grid, 128, smem, st>>>(A, Bm, C, M, N, K, lda, ldb, ldc, alpha, beta, lower_only);
    } else {
        size_t smem = sizeof(double) * 2 * (GT_M + GT_N) * GT_LD;
        if (int r = nmgp_opt_in_smem(k_gemm_nt<2>, smem, "nmgp_gemm_nt")) return r;
        dim3 grid((unsigned)((N + GT_N - 1) / GT_N), (unsigned)((M + GT_M - 1) / GT_M));
        k_gemm_nt<2><<<grid, 256, smem, st>>>(A, Bm, C, M, N, K, lda, ldb, ldc, alpha, beta, lower_only);
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

// diagonal block: in-place lower Cholesky of the nb x nb block at A (leading dimension lda), one CTA of 32 x 16
// threads, block resident in shared memory, right-looking (every step updates the whole trailing block in parallel)
#define PD_TX 32
#define PD_TY 16
__global__ void __launch_bounds__(PD_TX * PD_TY)
k_potrf_diag(double* __restrict__ A, long long lda, int nb, int* __restrict__ info, int blockno) {
    extern __shared__ double sm[];
    __shared__ double s_rinv;
    const int ld = nb | 1, tid = threadIdx.x, tx = tid & (PD_TX - 1), ty = tid / PD_TX;
    for (int e = tid; e < nb * nb; e += blockDim.x) {
        int a = e / nb, b = e - a * nb;
        sm[a * ld + b] = A[(long long)a * lda + b];
    }
    __syncthreads();
    for (int k = 0; k < nb; ++k) {
        if (tid == 0) {
            double dkk = sm[k * ld + k];
            if (!(dkk > 0.0)) atomicMax(info, blockno + k + 1);
            double piv = sqrt(dkk);
            sm[k * ld + k] = piv;
            s_rinv = 1.0 / piv;
        }
        __syncthreads();
        const double rinv = s_rinv;
        for (int i = k + 1 + tid; i < nb; i += blockDim.x) sm[i * ld + k] *= rinv;
        __syncthreads();
        for (int i = k + 1 + ty; i < nb; i += PD_TY) {
            const double lik = sm[i * ld + k];
            for (int j = k + 1 + tx; j <= i; j += PD_TX) sm[i * ld + j] = fma(-lik, sm[j * ld + k], sm[i * ld + j]);
        }
        __syncthreads();
    }
    for (int e = tid; e < nb * nb; e += blockDim.x) {
        int a = e / nb, b = e - a * nb;
        A[(long long)a * lda + b] = (b <= a) ? sm[a * ld + b] : 0.0;
    }
}
// panel: rows below the diagonal block, X L11^T = A21  ->  one thread per row, forward substitution, L11 in smem
__global__ void __launch_bounds__(128)
k_trsm_panel(const double* __restrict__ L11, double* __restrict__ A21, long long lda, long long nrows, int nb) {
    extern __shared__ double sm[];
    double* Ls = sm;                         // [nb][nb]
    double* tile = Ls + nb * nb;             // [128][nb+1]
    const int ldt = nb + 1, tid = threadIdx.x;
    const long long r0 = (long long)blockIdx.x * 128;
    const int nr = (int)min(128LL, nrows - r0);
    for (int e = tid; e < nb * nb; e += 128) {
        int a = e / nb, b = e - a * nb;
        Ls[e] = L11[(long long)a * lda + b];
    }
    for (int e = tid; e < nr * nb; e += 128) {
        int r = e / nb, a = e - r * nb;
        tile[r * ldt + a] = A21[(r0 + r) * lda + a];
    }
    __syncthreads();
    if (tid < nr) {
        double* y = tile + tid * ldt;
        for (int a = 0; a < nb; ++a) {
            double s0 = y[a], s1 = 0.0;
            int c = 0;
            for (; c + 1 < a; c += 2) {
                s0 = fma(-Ls[a * nb + c], y[c], s0);
                s1 = fma(-Ls[a * nb + c + 1], y[c + 1], s1);
            }
            if (c < a) s0 = fma(-Ls[a * nb + c], y[c], s0);
            y[a] = (s0 + s1) / Ls[a * nb + a];
        }
    }
    __syncthreads();
    for (int e = tid; e < nr * nb; e += 128) {
        int r = e / nb, a = e - r * nb;
        A21[(r0 + r) * lda + a] = tile[r * ldt + a];
    }
}
// same panel solve with the row held in registers (nb padded to 64 with an identity tail): X L^T = A
__global__ void __launch_bounds__(128)
k_trsm_panel_reg64(const double* __restrict__ L11, double* __restrict__ A21, long long lda, long long nrows, int nb) {
    extern __shared__ __align__(16) double sm[];
    constexpr int QP = 64, LDT = QP + 1;
    double* Ls = sm;                 // [64][64], identity padded
    double* rinv = Ls + QP * QP;     // [64]
    double* tile = rinv + QP;        // [128][65]
    const int tid = threadIdx.x;
    const long long r0 = (long long)blockIdx.x * 128;
    const int nr = (int)min(128LL, nrows - r0);
    for (int e = tid; e < QP * QP; e += 128) {
        int a = e / QP, b = e - a * QP;
        Ls[e] = (a < nb && b < nb) ? ((b <= a) ? L11[(long long)a * lda + b] : 0.0) : (a == b ? 1.0 : 0.0);
    }
    for (int a = tid; a < QP; a += 128) rinv[a] = a < nb ? 1.0 / L11[(long long)a * lda + a] : 1.0;
    for (int e = tid; e < 128 * QP; e += 128) {
        int r = e / QP, a = e - r * QP;
        tile[r * LDT + a] = (r < nr && a < nb) ? A21[(r0 + r) * lda + a] : 0.0;
    }
    __syncthreads();
    double yv[QP];
#pragma unroll
    for (int a = 0; a < QP; ++a) yv[a] = tile[tid * LDT + a];
#pragma unroll
    for (int a = 0; a < QP; ++a) {
        double s0 = yv[a], s1 = 0.0, s2 = 0.0, s3 = 0.0;
#pragma unroll
        for (int c = 0; c < a; ++c) {
            const double rv = Ls[a * QP + c];
            if ((c & 3) == 0) s0 = fma(-rv, yv[c], s0);
            else if ((c & 3) == 1) s1 = fma(-rv, yv[c], s1);
            else if ((c & 3) == 2) s2 = fma(-rv, yv[c], s2);
            else s3 = fma(-rv, yv[c], s3);
        }
        yv[a] = ((s0 + s1) + (s2 + s3)) * rinv[a];
    }
#pragma unroll
    for (int a = 0; a < QP; ++a) tile[tid * LDT + a] = yv[a];
    __syncthreads();
    for (int e = tid; e < nr * nb; e += 128) {
        int r = e / nb, a = e - r * nb;
        A21[(r0 + r) * lda + a] = tile[r * LDT + a];
    }
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
// factor one panel: diagonal block (recursing with 64-wide blocks when wider than 64) and the rows below it
static int potrf_blocked(double* A, long long T, long long lda, int pb, int* info, int pivot_base, cudaStream_t st);
static int factor_panel(double* Akk, long long lda, int nb, long long rest, int* info, int pivot_base, cudaStream_t st) {
    const size_t smem_r = sizeof(double) * (64 * 64 + 64 + 128 * 65);
    if (nb > 64) {
        if (int r = potrf_blocked(Akk, nb, lda, 64, info, pivot_base, st)) return r;
    } else {
        k_potrf_diag<<<1, PD_TX * PD_TY, sizeof(double) * nb * (nb | 1), st>>>(Akk, lda, nb, info, pivot_base);
    }
    if (rest > 0) {
        // panel solve X L11^T = A21 by 64-wide column blocks with register-resident rows:
        //   A21[:, c0:c0+h] -= X[:, :c0] L11[c0:c0+h, :c0]^T ;  X[:, c0:c0+h] = A21[:, c0:c0+h] L11[c0.., c0..]^-T
        double* A21 = Akk + (long long)nb * lda;
        for (int c0 = 0; c0 < nb; c0 += 64) {
            const int h = nb - c0 > 64 ? 64 : nb - c0;
            if (c0 > 0)
                if (int r = gemm_nt_launch(A21, Akk + (long long)c0 * lda, A21 + c0, rest, h, c0, lda, lda, lda, -1.0, 1.0, 0, st))
                    return r;
            k_trsm_panel_reg64<<<(unsigned)((rest + 127) / 128), 128, smem_r, st>>>(Akk + (long long)c0 * lda + c0, A21 + c0,
                                                                                     lda, rest, h);
        }
    }
    return 0;
}
// blocked right-looking factorisation, panel width pb, everything on one stream (used for the diagonal blocks)
static int potrf_blocked(double* A, long long T, long long lda, int pb, int* info, int pivot_base, cudaStream_t st) {
    for (long long k = 0; k < T; k += pb) {
        const int nb = (int)min((long long)pb, T - k);
        const long long rest = T - k - nb;
        double* Akk = A + k * lda + k;
        if (int r = factor_panel(Akk, lda, nb, rest, info, pivot_base + (int)k, st)) return r;
        if (rest > 0) {
            double* A21 = Akk + (long long)nb * lda;
            if (int r = gemm_nt_launch(A21, A21, A21 + nb, rest, rest, nb, lda, lda, lda, -1.0, 1.0, 1, st)) return r;
        }
    }
    return 0;
}
// top level with look-ahead: the next panel is factorised on a helper stream while the main stream applies the bulk
// of the trailing update, so the latency-bound panel work hides behind the DMMA SYRK.
static cudaStream_t g_helper_stream = nullptr;
static cudaEvent_t g_ev_upd = nullptr, g_ev_pan = nullptr;
static int potrf_lookahead(double* A, long long T, long long lda, int pb, int* info, cudaStream_t s0) {
    if (!g_helper_stream) {
        if (cudaStreamCreateWithFlags(&g_helper_stream, cudaStreamNonBlocking) != cudaSuccess ||
            cudaEventCreateWithFlags(&g_ev_upd, cudaEventDisableTiming) != cudaSuccess ||
            cudaEventCreateWithFlags(&g_ev_pan, cudaEventDisableTiming) != cudaSuccess) {
            nmgp_set_error("nmgp_potrf_big: cannot create the look-ahead stream");
            return -4;
        }
    }
    cudaStream_t s1 = g_helper_stream;
    {   // panel 0 on the main stream
        const int nb = (int)min((long long)pb, T);
        if (int r = factor_panel(A, lda, nb, T - nb, info, 0, s0)) return r;
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
        if (int r = factor_panel(A22, lda, nb2, rest - nb2, info, (int)(k + nb), s1)) return r;
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
NMGP_API int nmgp_potrf_big(double* A, long long T, long long lda, double* hld, int* info, cudaStream_t st) {
    NMGP_REQUIRE(T > 0 && lda >= T && T < 2147483647LL, "nmgp_potrf_big");
    if (int r = nmgp_opt_in_smem(k_potrf_diag, sizeof(double) * 64 * 65, "nmgp_potrf_big")) return r;
    if (int r = nmgp_opt_in_smem(k_trsm_panel_reg64, sizeof(double) * (64 * 64 + 64 + 128 * 65), "nmgp_potrf_big")) return r;
    if (T > 1024) {
        int pb = T >= 12288 ? 512 : (T >= 6144 ? 256 : PB);   // measured best on B200 (profiles/README.md)
        if (const char* e = getenv("NMGP_POTRF_PB")) pb = atoi(e) >= 64 ? (atoi(e) / 64) * 64 : pb;   // tuning knob
        if (int r = potrf_lookahead(A, T, lda, pb, info, st)) return r;
    } else {
        if (int r = potrf_blocked(A, T, lda, 64, info, 0, st)) return r;
    }
    dim3 gz((unsigned)((T + 255) / 256), (unsigned)min(T, 65535LL));
    if (T <= 65535) k_zero_upper<<<gz, 256, 0, st>>>(A, T, lda);
    if (hld) k_logdiag_sum<<<1, 1024, 0, st>>>(A, T, lda, hld);
    return nmgp_launch_status("nmgp_potrf_big");
}

// ------------------------------------------------------------------------------------------------------------
// x <- (L L^T)^-1 x for one vector: blocked forward then backward substitution (PB-wide diagonal solves in one CTA,
// the remaining update as a memory-bound matrix-vector product).
__global__ void __launch_bounds__(PB)
k_trsv_diag(const double* __restrict__ L, long long lda, double* __restrict__ x, int nb, int transposed) {
    extern __shared__ double sm[];
    double* Ls = sm;            // [nb][nb|1]
    double* xs = Ls + nb * (nb | 1);
    const int ld = nb | 1, tid = threadIdx.x;
    for (int e = tid; e < nb * nb; e += blockDim.x) {
        int a = e / nb, b = e - a * nb;
        Ls[a * ld + b] = L[(long long)a * lda + b];
    }
    if (tid < nb) xs[tid] = x[tid];
    __syncthreads();
    if (!transposed) {
        for (int a = 0; a < nb; ++a) {
            if (tid == a) xs[a] /= Ls[a * ld + a];
            __syncthreads();
            if (tid > a && tid < nb) xs[tid] = fma(-Ls[tid * ld + a], xs[a], xs[tid]);
            __syncthreads();
        }
    } else {
        for (int a = nb - 1; a >= 0; --a) {
            if (tid == a) xs[a] /= Ls[a * ld + a];
            __syncthreads();
            if (tid < a) xs[tid] = fma(-Ls[a * ld + tid], xs[a], xs[tid]);
            __syncthreads();
        }
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
    const size_t smem = sizeof(double) * (PB * (PB | 1) + PB);
    if (int r = nmgp_opt_in_smem(k_trsv_diag, smem, "nmgp_potrs_vec")) return r;
    for (long long k = 0; k < T; k += PB) {                       // L y = b
        const int nb = (int)min((long long)PB, T - k);
        k_trsv_diag<<<1, PB, sizeof(double) * (nb * (nb | 1) + nb), st>>>(L + k * lda + k, lda, x + k, nb, 0);
        const long long rest = T - k - nb;
        if (rest > 0)
            k_gemv_sub<<<(unsigned)((rest + 7) / 8), 256, 0, st>>>(L + (k + nb) * lda + k, lda, x + k, x + k + nb, rest, nb);
    }
    for (long long k = ((T - 1) / PB) * PB; k >= 0; k -= PB) {    // L^T x = y
        const int nb = (int)min((long long)PB, T - k);
        k_trsv_diag<<<1, PB, sizeof(double) * (nb * (nb | 1) + nb), st>>>(L + k * lda + k, lda, x + k, nb, 1);
        if (k > 0)   // x[0:k] -= L[k:k+nb, 0:k]^T x[k:k+nb]
            k_gemv_t_sub<<<(unsigned)((k + 127) / 128), 128, 0, st>>>(L + k * lda, lda, x + k, x, nb, k);
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
NMGP_API int nmgp_scale_add_diag(const double* K, double* A, long long T, double alpha, double sigma2,
                                 cudaStream_t st) {
    NMGP_REQUIRE(T > 0, "nmgp_scale_add_diag");
    dim3 grid((unsigned)((T + 255) / 256), (unsigned)min(T, 65535LL), (unsigned)((T + 65534) / 65535));
    k_scale_add_diag<<<grid, 256, 0, st>>>(K, A, T, alpha, sigma2);
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
    k_kron<<<(unsigned)((total + 255) / 256), 256, 0, st>>>(t1, t2, out, h1, w1, h2, w2);
    return nmgp_launch_status("nmgp_kron_product");
}

// ------------------------------------------------------------------------------------------------------------
// Cyclic Jacobi eigen-decomposition of a small symmetric n x n matrix (n <= 128), one CTA.  Reads the UPPER triangle
// (torch.symeig's default, kronecker_operation.py:45).  Eigenvalues ascending in w, eigenvectors in the columns of V.
__global__ void __launch_bounds__(128)
k_eigh_jacobi(const double* __restrict__ A, double* __restrict__ w, double* __restrict__ V,
              double* __restrict__ Vwork, int n) {
    extern __shared__ double sm[];
    double* S = sm;              // [n][n]
    double* U = Vwork;           // [n][n] eigenvector accumulator (global scratch, L2-resident)
    __shared__ double cs[2];
    __shared__ int order[128];
    const int tid = threadIdx.x;
    for (int e = tid; e < n * n; e += blockDim.x) {
        int a = e / n, b = e - a * n;
        S[e] = (b >= a) ? A[a * n + b] : A[b * n + a];
        U[e] = (a == b) ? 1.0 : 0.0;
    }
    __syncthreads();
    for (int sweep = 0; sweep < 30; ++sweep) {
        double off = 0.0;
        for (int e = tid; e < n * n; e += blockDim.x) {
            int a = e / n, b = e - a * n;
            if (a != b) off = fma(S[e], S[e], off);
        }
        off = block_sum(off);
        double diag = 0.0;
        for (int a = tid; a < n; a += blockDim.x) diag = fma(S[a * n + a], S[a * n + a], diag);
        diag = block_sum(diag);
        if (off <= 1e-34 * diag) break;
        for (int p = 0; p < n - 1; ++p) {
            for (int q = p + 1; q < n; ++q) {
                if (tid == 0) {
                    double apq = S[p * n + q];
                    double c = 1.0, s = 0.0;
                    if (apq != 0.0) {
                        double tau = (S[q * n + q] - S[p * n + p]) / (2.0 * apq);
                        double tt = (tau >= 0.0 ? 1.0 : -1.0) / (fabs(tau) + sqrt(1.0 + tau * tau));
                        c = 1.0 / sqrt(1.0 + tt * tt);
                        s = tt * c;
                    }
                    cs[0] = c; cs[1] = s;
                }
                __syncthreads();
                const double c = cs[0], s = cs[1];
                if (s != 0.0) {
                    for (int k = tid; k < n; k += blockDim.x) {      // columns p, q of S and U
                        double skp = S[k * n + p], skq = S[k * n + q];
                        S[k * n + p] = c * skp - s * skq;
                        S[k * n + q] = s * skp + c * skq;
                        double ukp = U[k * n + p], ukq = U[k * n + q];
                        U[k * n + p] = c * ukp - s * ukq;
                        U[k * n + q] = s * ukp + c * ukq;
                    }
                    __syncthreads();
                    for (int k = tid; k < n; k += blockDim.x) {      // rows p, q of S
                        double spk = S[p * n + k], sqk = S[q * n + k];
                        S[p * n + k] = c * spk - s * sqk;
                        S[q * n + k] = s * spk + c * sqk;
                    }
                }
                __syncthreads();
            }
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
    }
    __syncthreads();
    if (tid < n) w[tid] = S[order[tid] * n + order[tid]];
    for (int e = tid; e < n * n; e += blockDim.x) {
        int a = e / n, b = e - a * n;
        V[e] = U[a * n + order[b]];
    }
}
NMGP_API int nmgp_eigh_small(const double* A, double* w, double* V, double* work /* n*n */, int n, cudaStream_t st) {
    NMGP_REQUIRE(n > 0 && n <= 128 && work != nullptr, "nmgp_eigh_small");
    size_t smem = sizeof(double) * n * n;
    if (int r = nmgp_opt_in_smem(k_eigh_jacobi, smem, "nmgp_eigh_small")) return r;
    k_eigh_jacobi<<<1, 128, smem, st>>>(A, w, V, work, n);
    return nmgp_launch_status("nmgp_eigh_small");
}

// ------------------------------------------------------------------------------------------------------------
// small vector helpers of the Kronecker log-density (distributions.py:42-51)
__global__ void k_axpby(const double* __restrict__ x, const double* __restrict__ y, double* __restrict__ out,
                        long long n, double a, double b) {
    long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) out[i] = a * x[i] + b * y[i];
}
NMGP_API int nmgp_axpby(const double* x, const double* y, double* out, long long n, double a, double b,
                        cudaStream_t st) {
    if (n <= 0) return 0;
    k_axpby<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(x, y, out, n, a, b);
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
    k_dot<<<(unsigned)blocks, 256, 0, st>>>(x, y, out, n);
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
    k_pairwise<<<grid, 256, 0, st>>>(X1, X2, out, T1, T2, dx);
    return nmgp_launch_status("nmgp_pairwise_dist");
}
