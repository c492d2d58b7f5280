// Small batched Q x Q kernels of the DSVI step: Sigma = tril(S) tril(S)^T, Cholesky (+ log-det),
// their adjoints, the reference-exact Gaussian KL (quirk q10) and the inducing draw of v.
// One CTA per matrix, matrix resident in shared memory (Q <= 112), FP64 FMA.
// Reference sites: code/nmgp_dsvi.py:172-177 (LL^T), code/utils.py:46,276,347-348 (cholesky),
// code/utils.py:332-351 (KL_Gaussian), code/utils.py:225-227 (sample of v).
#include <stdarg.h>
#include <string.h>

#include "common.cuh"

// ------------------------------------------------------------------------------------------
// error plumbing
static thread_local char g_err[512] = "";
void nmgp_set_error(const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
}
int nmgp_launch_status(const char* what) {
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) {
        nmgp_set_error("%s: CUDA launch failed: %s", what, cudaGetErrorString(e));
        return -10;
    }
    return 0;
}
NMGP_API const char* nmgp_last_error(void) { return g_err; }
NMGP_API int nmgp_version(void) { return 200; }
// kernel launches issued by this library since it was loaded (host-side counter; CUDA-graph replays are added by the
// host code that replays them)
unsigned long long g_nmgp_launches = 0;
NMGP_API unsigned long long nmgp_launch_count(void) { return g_nmgp_launches; }

// ------------------------------------------------------------------------------------------
__global__ void k_hyper_exp(const double* __restrict__ logs, double* __restrict__ out, int n) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) out[i] = exp(logs[i]);
}
NMGP_API int nmgp_hyper_exp(const double* logs, double* hyp, int n, cudaStream_t st) {
    NMGP_REQUIRE(n > 0, "nmgp_hyper_exp");
    k_hyper_exp<<<NMGP_L((n + 63) / 64), 64, 0, st>>>(logs, hyp, n);
    return nmgp_launch_status("nmgp_hyper_exp");
}

// seg[d] = first row with I >= d (I sorted ascending), seg[D] = B
__global__ void k_segment_offsets(const int* __restrict__ I, int* __restrict__ seg, long long B, int D) {
    int d = blockIdx.x * blockDim.x + threadIdx.x;
    if (d > D) return;
    long long lo = 0, hi = B;
    while (lo < hi) {
        long long mid = (lo + hi) >> 1;
        if (I[mid] < d) lo = mid + 1; else hi = mid;
    }
    seg[d] = (int)lo;
}
NMGP_API int nmgp_segment_offsets(const int* I, int* seg, long long B, int D, cudaStream_t st) {
    NMGP_REQUIRE(B >= 0 && D > 0 && B < 2147483647LL, "nmgp_segment_offsets");
    k_segment_offsets<<<NMGP_L((D + 1 + 127) / 128), 128, 0, st>>>(I, seg, B, D);
    return nmgp_launch_status("nmgp_segment_offsets");
}

// ------------------------------------------------------------------------------------------
// Sigma = tril(S) tril(S)^T
__global__ void k_tril_syrk_fwd(const double* __restrict__ S, double* __restrict__ Sig, int Q) {
    extern __shared__ double sm[];
    const int ld = Q | 1;
    const double* Sm = S + (size_t)blockIdx.x * Q * Q;
    double* Om = Sig + (size_t)blockIdx.x * Q * Q;
    for (int e = threadIdx.x; e < Q * Q; e += blockDim.x) {
        int a = e / Q, b = e - a * Q;
        sm[a * ld + b] = (b <= a) ? Sm[e] : 0.0;
    }
    __syncthreads();
    for (int e = threadIdx.x; e < Q * Q; e += blockDim.x) {
        int a = e / Q, b = e - a * Q;
        int kmax = a < b ? a : b;
        double s = 0.0;
        for (int c = 0; c <= kmax; ++c) s = fma(sm[a * ld + c], sm[b * ld + c], s);
        Om[e] = s;
    }
}
NMGP_API int nmgp_tril_syrk_fwd(const double* S, double* Sigma, int nb, int Q, cudaStream_t st) {
    NMGP_REQUIRE(nb >= 0 && Q > 0, "nmgp_tril_syrk_fwd");
    if (nb == 0) return 0;
    size_t smem = (size_t)Q * (Q | 1) * sizeof(double);
    if (int r = nmgp_opt_in_smem(k_tril_syrk_fwd, smem, "nmgp_tril_syrk_fwd")) return r;
    k_tril_syrk_fwd<<<NMGP_L(nb), 256, smem, st>>>(S, Sigma, Q);
    return nmgp_launch_status("nmgp_tril_syrk_fwd");
}

// Sbar = tril((G + G^T) L), L = tril(S)
__global__ void k_tril_syrk_bwd(const double* __restrict__ S, const double* __restrict__ G,
                                double* __restrict__ Sbar, int Q) {
    extern __shared__ double sm[];
    const int ld = Q | 1;
    double* L = sm;
    double* M = sm + (size_t)Q * ld;
    const size_t off = (size_t)blockIdx.x * Q * Q;
    for (int e = threadIdx.x; e < Q * Q; e += blockDim.x) {
        int a = e / Q, b = e - a * Q;
        L[a * ld + b] = (b <= a) ? S[off + e] : 0.0;
        M[a * ld + b] = G[off + e] + G[off + (size_t)b * Q + a];
    }
    __syncthreads();
    for (int e = threadIdx.x; e < Q * Q; e += blockDim.x) {
        int a = e / Q, b = e - a * Q;
        double s = 0.0;
        if (b <= a)
            for (int c = b; c < Q; ++c) s = fma(M[a * ld + c], L[c * ld + b], s);
        Sbar[off + e] = s;
    }
}
NMGP_API int nmgp_tril_syrk_bwd(const double* S, const double* SigBar, double* Sbar, int nb, int Q, cudaStream_t st) {
    NMGP_REQUIRE(nb >= 0 && Q > 0, "nmgp_tril_syrk_bwd");
    if (nb == 0) return 0;
    size_t smem = 2 * (size_t)Q * (Q | 1) * sizeof(double);
    if (int r = nmgp_opt_in_smem(k_tril_syrk_bwd, smem, "nmgp_tril_syrk_bwd")) return r;
    k_tril_syrk_bwd<<<NMGP_L(nb), 256, smem, st>>>(S, SigBar, Sbar, Q);
    return nmgp_launch_status("nmgp_tril_syrk_bwd");
}

// ------------------------------------------------------------------------------------------
// C = chol(A + jitter I) (lower, strict upper zeroed), hld = sum log diag C.  Thread i owns row i.
__global__ void k_potrf(const double* __restrict__ A, double jitter, double* __restrict__ C,
                        double* __restrict__ hld, int* __restrict__ info, int Q) {
    extern __shared__ double sm[];
    __shared__ double s_piv;
    const int ld = Q | 1;
    const size_t off = (size_t)blockIdx.x * Q * Q;
    for (int e = threadIdx.x; e < Q * Q; e += blockDim.x) {
        int a = e / Q, b = e - a * Q;
        sm[a * ld + b] = A[off + e] + (a == b ? jitter : 0.0);
    }
    __syncthreads();
    const int i = threadIdx.x;
    for (int k = 0; k < Q; ++k) {
        double s = 0.0;
        if (i >= k && i < Q) {
            double s0 = sm[i * ld + k], s1 = 0.0;
            int c = 0;
            for (; c + 1 < k; c += 2) {
                s0 = fma(-sm[i * ld + c], sm[k * ld + c], s0);
                s1 = fma(-sm[i * ld + c + 1], sm[k * ld + c + 1], s1);
            }
            if (c < k) s0 = fma(-sm[i * ld + c], sm[k * ld + c], s0);
            s = s0 + s1;
            if (i == k) {
                if (!(s > 0.0)) atomicMax(info, (int)blockIdx.x + 1);
                s_piv = sqrt(s);
            }
        }
        __syncthreads();
        if (i >= k && i < Q) sm[i * ld + k] = (i == k) ? s_piv : s / s_piv;
        __syncthreads();
    }
    double lg = (i < Q) ? log(sm[i * ld + i]) : 0.0;
    lg = block_sum(lg);
    if (threadIdx.x == 0) hld[blockIdx.x] = lg;
    for (int e = threadIdx.x; e < Q * Q; e += blockDim.x) {
        int a = e / Q, b = e - a * Q;
        C[off + e] = (b <= a) ? sm[a * ld + b] : 0.0;
    }
}
int nmgp_potrf_batched_blocked(const double* A, double jitter, double* C, double* hld, int* info, int nb_mat, int Q,
                               cudaStream_t st);   // nmgp_dense.cu: DMMA blocked factorisation for 64 < Q <= 128
NMGP_API int nmgp_potrf_batched(const double* A, double jitter, double* C, double* hld, int* info, int nb, int Q,
                                cudaStream_t st) {
    NMGP_REQUIRE(nb >= 0 && Q > 0 && Q <= 128, "nmgp_potrf_batched");
    if (nb == 0) return 0;
    if (Q > 64) return nmgp_potrf_batched_blocked(A, jitter, C, hld, info, nb, Q, st);
    size_t smem = (size_t)Q * (Q | 1) * sizeof(double);
    if (int r = nmgp_opt_in_smem(k_potrf, smem, "nmgp_potrf_batched")) return r;
    k_potrf<<<NMGP_L(nb), 128, smem, st>>>(A, jitter, C, hld, info, Q);
    return nmgp_launch_status("nmgp_potrf_batched");
}

// Abar = sym(C^-T Phi(C^T Cb) C^-1), Cb = tril(Cbar) + diag(hldbar / diag C).  Thread b owns column b.
__global__ void k_potrf_bwd(const double* __restrict__ Cg, const double* __restrict__ Cbar,
                            const double* __restrict__ hldbar, double* __restrict__ Abar, int Q) {
    extern __shared__ double sm[];
    const int ld = Q | 1;
    double* C = sm;
    double* W = sm + (size_t)Q * ld;
    const size_t off = (size_t)blockIdx.x * Q * Q;
    const double hb = hldbar[blockIdx.x];
    for (int e = threadIdx.x; e < Q * Q; e += blockDim.x) {
        int a = e / Q, b = e - a * Q;
        double c = (b <= a) ? Cg[off + e] : 0.0;
        C[a * ld + b] = c;
        double w = (b <= a) ? Cbar[off + e] : 0.0;
        if (a == b) w += hb / c;
        W[a * ld + b] = w;
    }
    __syncthreads();
    const int b = threadIdx.x;
    // step 1: W <- Phi(C^T W) in place (rows ascending; entry (a,b) only reads rows >= a of column b)
    if (b < Q) {
        for (int a = 0; a < Q; ++a) {
            double s = 0.0;
            if (a >= b) {
                double s0 = 0.0, s1 = 0.0;
                int c = a;
                for (; c + 1 < Q; c += 2) {
                    s0 = fma(C[c * ld + a], W[c * ld + b], s0);
                    s1 = fma(C[(c + 1) * ld + a], W[(c + 1) * ld + b], s1);
                }
                if (c < Q) s0 = fma(C[c * ld + a], W[c * ld + b], s0);
                s = s0 + s1;
                if (a == b) s *= 0.5;
            }
            W[a * ld + b] = s;
        }
    }
    __syncthreads();
    // step 2: W <- C^-T W (back substitution down each column)
    if (b < Q) {
        for (int a = Q - 1; a >= 0; --a) {
            double s0 = W[a * ld + b], s1 = 0.0;
            int c = a + 1;
            for (; c + 1 < Q; c += 2) {
                s0 = fma(-C[c * ld + a], W[c * ld + b], s0);
                s1 = fma(-C[(c + 1) * ld + a], W[(c + 1) * ld + b], s1);
            }
            if (c < Q) s0 = fma(-C[c * ld + a], W[c * ld + b], s0);
            W[a * ld + b] = (s0 + s1) / C[a * ld + a];
        }
    }
    __syncthreads();
    // step 3: W <- W C^-1, i.e. row b of W solved against C^T from the right
    if (b < Q) {
        for (int a = Q - 1; a >= 0; --a) {
            double s0 = W[b * ld + a], s1 = 0.0;
            int c = a + 1;
            for (; c + 1 < Q; c += 2) {
                s0 = fma(-C[c * ld + a], W[b * ld + c], s0);
                s1 = fma(-C[(c + 1) * ld + a], W[b * ld + c + 1], s1);
            }
            if (c < Q) s0 = fma(-C[c * ld + a], W[b * ld + c], s0);
            W[b * ld + a] = (s0 + s1) / C[a * ld + a];
        }
    }
    __syncthreads();
    for (int e = threadIdx.x; e < Q * Q; e += blockDim.x) {
        int a = e / Q, c = e - a * Q;
        Abar[off + e] = 0.5 * (W[a * ld + c] + W[c * ld + a]);
    }
}
NMGP_API int nmgp_potrf_bwd_batched(const double* C, const double* Cbar, const double* hldbar, double* Abar, int nb,
                                    int Q, cudaStream_t st) {
    NMGP_REQUIRE(nb >= 0 && Q > 0 && Q <= 128, "nmgp_potrf_bwd_batched");
    if (nb == 0) return 0;
    size_t smem = 2 * (size_t)Q * (Q | 1) * sizeof(double);
    if (int r = nmgp_opt_in_smem(k_potrf_bwd, smem, "nmgp_potrf_bwd_batched")) return r;
    k_potrf_bwd<<<NMGP_L(nb), 128, smem, st>>>(C, Cbar, hldbar, Abar, Q);
    return nmgp_launch_status("nmgp_potrf_bwd_batched");
}

// ---- 64 < Q <= 128: the same adjoint on the tensor cores ------------------------------------------------------------
// Abar = sym(C^-T Phi(C^T Cb) C^-1) = sym(((Phi^T C^-1)^T) C^-1): one DMMA reduction (C^T Cb, k_atb_mma) and two right
// solves "rows x C^-1" (the row-solve kernel's second sweep) with a transpose in between -- no per-column serial
// substitution.  work1 / work2: [nb, Q, Q] scratch.
int nmgp_right_solve_rows(const double* K, const double* R, double* X, int ns, long long B, int Q, cudaStream_t st);
int nmgp_atb_mma(const double* A, const double* Bm, double* C, double sign, int ns, long long B, int Q,
                 const double* cbar, double* Kbar, cudaStream_t st);
// W[b] = tril(Cbar[b]) + diag(hldbar[b] / diag C[b])
__global__ void k_pbw_prep(const double* __restrict__ C, const double* __restrict__ Cbar, const double* __restrict__ hldbar,
                           double* __restrict__ W, long long n, int Q) {
    const long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= n) return;
    const long long m = e / ((long long)Q * Q);
    const int r = (int)(e % ((long long)Q * Q)), a = r / Q, b = r - a * Q;
    double w = (b <= a) ? Cbar[e] : 0.0;
    if (a == b) w += hldbar[m] / C[e];
    W[e] = w;
}
// out[b] = Phi(M[b])^T  (Phi: lower triangle with halved diagonal)   or, phi == 0, plain transpose
__global__ void k_pbw_transpose(const double* __restrict__ M, double* __restrict__ out, long long n, int Q, int phi) {
    const long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= n) return;
    const long long m = e / ((long long)Q * Q);
    const int r = (int)(e % ((long long)Q * Q)), a = r / Q, b = r - a * Q;       // out[a,b] = f(M[b,a])
    double v = M[m * Q * Q + (long long)b * Q + a];
    if (phi) v = (a > b) ? 0.0 : (a == b ? 0.5 * v : v);                          // M[b,a] with b >= a kept
    out[e] = v;
}
__global__ void k_pbw_sym(const double* __restrict__ M, double* __restrict__ out, long long n, int Q) {
    const long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= n) return;
    const long long m = e / ((long long)Q * Q);
    const int r = (int)(e % ((long long)Q * Q)), a = r / Q, b = r - a * Q;
    out[e] = 0.5 * (M[e] + M[m * Q * Q + (long long)b * Q + a]);
}
NMGP_API int nmgp_potrf_bwd_batched_lq(const double* C, const double* Cbar, const double* hldbar, double* Abar,
                                       double* work1, double* work2, int nb, int Q, cudaStream_t st) {
    NMGP_REQUIRE(nb >= 0 && Q > 64 && Q <= 128, "nmgp_potrf_bwd_batched_lq");
    if (nb == 0) return 0;
    const long long n = (long long)nb * Q * Q;
    const unsigned blocks = (unsigned)((n + 255) / 256);
    k_pbw_prep<<<NMGP_L(blocks), 256, 0, st>>>(C, Cbar, hldbar, work1, n, Q);                      // work1 = Cb
    if (cudaMemsetAsync(work2, 0, sizeof(double) * n, st) != cudaSuccess) {
        nmgp_set_error("nmgp_potrf_bwd_batched_lq: cudaMemsetAsync failed");
        return -11;
    }
    if (int r = nmgp_atb_mma(C, work1, work2, 1.0, nb, Q, Q, nullptr, nullptr, st)) return r;        // work2 = C^T Cb
    k_pbw_transpose<<<NMGP_L(blocks), 256, 0, st>>>(work2, work1, n, Q, 1);                          // work1 = Phi^T
    if (int r = nmgp_right_solve_rows(work1, C, work2, nb, Q, Q, st)) return r;                      // work2 = Phi^T C^-1
    k_pbw_transpose<<<NMGP_L(blocks), 256, 0, st>>>(work2, work1, n, Q, 0);                          // work1 = C^-T Phi
    if (int r = nmgp_right_solve_rows(work1, C, work2, nb, Q, Q, st)) return r;                      // work2 = C^-T Phi C^-1
    k_pbw_sym<<<NMGP_L(blocks), 256, 0, st>>>(work2, Abar, n, Q);
    return nmgp_launch_status("nmgp_potrf_bwd_batched_lq");
}


// ------------------------------------------------------------------------------------------
// KL(N(mu_b, C_b C_b^T) || N(0, R_p R_p^T)) in the reference's form (code/utils.py:346-351, quirk q10):
//   kl[p,b] = hldR[p] - hldS[b] + 0.5 ( sum_a rs[b,a] / R_p[a,a]^2 + mu_b^T (R_p R_p^T)^-1 mu_b - Q ),
//   rs[b,a] = sum_{c<=a} C_b[a,c]^2.
// Batched formulation (every (p,b) pair in parallel, no per-pair serial substitution):
//   * rs once per b (one warp per matrix row, coalesced);
//   * t[p,b,:] = (R_p R_p^T)^-1 mu_b and the Mahalanobis term mu_b . t[p,b,:] are exactly what the DMMA row-solve
//     kernel produces for the "rows" mu_b (nmgp_solve_rows_fwd_mma: P = K (R R^T)^-1, c = rowsum(P o K));
//   * adjoint: mubar_b = sum_p klbar t;  A_p = R_p R_p^T receives -1/2 sum_b klbar t t^T, i.e.
//     Rbar_p -= tril(G_p R_p) with the weighted Gram matrix G_p = sum_b klbar[p,b] t t^T (DMMA reduction k_atb_mma);
//     the diagonal-only trace term contributes Rbar_p[a,a] -= sum_b klbar rs[b,a] / d^3 and
//     CSbar_b = tril(2 rsb[b,a] C_b), rsb[b,a] = 1/2 sum_p klbar / R_p[a,a]^2.
int nmgp_solve_rows_fwd_mma(const double* K, const double* R, double* P, double* c, int ns, long long B, int Q,
                            cudaStream_t st);
int nmgp_atb_mma(const double* A, const double* Bm, double* C, double sign, int ns, long long B, int Q,
                 const double* cbar, double* Kbar, cudaStream_t st);

// rs[b,a] = sum_{c<=a} CS[b,a,c]^2 : one warp per (b,a)
__global__ void k_kl_rowsq(const double* __restrict__ CS, double* __restrict__ rs, long long nrows, int Q) {
    const long long row = (long long)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (row >= nrows) return;
    const int lane = threadIdx.x & 31, a = (int)(row % Q);
    const double* src = CS + row * Q;
    double s = 0.0;
    for (int c = lane; c <= a; c += 32) s = fma(src[c], src[c], s);
    s = warp_sum(s);
    if (lane == 0) rs[row] = s;
}
// out[p,b,:] = mu[b,:]
__global__ void k_kl_expand(const double* __restrict__ mu, double* __restrict__ out, int np_, long long nbQ) {
    const long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= nbQ) return;
    const double v = mu[e];
    for (int p = 0; p < np_; ++p) out[(size_t)p * nbQ + e] = v;
}
// kl[p,b] (in: the Mahalanobis term) = hldR[p] - hldS[b] + 0.5 (sum_a rs[b,a] w_p[a] + kl[p,b] - Q), w_p = 1/diag(R_p)^2
__global__ void k_kl_combine(double* __restrict__ kl, const double* __restrict__ rs, const double* __restrict__ R,
                             const double* __restrict__ hldR, const double* __restrict__ hldS, int nb, int Q) {
    extern __shared__ double wsm[];
    const int p = blockIdx.y;
    for (int a = threadIdx.x; a < Q; a += blockDim.x) {
        const double d = R[(size_t)p * Q * Q + (size_t)a * Q + a];
        wsm[a] = 1.0 / (d * d);
    }
    __syncthreads();
    const int wpb = blockDim.x >> 5, lane = threadIdx.x & 31;
    for (int b = blockIdx.x * wpb + (threadIdx.x >> 5); b < nb; b += gridDim.x * wpb) {
        double s = 0.0;
        for (int a = lane; a < Q; a += 32) s = fma(rs[(size_t)b * Q + a], wsm[a], s);
        s = warp_sum(s);
        if (lane == 0) {
            const size_t o = (size_t)p * nb + b;
            kl[o] = hldR[p] - hldS[b] + 0.5 * (s + kl[o] - (double)Q);
        }
    }
}
NMGP_API int nmgp_kl_fwd(const double* CS, const double* hldS, const double* mu, const double* R, const double* hldR,
                         double* kl, double* t, double* rs, double* work, int np_, int nb, int Q, cudaStream_t st) {
    NMGP_REQUIRE(np_ > 0 && nb > 0 && Q > 0 && Q <= 128, "nmgp_kl_fwd");
    const long long nrows = (long long)nb * Q;
    k_kl_rowsq<<<NMGP_L((unsigned)((nrows + 7) / 8)), 256, 0, st>>>(CS, rs, nrows, Q);
    k_kl_expand<<<NMGP_L((unsigned)((nrows + 255) / 256)), 256, 0, st>>>(mu, work, np_, nrows);
    if (int r = nmgp_solve_rows_fwd_mma(work, R, t, kl, np_, nb, Q, st)) return r;
    dim3 grid((unsigned)min((nb + 7) / 8, 1024), np_);
    k_kl_combine<<<NMGP_L(grid), 256, Q * sizeof(double), st>>>(kl, rs, R, hldR, hldS, nb, Q);
    return nmgp_launch_status("nmgp_kl_fwd");
}

// work[p,b,a] = klbar[p,b] t[p,b,a];  mubar[b,a] = sum_p work;  rsb[b,a] = 1/2 sum_p klbar[p,b] / R_p[a,a]^2;
// hldSbar[b] = -sum_p klbar[p,b].   Thread per (b,a).
__global__ void k_kl_bwd_rows(const double* __restrict__ klbar, const double* __restrict__ R, const double* __restrict__ t,
                              double* __restrict__ work, double* __restrict__ mubar, double* __restrict__ rsb,
                              double* __restrict__ hldSbar, int np_, int nb, int Q) {
    const long long gid = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (gid >= (long long)nb * Q) return;
    const int b = (int)(gid / Q), a = (int)(gid - (long long)b * Q);
    double mb = 0.0, rb = 0.0, ks = 0.0;
    for (int p = 0; p < np_; ++p) {
        const double kb = klbar[(size_t)p * nb + b];
        const size_t o = ((size_t)p * nb + b) * Q + a;
        const double wv = kb * t[o];
        work[o] = wv;
        mb += wv;
        const double d = R[(size_t)p * Q * Q + (size_t)a * Q + a];
        rb = fma(0.5 * kb, 1.0 / (d * d), rb);
        ks += kb;
    }
    mubar[gid] = mb;
    rsb[gid] = rb;
    if (a == 0) hldSbar[b] = -ks;
}
// CSbar[b,a,c] = (c <= a) ? 2 rsb[b,a] CS[b,a,c] : 0
__global__ void k_kl_bwd_CS(const double* __restrict__ CS, const double* __restrict__ rsb, double* __restrict__ CSbar,
                            long long n, int Q) {
    const long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= n) return;
    const long long row = e / Q;
    const int c = (int)(e - row * Q), a = (int)(row % Q);
    CSbar[e] = (c <= a) ? 2.0 * rsb[row] * CS[e] : 0.0;
}
// Rbar_p[a,c] += -(G_p R_p)[a,c] (c <= a);  Rbar_p[a,a] += -sum_b klbar[p,b] rs[b,a] / d^3.  grid (ceil(Q/8), np), one
// warp per row a.
__global__ void k_kl_bwd_R(const double* __restrict__ klbar, const double* __restrict__ rs, const double* __restrict__ R,
                           const double* __restrict__ G, double* __restrict__ Rbar, int nb, int Q) {
    const int p = blockIdx.y, a = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5), lane = threadIdx.x & 31;
    if (a >= Q) return;
    const double* Gp = G + (size_t)p * Q * Q + (size_t)a * Q;
    const double* Rp = R + (size_t)p * Q * Q;
    double* Ob = Rbar + (size_t)p * Q * Q + (size_t)a * Q;
    for (int c = lane; c <= a; c += 32) {
        double s0 = 0.0, s1 = 0.0;
        int k = c;                                   // R is lower triangular: R[k,c] = 0 for k < c
        for (; k + 1 < Q; k += 2) {
            s0 = fma(Gp[k], Rp[(size_t)k * Q + c], s0);
            s1 = fma(Gp[k + 1], Rp[(size_t)(k + 1) * Q + c], s1);
        }
        if (k < Q) s0 = fma(Gp[k], Rp[(size_t)k * Q + c], s0);
        Ob[c] -= s0 + s1;
    }
    if (rs == nullptr) return;                       // exact-KL variant: no diagonal-only trace term
    double ws = 0.0;
    for (int b = lane; b < nb; b += 32) ws = fma(klbar[(size_t)p * nb + b], rs[(size_t)b * Q + a], ws);
    ws = warp_sum(ws);
    if (lane == 0) {
        const double d = Rp[(size_t)a * Q + a];
        Ob[a] -= ws / (d * d * d);
    }
}
__global__ void k_kl_bwd_hldR(const double* __restrict__ klbar, double* __restrict__ hldRbar, int np_, int nb) {
    int p = blockIdx.x;
    double s = 0.0;
    for (int b = threadIdx.x; b < nb; b += blockDim.x) s += klbar[(size_t)p * nb + b];
    s = block_sum(s);
    if (threadIdx.x == 0) hldRbar[p] = s;
}
NMGP_API int nmgp_kl_bwd(const double* klbar, const double* CS, const double* R, const double* t, const double* rs,
                         double* CSbar, double* hldSbar, double* mubar, double* Rbar /* += */, double* hldRbar,
                         double* work /* [np,nb,Q] */, double* rsb /* [nb,Q] */, double* G /* [np,Q,Q] */, int np_, int nb,
                         int Q, cudaStream_t st) {
    NMGP_REQUIRE(np_ > 0 && nb > 0 && Q > 0 && Q <= 128, "nmgp_kl_bwd");
    const long long n3 = (long long)nb * Q;
    k_kl_bwd_rows<<<NMGP_L((unsigned)((n3 + 127) / 128)), 128, 0, st>>>(klbar, R, t, work, mubar, rsb, hldSbar, np_, nb, Q);
    const long long n4 = n3 * Q;
    k_kl_bwd_CS<<<NMGP_L((unsigned)((n4 + 255) / 256)), 256, 0, st>>>(CS, rsb, CSbar, n4, Q);
    if (cudaMemsetAsync(G, 0, sizeof(double) * (size_t)np_ * Q * Q, st) != cudaSuccess) {
        nmgp_set_error("nmgp_kl_bwd: cudaMemsetAsync failed");
        return -11;
    }
    if (int r = nmgp_atb_mma(work, t, G, 1.0, np_, nb, Q, nullptr, nullptr, st)) return r;
    dim3 g2((Q + 7) / 8, np_);
    k_kl_bwd_R<<<NMGP_L(g2), 256, 0, st>>>(klbar, rs, R, G, Rbar, nb, Q);
    k_kl_bwd_hldR<<<NMGP_L(np_), 128, 0, st>>>(klbar, hldRbar, np_, nb);
    return nmgp_launch_status("nmgp_kl_bwd");
}

// Building blocks of the mathematically exact KL variant (flag of utils.KL_Gaussian / NMGP(exact_kl=True), quirk q10):
// C[s] += sign * A[s]^T B[s] over the rows (DMMA reduction) and Rbar_p += -tril(G_p R_p).
NMGP_API int nmgp_atb(const double* A, const double* Bm, double* C /* += */, double sign, int ns, long long B, int Q,
                      cudaStream_t st) {
    NMGP_REQUIRE(ns > 0 && B >= 0 && Q > 0 && Q <= 128, "nmgp_atb");
    if (B == 0) return 0;
    return nmgp_atb_mma(A, Bm, C, sign, ns, B, Q, nullptr, nullptr, st);
}
NMGP_API int nmgp_kl_rbar(const double* R, const double* G, double* Rbar /* += */, int np_, int Q, cudaStream_t st) {
    NMGP_REQUIRE(np_ > 0 && Q > 0, "nmgp_kl_rbar");
    dim3 g2((Q + 7) / 8, np_);
    k_kl_bwd_R<<<NMGP_L(g2), 256, 0, st>>>(nullptr, nullptr, R, G, Rbar, 0, Q);
    return nmgp_launch_status("nmgp_kl_rbar");
}

// ------------------------------------------------------------------------------------------
// v_s = mu_v + C_v z_s, ellz = exp(v)   (code/utils.py:225-227 + code/nmgp_dsvi.py:215)
__global__ void k_sample_v_fwd(const double* __restrict__ mu, const double* __restrict__ Cv, const double* __restrict__ zv,
                               double* __restrict__ v, double* __restrict__ ellz, int S, int Q) {
    int gid = blockIdx.x * blockDim.x + threadIdx.x;
    if (gid >= S * Q) return;
    int s = gid / Q, a = gid - s * Q;
    double acc = 0.0;
    for (int c = 0; c <= a; ++c) acc = fma(Cv[(size_t)a * Q + c], zv[(size_t)s * Q + c], acc);
    double val = mu[a] + acc;
    v[gid] = val;
    ellz[gid] = exp(val);
}
NMGP_API int nmgp_sample_v_fwd(const double* mu_v, const double* Cv, const double* zv, double* v, double* ellz, int S,
                               int Q, cudaStream_t st) {
    NMGP_REQUIRE(S > 0 && Q > 0, "nmgp_sample_v_fwd");
    k_sample_v_fwd<<<NMGP_L((S * Q + 127) / 128), 128, 0, st>>>(mu_v, Cv, zv, v, ellz, S, Q);
    return nmgp_launch_status("nmgp_sample_v_fwd");
}
// vb = vbar + ellzbar*ellz;  mu_v_bar[a] += sum_s vb[s,a];  Cvbar[a,c] += sum_s vb[s,a] z[s,c] (c<=a)
__global__ void k_sample_v_bwd(const double* __restrict__ ellzbar, const double* __restrict__ vbar,
                               const double* __restrict__ ellz, const double* __restrict__ zv,
                               double* __restrict__ mubar, double* __restrict__ Cvbar, int S, int Q) {
    int gid = blockIdx.x * blockDim.x + threadIdx.x;
    if (gid >= Q * Q) return;
    int a = gid / Q, c = gid - a * Q;
    if (c > a) return;
    double acc = 0.0, accm = 0.0;
    for (int s = 0; s < S; ++s) {
        double vb = vbar[(size_t)s * Q + a] + ellzbar[(size_t)s * Q + a] * ellz[(size_t)s * Q + a];
        acc = fma(vb, zv[(size_t)s * Q + c], acc);
        accm += vb;
    }
    Cvbar[gid] += acc;
    if (c == 0) mubar[a] += accm;
}
NMGP_API int nmgp_sample_v_bwd(const double* ellzbar, const double* vbar, const double* ellz, const double* zv,
                               double* mu_v_bar, double* Cvbar, int S, int Q, cudaStream_t st) {
    NMGP_REQUIRE(S > 0 && Q > 0, "nmgp_sample_v_bwd");
    k_sample_v_bwd<<<NMGP_L((Q * Q + 127) / 128), 128, 0, st>>>(ellzbar, vbar, ellz, zv, mu_v_bar, Cvbar, S, Q);
    return nmgp_launch_status("nmgp_sample_v_bwd");
}
