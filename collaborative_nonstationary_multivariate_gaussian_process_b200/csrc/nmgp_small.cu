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
NMGP_API int nmgp_potrf_batched(const double* A, double jitter, double* C, double* hld, int* info, int nb, int Q,
                                cudaStream_t st) {
    NMGP_REQUIRE(nb >= 0 && Q > 0 && Q <= 128, "nmgp_potrf_batched");
    if (nb == 0) return 0;
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

// ------------------------------------------------------------------------------------------
// KL(N(mu_b, C_b C_b^T) || N(0, R_p R_p^T)) in the reference's form (code/utils.py:346-351, quirk q10):
//   kl[p,b] = hldR[p] - hldS[b] + 0.5 ( sum_{a,c} (C_b[a,c]/R_p[a,a])^2 + ||R_p^-1 mu_b||^2 - Q )
// grid (ceil(nb/128), np); thread = one (p,b); R_p in shared memory; t = R_p^-1 mu_b kept in global (saved for bwd).
__global__ void k_kl_fwd(const double* __restrict__ CS, const double* __restrict__ hldS, const double* __restrict__ mu,
                         const double* __restrict__ R, const double* __restrict__ hldR, double* __restrict__ kl,
                         double* __restrict__ t, int np_, int nb, int Q) {
    extern __shared__ double Rs[];
    const int p = blockIdx.y;
    const double* Rp = R + (size_t)p * Q * Q;
    for (int e = threadIdx.x; e < Q * Q; e += blockDim.x) Rs[e] = Rp[e];
    __syncthreads();
    int b = blockIdx.x * blockDim.x + threadIdx.x;
    if (b >= nb) return;
    double* tb = t + ((size_t)p * nb + b) * Q;
    const double* mub = mu + (size_t)b * Q;
    const double* Cb = CS + (size_t)b * Q * Q;
    double term3 = 0.0, term2 = 0.0;
    for (int a = 0; a < Q; ++a) {
        double s = mub[a];
        for (int c = 0; c < a; ++c) s = fma(-Rs[a * Q + c], tb[c], s);
        double d = Rs[a * Q + a];
        s /= d;
        tb[a] = s;
        term3 = fma(s, s, term3);
        double rs = 0.0;
        for (int c = 0; c <= a; ++c) {
            double v = Cb[a * Q + c];
            rs = fma(v, v, rs);
        }
        term2 += rs / (d * d);
    }
    kl[(size_t)p * nb + b] = hldR[p] - hldS[b] + 0.5 * (term2 + term3 - (double)Q);
}
NMGP_API int nmgp_kl_fwd(const double* CS, const double* hldS, const double* mu, const double* R, const double* hldR,
                         double* kl, double* t, int np_, int nb, int Q, cudaStream_t st) {
    NMGP_REQUIRE(np_ > 0 && nb > 0 && Q > 0, "nmgp_kl_fwd");
    size_t smem = (size_t)Q * Q * sizeof(double);
    if (int r = nmgp_opt_in_smem(k_kl_fwd, smem, "nmgp_kl_fwd")) return r;
    dim3 grid((nb + 127) / 128, np_);
    k_kl_fwd<<<NMGP_L(grid), 128, smem, st>>>(CS, hldS, mu, R, hldR, kl, t, np_, nb, Q);
    return nmgp_launch_status("nmgp_kl_fwd");
}

// backward part 1: work[p,b,:] = R_p^-T (klbar[p,b] * t[p,b,:])
__global__ void k_kl_bwd_solve(const double* __restrict__ klbar, const double* __restrict__ R,
                               const double* __restrict__ t, double* __restrict__ work, int np_, int nb, int Q) {
    extern __shared__ double Rs[];
    const int p = blockIdx.y;
    const double* Rp = R + (size_t)p * Q * Q;
    for (int e = threadIdx.x; e < Q * Q; e += blockDim.x) Rs[e] = Rp[e];
    __syncthreads();
    int b = blockIdx.x * blockDim.x + threadIdx.x;
    if (b >= nb) return;
    const double kb = klbar[(size_t)p * nb + b];
    const double* tb = t + ((size_t)p * nb + b) * Q;
    double* wb = work + ((size_t)p * nb + b) * Q;
    for (int a = Q - 1; a >= 0; --a) {
        double s = kb * tb[a];
        for (int c = a + 1; c < Q; ++c) s = fma(-Rs[c * Q + a], wb[c], s);
        wb[a] = s / Rs[a * Q + a];
    }
}
// part 2: Rbar[p][a,c] (a>=c) -= sum_b work[p,b,a] t[p,b,c];  diagonal += -2/d^3 * 0.5 sum_b klbar[p,b] rs[b,a]
// grid (nchunks, np); each block handles a chunk of b and atomically adds its partial.
__global__ void k_kl_bwd_R(const double* __restrict__ klbar, const double* __restrict__ CS, const double* __restrict__ R,
                           const double* __restrict__ t, const double* __restrict__ work, double* __restrict__ Rbar,
                           int np_, int nb, int Q, int chunk) {
    const int p = blockIdx.y;
    const int b0 = blockIdx.x * chunk, b1 = min(nb, b0 + chunk);
    for (int e = threadIdx.x; e < Q * Q; e += blockDim.x) {
        int a = e / Q, c = e - a * Q;
        if (c > a) continue;
        double s = 0.0;
        for (int b = b0; b < b1; ++b) {
            size_t o = ((size_t)p * nb + b) * Q;
            s = fma(-work[o + a], t[o + c], s);
        }
        if (a == c) {
            double d = R[(size_t)p * Q * Q + (size_t)a * Q + a];
            double wbar = 0.0;
            for (int b = b0; b < b1; ++b) {
                const double* Cb = CS + (size_t)b * Q * Q + (size_t)a * Q;
                double rs = 0.0;
                for (int k = 0; k <= a; ++k) rs = fma(Cb[k], Cb[k], rs);
                wbar = fma(0.5 * klbar[(size_t)p * nb + b], rs, wbar);
            }
            s += wbar * (-2.0) / (d * d * d);
        }
        atomicAdd(&Rbar[(size_t)p * Q * Q + e], s);
    }
}
// part 3: mubar, CSbar, hldSbar (thread per (b, a)); hldRbar by block 0
__global__ void k_kl_bwd_S(const double* __restrict__ klbar, const double* __restrict__ CS, const double* __restrict__ R,
                           const double* __restrict__ work, double* __restrict__ CSbar, double* __restrict__ hldSbar,
                           double* __restrict__ mubar, int np_, int nb, int Q) {
    long long gid = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (gid >= (long long)nb * Q) return;
    int b = (int)(gid / Q), a = (int)(gid - (long long)b * Q);
    double mb = 0.0, rsbar = 0.0, ks = 0.0;
    for (int p = 0; p < np_; ++p) {
        double kb = klbar[(size_t)p * nb + b];
        mb += work[((size_t)p * nb + b) * Q + a];
        double d = R[(size_t)p * Q * Q + (size_t)a * Q + a];
        rsbar = fma(0.5 * kb, 1.0 / (d * d), rsbar);
        ks += kb;
    }
    mubar[(size_t)b * Q + a] = mb;
    if (a == 0) hldSbar[b] = -ks;
    const double* Cb = CS + (size_t)b * Q * Q + (size_t)a * Q;
    double* Ob = CSbar + (size_t)b * Q * Q + (size_t)a * Q;
    for (int c = 0; c < Q; ++c) Ob[c] = (c <= a) ? 2.0 * rsbar * Cb[c] : 0.0;
}
__global__ void k_kl_bwd_hldR(const double* __restrict__ klbar, double* __restrict__ hldRbar, int np_, int nb) {
    int p = blockIdx.x;
    double s = 0.0;
    for (int b = threadIdx.x; b < nb; b += blockDim.x) s += klbar[(size_t)p * nb + b];
    s = block_sum(s);
    if (threadIdx.x == 0) hldRbar[p] = s;
}
NMGP_API int nmgp_kl_bwd(const double* klbar, const double* CS, const double* mu, const double* R, const double* t,
                         double* CSbar, double* hldSbar, double* mubar, double* Rbar /* pre-zeroed */, double* hldRbar,
                         double* work, int np_, int nb, int Q, cudaStream_t st) {
    (void)mu;
    NMGP_REQUIRE(np_ > 0 && nb > 0 && Q > 0, "nmgp_kl_bwd");
    size_t smem = (size_t)Q * Q * sizeof(double);
    if (int r = nmgp_opt_in_smem(k_kl_bwd_solve, smem, "nmgp_kl_bwd")) return r;
    dim3 g1((nb + 127) / 128, np_);
    k_kl_bwd_solve<<<NMGP_L(g1), 128, smem, st>>>(klbar, R, t, work, np_, nb, Q);
    const int chunk = 32;
    dim3 g2((nb + chunk - 1) / chunk, np_);
    k_kl_bwd_R<<<NMGP_L(g2), 256, 0, st>>>(klbar, CS, R, t, work, Rbar, np_, nb, Q, chunk);
    long long n3 = (long long)nb * Q;
    k_kl_bwd_S<<<NMGP_L((unsigned)((n3 + 127) / 128)), 128, 0, st>>>(klbar, CS, R, work, CSbar, hldSbar, mubar, np_, nb, Q);
    k_kl_bwd_hldR<<<NMGP_L(np_), 128, 0, st>>>(klbar, hldRbar, np_, nb);
    return nmgp_launch_status("nmgp_kl_bwd");
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
