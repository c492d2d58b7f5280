// Covariance builds: stationary RBF (code/utils.py:75-94), Gibbs/Paciorek nonstationary RBF
// (code/utils.py:97-103) with adjoints, and the SIM_code builds (code/SIM_code/Utility/kernels.py:5-73).
// Bound by HBM stores and the FP64 exp/sqrt/div ALU sequences; one element per thread, coalesced
// along the inducing/column index.
#include "common.cuh"

// K[n,q] = s2 * exp(-0.5 (x_n/len - z_q/len)^2) (+ jitter on n==q)
__global__ void k_rbf_fwd(const double* __restrict__ x, const double* __restrict__ z, const double* __restrict__ hyp,
                          int is2, int ilen, double jitter, double* __restrict__ K, long long B, int Q) {
    long long gid = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (gid >= B * Q) return;
    long long n = gid / Q;
    int q = (int)(gid - n * Q);
    const double s2 = hyp[is2], len = hyp[ilen];
    double r = x[n] / len - z[q] / len;
    double k = s2 * exp(-0.5 * (r * r));
    if (jitter != 0.0 && n == q) k += jitter;
    K[gid] = k;
}
NMGP_API int nmgp_rbf_build_fwd(const double* x, const double* z, const double* hyp, int is2, int ilen, double jitter,
                                double* K, long long B, int Q, cudaStream_t st) {
    NMGP_REQUIRE(B >= 0 && Q > 0 && is2 >= 0 && is2 < H_COUNT && ilen >= 0 && ilen < H_COUNT, "nmgp_rbf_build_fwd");
    if (B == 0) return 0;
    long long n = B * Q;
    k_rbf_fwd<<<NMGP_L((unsigned)((n + 255) / 256)), 256, 0, st>>>(x, z, hyp, is2, ilen, jitter, K, B, Q);
    return nmgp_launch_status("nmgp_rbf_build_fwd");
}

// ghyp[is2] += sum Kbar*K (d/dlog s2);  ghyp[ilen] += sum Kbar*K*r^2 (d/dlog len)
__global__ void k_rbf_bwd(const double* __restrict__ x, const double* __restrict__ z, const double* __restrict__ hyp,
                          int is2, int ilen, const double* __restrict__ Kbar, double* __restrict__ ghyp, long long B,
                          int Q) {
    const double s2 = hyp[is2], len = hyp[ilen];
    double a0 = 0.0, a1 = 0.0;
    const long long total = B * Q;
    for (long long gid = (long long)blockIdx.x * blockDim.x + threadIdx.x; gid < total;
         gid += (long long)gridDim.x * blockDim.x) {
        long long n = gid / Q;
        int q = (int)(gid - n * Q);
        double r = x[n] / len - z[q] / len;
        double r2 = r * r;
        double g = Kbar[gid] * (s2 * exp(-0.5 * r2));
        a0 += g;
        a1 = fma(g, r2, a1);
    }
    a0 = block_sum(a0);
    a1 = block_sum(a1);
    if (threadIdx.x == 0) {
        atomicAdd(&ghyp[is2], a0);
        atomicAdd(&ghyp[ilen], a1);
    }
}
NMGP_API int nmgp_rbf_build_bwd(const double* x, const double* z, const double* hyp, int is2, int ilen,
                                const double* Kbar, double* ghyp, long long B, int Q, cudaStream_t st) {
    NMGP_REQUIRE(B >= 0 && Q > 0 && is2 >= 0 && is2 < H_COUNT && ilen >= 0 && ilen < H_COUNT, "nmgp_rbf_build_bwd");
    if (B == 0) return 0;
    long long n = B * Q;
    long long blocks = (n + 255) / 256;
    if (blocks > 148 * 8) blocks = 148 * 8;
    k_rbf_bwd<<<NMGP_L((unsigned)blocks), 256, 0, st>>>(x, z, hyp, is2, ilen, Kbar, ghyp, B, Q);
    return nmgp_launch_status("nmgp_rbf_build_bwd");
}

// ---- Gibbs kernel -----------------------------------------------------------------------------
__device__ __forceinline__ double gibbs_value(double xn, double zq, double a, double b) {
    double d = xn - zq;
    double r2 = d * d;
    double den = a * a + b * b;
    return sqrt(2.0 * (a * b) / den) * exp(-r2 / den);
}

// one warp per GB_ROWS_PER_WARP rows, lanes over q (Q <= 128): the per-column terms (z, b, b^2) are loaded once per
// warp and every entry costs one reciprocal, one sqrt and one exp
#define GB_ROWS_PER_WARP 8
template <int NU>
__global__ void __launch_bounds__(256, 4) k_gibbs_fwd(const double* __restrict__ x, const double* __restrict__ z, const double* __restrict__ ellx,
                            const double* __restrict__ ellz, double jitter, double* __restrict__ K, long long B, int Q) {
    const int s = blockIdx.y;
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5, nw = blockDim.x >> 5;
    const long long row0 = ((long long)blockIdx.x * nw + w) * GB_ROWS_PER_WARP;
    if (Q > 128) {   // wide case (not on the DSVI path): same arithmetic, column terms re-read per entry
        for (int rr = 0; rr < GB_ROWS_PER_WARP; ++rr) {
            const long long n = row0 + rr;
            if (n >= B) break;
            const double a = ellx[(size_t)s * B + n], xn = x[n];
            for (int q = lane; q < Q; q += 32) {
                const double d = xn - z[q], b = ellz[(size_t)s * Q + q];
                const double rden = 1.0 / fma(a, a, b * b);
                double k = sqrt(2.0 * (a * b) * rden) * exp(-(d * d) * rden);
                if (jitter != 0.0 && n == q) k += jitter;
                K[((size_t)s * B + n) * Q + q] = k;
            }
        }
        return;
    }
    double zq[NU], bq[NU];
#pragma unroll
    for (int u = 0; u < NU; ++u) {
        const int q = lane + 32 * u;
        zq[u] = q < Q ? z[q] : 0.0;
        bq[u] = q < Q ? ellz[(size_t)s * Q + q] : 1.0;
    }
    for (int rr = 0; rr < GB_ROWS_PER_WARP; ++rr) {
        const long long n = row0 + rr;
        if (n >= B) break;
        const double a = ellx[(size_t)s * B + n], xn = x[n];
        double* krow = K + ((size_t)s * B + n) * Q;
#pragma unroll
        for (int u = 0; u < NU; ++u) {
            const int q = lane + 32 * u;
            if (q < Q) {
                const double d = xn - zq[u], b = bq[u];
                const double rden = 1.0 / fma(a, a, b * b);
                double k = sqrt(2.0 * (a * b) * rden) * exp(-(d * d) * rden);
                if (jitter != 0.0 && n == q) k += jitter;
                krow[q] = k;
            }
        }
    }
}
NMGP_API int nmgp_gibbs_build_fwd(const double* x, const double* z, const double* ellx, const double* ellz,
                                  double jitter, double* K, int ns, long long B, int Q, cudaStream_t st) {
    NMGP_REQUIRE(ns >= 0 && ns <= 65535 && B >= 0 && Q > 0, "nmgp_gibbs_build_fwd");
    if (B == 0 || ns == 0) return 0;
    const int threads = 256, rows_per_block = (threads / 32) * GB_ROWS_PER_WARP;
    dim3 grid((unsigned)((B + rows_per_block - 1) / rows_per_block), ns);
    switch (Q > 128 ? 4 : (Q + 31) / 32) {
        case 1: k_gibbs_fwd<1><<<NMGP_L(grid), threads, 0, st>>>(x, z, ellx, ellz, jitter, K, B, Q); break;
        case 2: k_gibbs_fwd<2><<<NMGP_L(grid), threads, 0, st>>>(x, z, ellx, ellz, jitter, K, B, Q); break;
        case 3: k_gibbs_fwd<3><<<NMGP_L(grid), threads, 0, st>>>(x, z, ellx, ellz, jitter, K, B, Q); break;
        default: k_gibbs_fwd<4><<<NMGP_L(grid), threads, 0, st>>>(x, z, ellx, ellz, jitter, K, B, Q); break;
    }
    return nmgp_launch_status("nmgp_gibbs_build_fwd");
}

// ellxbar[s,n] = sum_q Kbar k dlog k/da ;  ellzbar[s,q] += sum_n Kbar k dlog k/db
// dlog k/da = 1/(2a) - a/den + 2 a r2/den^2.  One warp per row, lanes over q.
template <int NU>
__global__ void __launch_bounds__(256, 4) k_gibbs_bwd(const double* __restrict__ x, const double* __restrict__ z, const double* __restrict__ ellx,
                            const double* __restrict__ ellz, const double* __restrict__ Kbar,
                            const double* __restrict__ Kfwd, double* __restrict__ ellxbar, double* __restrict__ ellzbar,
                            long long B, int Q) {
    extern __shared__ double colsum[];  // [Q]
    const int s = blockIdx.y;
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5, nw = blockDim.x >> 5;
    for (int q = threadIdx.x; q < Q; q += blockDim.x) colsum[q] = 0.0;
    __syncthreads();
    const long long row0 = ((long long)blockIdx.x * nw + w) * GB_ROWS_PER_WARP;
    double cacc[NU];   // Q <= 32 NU
#pragma unroll
    for (int u = 0; u < NU; ++u) cacc[u] = 0.0;
    double zq[NU], bq[NU], hb[NU];
#pragma unroll
    for (int u = 0; u < NU; ++u) {
        const int q = lane + 32 * u;
        zq[u] = q < Q ? z[q] : 0.0;
        bq[u] = q < Q ? ellz[(size_t)s * Q + q] : 1.0;
        hb[u] = 0.5 / bq[u];
    }
    for (int rr = 0; rr < GB_ROWS_PER_WARP; ++rr) {
        long long n = row0 + rr;
        if (n >= B) break;
        const double a = ellx[(size_t)s * B + n], xn = x[n], ha = 0.5 / a;
        const double* kb = Kbar + ((size_t)s * B + n) * Q;
        const double* kf = Kfwd ? Kfwd + ((size_t)s * B + n) * Q : nullptr;   // forward values (no jitter) if kept
        double racc = 0.0;
#pragma unroll
        for (int u = 0; u < NU; ++u) {
            int q = lane + 32 * u;
            if (q < Q) {
                const double b = bq[u], d = xn - zq[u];
                const double r2 = d * d, rden = 1.0 / fma(a, a, b * b);
                const double k = kf ? kf[q] : sqrt(2.0 * (a * b) * rden) * exp(-r2 * rden);
                const double g = kb[q] * k;
                const double common = (2.0 * r2 * rden - 1.0) * rden;
                racc = fma(g, fma(a, common, ha), racc);
                cacc[u] = fma(g, fma(b, common, hb[u]), cacc[u]);
            }
        }
        racc = warp_sum(racc);
        if (lane == 0) ellxbar[(size_t)s * B + n] = racc;
    }
#pragma unroll
    for (int u = 0; u < NU; ++u) {
        int q = lane + 32 * u;
        if (q < Q && cacc[u] != 0.0) atomicAdd(&colsum[q], cacc[u]);
    }
    __syncthreads();
    for (int q = threadIdx.x; q < Q; q += blockDim.x)
        if (colsum[q] != 0.0) atomicAdd(&ellzbar[(size_t)s * Q + q], colsum[q]);
}
NMGP_API int nmgp_gibbs_build_bwd(const double* x, const double* z, const double* ellx, const double* ellz,
                                  const double* Kbar, const double* Kfwd, double* ellxbar, double* ellzbar, int ns,
                                  long long B, int Q, cudaStream_t st) {
    NMGP_REQUIRE(ns >= 0 && ns <= 65535 && B >= 0 && Q > 0 && Q <= 128, "nmgp_gibbs_build_bwd");
    if (B == 0 || ns == 0) return 0;
    const int threads = 256, rows_per_block = (threads / 32) * GB_ROWS_PER_WARP;
    dim3 grid((unsigned)((B + rows_per_block - 1) / rows_per_block), ns);
    switch ((Q + 31) / 32) {
        case 1: k_gibbs_bwd<1><<<NMGP_L(grid), threads, Q * sizeof(double), st>>>(x, z, ellx, ellz, Kbar, Kfwd, ellxbar, ellzbar, B, Q); break;
        case 2: k_gibbs_bwd<2><<<NMGP_L(grid), threads, Q * sizeof(double), st>>>(x, z, ellx, ellz, Kbar, Kfwd, ellxbar, ellzbar, B, Q); break;
        case 3: k_gibbs_bwd<3><<<NMGP_L(grid), threads, Q * sizeof(double), st>>>(x, z, ellx, ellz, Kbar, Kfwd, ellxbar, ellzbar, B, Q); break;
        default: k_gibbs_bwd<4><<<NMGP_L(grid), threads, Q * sizeof(double), st>>>(x, z, ellx, ellz, Kbar, Kfwd, ellxbar, ellzbar, B, Q); break;
    }
    return nmgp_launch_status("nmgp_gibbs_build_bwd");
}

// ---- SIM_code builds (kernels.py): nmgp_simcov.cu ---------------------------------------------------

// Hadamard (irregular-observation) covariance of the SIM_code line: out[i,j] = Kx[i,j] * Bf[indx1[i], indx2[j]]
// (+ diag on i == j)   -- logpos.generate_K_index (logpos.py:87-98) fused with the elementwise product K_x * K_i
// (prediction.py:746-748) and the sigma2_err I of the observation covariance.
__global__ void k_hadamard_index_cov(const double* __restrict__ Kx, const double* __restrict__ Bf,
                                     const int* __restrict__ indx1, const int* __restrict__ indx2, double diag,
                                     double* __restrict__ out, long long N1, long long N2, int M) {
    const long long j = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    const long long i = blockIdx.y + (long long)blockIdx.z * 65535;
    if (i >= N1 || j >= N2) return;
    double v = Kx[i * N2 + j] * Bf[(size_t)indx1[i] * M + indx2[j]];
    if (i == j) v += diag;
    out[i * N2 + j] = v;
}
NMGP_API int nmgp_hadamard_index_cov(const double* Kx, const double* Bf, const int* indx1, const int* indx2, double diag,
                                     double* out, long long N1, long long N2, int M, cudaStream_t st) {
    NMGP_REQUIRE(N1 >= 0 && N2 >= 0 && M > 0, "nmgp_hadamard_index_cov");
    if (N1 == 0 || N2 == 0) return 0;
    dim3 grid((unsigned)((N2 + 255) / 256), (unsigned)min(N1, 65535LL), (unsigned)((N1 + 65534) / 65535));
    k_hadamard_index_cov<<<NMGP_L(grid), 256, 0, st>>>(Kx, Bf, indx1, indx2, diag, out, N1, N2, M);
    return nmgp_launch_status("nmgp_hadamard_index_cov");
}

// Adjoint of the dense indexed log-likelihood  -1/2 logdet S - 1/2 y^T S^-1 y,  S = A o Bt[i1, i2] + sigma2 I
// (logpos.py:521-526 / 350-352 / 611-616 under autograd): with G = dloglik/dS = 1/2 (alpha alpha^T - S^-1), alpha = S^-1 y,
//     Abar[a,b] = g G[a,b] Bt[i1[a], i2[b]],   Btbar[i1[a], i2[b]] += g G[a,b] A[a,b],   s2bar += g tr G.
// G is formed on the fly from Sinv and alpha.  The table cotangent is accumulated in shared memory per CTA when the table
// is small (M*M <= HB_SMEM doubles: the M x M coregionalisation matrix), else straight with global atomics (the N x N
// time kernel indexed by the identity).
#define HB_SMEM 1024
__global__ void k_dense_loglik_bwd(const double* __restrict__ Sinv, const double* __restrict__ alpha,
                                   const double* __restrict__ A, const double* __restrict__ Bt,
                                   const int* __restrict__ i1, const int* __restrict__ i2, const double* __restrict__ gdev,
                                   double* __restrict__ Abar, double* __restrict__ Btbar, double* __restrict__ s2bar,
                                   long long N, int M, int Mrows) {
    __shared__ double tab[HB_SMEM];
    const bool priv = (long long)M * Mrows <= HB_SMEM;
    if (priv) {
        for (int e = threadIdx.x; e < M * Mrows; e += blockDim.x) tab[e] = 0.0;
        __syncthreads();
    }
    const double g = gdev[0];
    const long long a = blockIdx.y + (long long)blockIdx.z * 65535;
    double tr = 0.0;
    if (a < N) {
        const double al_a = alpha[a];
        const int ra = i1[a];
        for (long long b = (long long)blockIdx.x * blockDim.x + threadIdx.x; b < N; b += (long long)gridDim.x * blockDim.x) {
            const double G = 0.5 * g * (al_a * alpha[b] - Sinv[a * N + b]);
            const size_t slot = (size_t)ra * M + i2[b];
            Abar[a * N + b] = G * Bt[slot];
            const double c = G * A[a * N + b];
            if (priv) atomicAdd(&tab[slot], c);
            else atomicAdd(&Btbar[slot], c);
            if (a == b) tr += G;
        }
    }
    if (priv) {
        __syncthreads();
        for (int e = threadIdx.x; e < M * Mrows; e += blockDim.x)
            if (tab[e] != 0.0) atomicAdd(&Btbar[e], tab[e]);
    }
    if (tr != 0.0) atomicAdd(s2bar, tr);
}
NMGP_API int nmgp_dense_loglik_bwd(const double* Sinv, const double* alpha, const double* A, const double* Bt,
                                   const int* i1, const int* i2, const double* g, double* Abar, double* Btbar,
                                   double* s2bar, long long N, int Mrows, int M, cudaStream_t st) {
    NMGP_REQUIRE(N >= 0 && M > 0 && Mrows > 0, "nmgp_dense_loglik_bwd");
    if (N == 0) return 0;
    dim3 grid((unsigned)min((N + 255) / 256, 64LL), (unsigned)min(N, 65535LL), (unsigned)((N + 65534) / 65535));
    k_dense_loglik_bwd<<<NMGP_L(grid), 256, 0, st>>>(Sinv, alpha, A, Bt, i1, i2, g, Abar, Btbar, s2bar, N, M, Mrows);
    return nmgp_launch_status("nmgp_dense_loglik_bwd");
}
