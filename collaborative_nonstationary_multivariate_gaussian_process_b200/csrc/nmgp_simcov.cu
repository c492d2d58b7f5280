// Covariance builds of the SIM_code (exact / Kronecker) line: code/SIM_code/Utility/kernels.py
//   Nonstationary_RBF_cov (kernels.py:46-73)  K[i,j] = s_i s_j sqrt(2 a_i b_j / (a_i^2 + b_j^2)) exp(-d_ij / (a_i^2 + b_j^2))
//   RBF_cov               (kernels.py:24-43)  K[i,j] = alpha^2 exp(-d(x_i/beta, y_j/beta) / 2)
// with d in the reference's GEMM form |x|^2 + |y|^2 - 2 x.y (kernels.py:14-21; may be slightly negative, not clamped)
// and + jitter on the diagonal of a self-covariance.
//
// The builds are FP64-ALU / HBM-store co-limited (8 bytes written per entry against one rsqrt + one exp per entry), so
//   * per-point factors are hoisted out of the pair loop: s_i sqrt(a_i), s_j sqrt(b_j), a_i^2, b_j^2, |x|^2 are
//     staged in shared memory per 64-point strip, and sqrt(2ab/A) exp(-d/A) becomes ONE rsqrt(A) (A^-1 = rsqrt^2) and
//     one exp per entry -- no division, no sqrt;
//   * a self-covariance only evaluates the tiles on and below the diagonal; each 64 x 64 tile goes through shared
//     memory once and is written twice, as itself and transposed, both with 16-byte coalesced stores;
//   * every thread holds its 4 row points in registers and sweeps 4 column points (16 entries per thread).
// The adjoint (w.r.t. the per-point sigma and ell; used by the differentiable log-posteriors, logpos.py:216-296)
// accumulates three row sums and three column sums per tile: sum w, sum w/A, sum w d/A^2 with w = Kbar o K.
#include "common.cuh"

namespace {

#define SC_T 64
#define SC_LD 66
#define SC_THREADS 256

struct Strip {           // per-point factors of a 64-point strip
    double x[SC_T], n[SC_T], a2[SC_T], f[SC_T];
};

// KIND 0: nonstationary; KIND 1: stationary RBF (x pre-divided by beta)
template <int KIND>
__device__ __forceinline__ void load_strip(Strip& s, const double* __restrict__ X, const double* __restrict__ sg,
                                           const double* __restrict__ ell, long long p0, long long T, double beta,
                                           bool row_side, int tid) {
    if (tid < SC_T) {
        const long long p = p0 + tid;
        double x = 0.0, a = 1.0, sig = 1.0;
        if (p < T) {
            x = X[p];
            if (KIND == 0) {
                if (ell) a = ell[p];
                if (sg) sig = sg[p];
            }
        }
        if (KIND == 1) x = x / beta;
        s.x[tid] = x;
        s.n[tid] = x * x;
        s.a2[tid] = a * a;
        s.f[tid] = sig * sqrt(a);          // same factor on both sides: K[i,j] and K[j,i] are then bitwise equal
        (void)row_side;
    }
}

template <int KIND>
__device__ __forceinline__ double entry(double xi, double ni, double a2i, double fi, double xj, double nj, double b2j,
                                        double fj, double alpha2) {
    const double dist = (ni + nj) - 2.0 * __dmul_rn(xi, xj);      // the reference rounds the product (torch.mm) first
    if (KIND == 1) return exp(-0.5 * dist) * alpha2;
    const double A = a2i + b2j;
    const double rs = rsqrt(A);
    return (fi * fj) * (rs * 1.4142135623730951) * exp(-dist * (rs * rs));
}

// SYM: blockIdx.x enumerates the tile pairs (ti >= tj) of a self-covariance; otherwise grid (tiles of T2, tiles of T1).
template <int KIND, bool SYM>
__global__ void __launch_bounds__(SC_THREADS, 3)
k_simcov(const double* __restrict__ X1, const double* __restrict__ sg1, const double* __restrict__ l1,
         const double* __restrict__ X2, const double* __restrict__ sg2, const double* __restrict__ l2, double alpha,
         double beta, double jitter, double* __restrict__ K, long long T1, long long T2) {
    extern __shared__ __align__(16) double smd[];
    double* tile = smd;                                          // [SC_T][SC_LD]
    double* tileT = smd + SC_T * SC_LD;                          // [SC_T][SC_LD]  (SYM only)
    Strip& rs_ = *reinterpret_cast<Strip*>(smd + (SYM ? 2 : 1) * SC_T * SC_LD);
    Strip& cs_ = *(&rs_ + 1);
    long long ti, tj;
    if (SYM) {
        const long long k = blockIdx.x;
        ti = (long long)((sqrt(8.0 * (double)k + 1.0) - 1.0) * 0.5);
        while (ti * (ti + 1) / 2 > k) --ti;
        while ((ti + 1) * (ti + 2) / 2 <= k) ++ti;
        tj = k - ti * (ti + 1) / 2;
    } else {
        ti = blockIdx.y;
        tj = blockIdx.x;
    }
    const long long i0 = ti * SC_T, j0 = tj * SC_T;
    const int tid = threadIdx.x, ty = tid >> 4, tx = tid & 15;
    load_strip<KIND>(rs_, X1, sg1, l1, i0, T1, beta, true, tid);
    if (tid >= SC_T && tid < 2 * SC_T) load_strip<KIND>(cs_, X2, sg2, l2, j0, T2, beta, false, tid - SC_T);
    __syncthreads();
    const double alpha2 = alpha * alpha;
    const bool diag_tile = SYM && (ti == tj);
    double cx[4], cn[4], cb[4], cf[4];
#pragma unroll
    for (int cc = 0; cc < 4; ++cc) {
        const int c = tx + 16 * cc;
        cx[cc] = cs_.x[c]; cn[cc] = cs_.n[c]; cb[cc] = cs_.a2[c]; cf[cc] = cs_.f[c];
    }
#pragma unroll
    for (int rr = 0; rr < 4; ++rr) {
        const int r = ty + 16 * rr;
        const double xi = rs_.x[r], ni = rs_.n[r], a2i = rs_.a2[r], fi = rs_.f[r];
#pragma unroll
        for (int cc = 0; cc < 4; ++cc) {
            const int c = tx + 16 * cc;
            double k = entry<KIND>(xi, ni, a2i, fi, cx[cc], cn[cc], cb[cc], cf[cc], alpha2);
            if ((SYM ? diag_tile : true) && (i0 + r == j0 + c)) k += jitter;
            tile[r * SC_LD + c] = k;
            if (SYM) tileT[c * SC_LD + r] = k;
        }
    }
    __syncthreads();
    const bool vec = (T2 & 1) == 0;
    const int nr = (int)min((long long)SC_T, T1 - i0), nc = (int)min((long long)SC_T, T2 - j0);
#pragma unroll
    for (int it = 0; it < (SC_T * SC_T / 2) / SC_THREADS; ++it) {
        const int idx = it * SC_THREADS + tid, row = idx >> 5, c2 = 2 * (idx & 31);
        if (row < nr && c2 < nc) {
            const double2 v = *reinterpret_cast<const double2*>(&tile[row * SC_LD + c2]);
            double* dst = K + (size_t)(i0 + row) * T2 + j0 + c2;
            if (vec && c2 + 1 < nc) *reinterpret_cast<double2*>(dst) = v;
            else {
                dst[0] = v.x;
                if (c2 + 1 < nc) dst[1] = v.y;
            }
        }
    }
    if (SYM && !diag_tile) {              // the mirrored tile: rows j0.., columns i0..   (T1 == T2)
#pragma unroll
        for (int it = 0; it < (SC_T * SC_T / 2) / SC_THREADS; ++it) {
            const int idx = it * SC_THREADS + tid, row = idx >> 5, c2 = 2 * (idx & 31);
            if (row < nc && c2 < nr) {
                const double2 v = *reinterpret_cast<const double2*>(&tileT[row * SC_LD + c2]);
                double* dst = K + (size_t)(j0 + row) * T2 + i0 + c2;
                if (vec && c2 + 1 < nr) *reinterpret_cast<double2*>(dst) = v;
                else {
                    dst[0] = v.x;
                    if (c2 + 1 < nr) dst[1] = v.y;
                }
            }
        }
    }
}

// generic input dimension (dx > 1): one thread per entry, 2-D tiled indexing (no grid.y limit on T1)
template <int KIND>
__global__ void k_simcov_generic(const double* __restrict__ X1, const double* __restrict__ sg1,
                                 const double* __restrict__ l1, const double* __restrict__ X2,
                                 const double* __restrict__ sg2, const double* __restrict__ l2, double alpha, double beta,
                                 double jitter, double* __restrict__ K, long long T1, long long T2, int dx, int add_jitter) {
    const long long j = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= T2) return;
    for (long long i = blockIdx.y; i < T1; i += gridDim.y) {
        double xn = 0.0, yn = 0.0, xy = 0.0;
        for (int k = 0; k < dx; ++k) {
            double a = X1[i * dx + k], b = X2[j * dx + k];
            if (KIND == 1) { a = a / beta; b = b / beta; }
            xn = fma(a, a, xn);
            yn = fma(b, b, yn);
            xy = fma(a, b, xy);
        }
        const double dist = xn + yn - 2.0 * xy;
        double k;
        if (KIND == 1) k = exp(-0.5 * dist) * (alpha * alpha);
        else {
            const double a = l1 ? l1[i] : 1.0, b = l2 ? l2[j] : 1.0;
            const double c = (sg1 ? sg1[i] : 1.0) * (sg2 ? sg2[j] : 1.0);
            const double A = a * a + b * b;
            k = c * sqrt(2.0 * (a * b) / A) * exp(-dist / A);
        }
        if (add_jitter && i == j) k += jitter;
        K[i * T2 + j] = k;
    }
}

// ---- adjoint of the nonstationary build w.r.t. sigma and ell (inputs 1-D, as everywhere in the reference: q11) ----
// per 64 x 64 tile: w = Kbar o K (the jitter carries no parameter);  row sums R0 = sum w, R1 = sum w/A, R2 = sum w d/A^2
// and the same column sums, then
//   g_sigma1[i] += R0 / s_i;        g_ell1[i] += R0/(2 a_i) - a_i R1 + 2 a_i R2       
//   g_sigma2[j] += C0 / s_j;        g_ell2[j] += C0/(2 b_j) - b_j C1 + 2 b_j C2
__global__ void __launch_bounds__(SC_THREADS, 3)
k_nonstat_cov_bwd(const double* __restrict__ X1, const double* __restrict__ sg1, const double* __restrict__ l1,
                  const double* __restrict__ X2, const double* __restrict__ sg2, const double* __restrict__ l2,
                  const double* __restrict__ Kbar, double* __restrict__ g_sg1, double* __restrict__ g_l1,
                  double* __restrict__ g_sg2, double* __restrict__ g_l2, long long T1, long long T2) {
    __shared__ Strip rs_, cs_;
    __shared__ double rsum[3][SC_T], csum[3][SC_T];
    const long long i0 = (long long)blockIdx.y * SC_T, j0 = (long long)blockIdx.x * SC_T;
    const int tid = threadIdx.x, ty = tid >> 4, tx = tid & 15;
    load_strip<0>(rs_, X1, sg1, l1, i0, T1, 1.0, true, tid);
    if (tid >= SC_T && tid < 2 * SC_T) load_strip<0>(cs_, X2, sg2, l2, j0, T2, 1.0, false, tid - SC_T);
    if (tid < SC_T)
        for (int k = 0; k < 3; ++k) rsum[k][tid] = csum[k][tid] = 0.0;
    __syncthreads();
    double cacc[4][3];
#pragma unroll
    for (int cc = 0; cc < 4; ++cc) cacc[cc][0] = cacc[cc][1] = cacc[cc][2] = 0.0;
#pragma unroll
    for (int rr = 0; rr < 4; ++rr) {
        const int r = ty + 16 * rr;
        const long long gi = i0 + r;
        const double xi = rs_.x[r], ni = rs_.n[r], a2i = rs_.a2[r], fi = rs_.f[r];
        double r0 = 0.0, r1 = 0.0, r2 = 0.0;
#pragma unroll
        for (int cc = 0; cc < 4; ++cc) {
            const int c = tx + 16 * cc;
            const long long gj = j0 + c;
            if (gi < T1 && gj < T2) {
                const double dist = (ni + cs_.n[c]) - 2.0 * __dmul_rn(xi, cs_.x[c]);
                const double A = a2i + cs_.a2[c];
                const double rsq = rsqrt(A), rA = rsq * rsq;
                const double k = (fi * cs_.f[c]) * (rsq * 1.4142135623730951) * exp(-dist * rA);
                const double w = Kbar[(size_t)gi * T2 + gj] * k;
                const double w1 = w * rA, w2 = w1 * dist * rA;
                r0 += w; r1 += w1; r2 += w2;
                cacc[cc][0] += w; cacc[cc][1] += w1; cacc[cc][2] += w2;
            }
        }
        // reduce over the 16 tx lanes sharing this row (lanes of a half warp)
#pragma unroll
        for (int o = 8; o > 0; o >>= 1) {
            r0 += __shfl_xor_sync(0xffffffffu, r0, o);
            r1 += __shfl_xor_sync(0xffffffffu, r1, o);
            r2 += __shfl_xor_sync(0xffffffffu, r2, o);
        }
        if (tx == 0) { rsum[0][r] = r0; rsum[1][r] = r1; rsum[2][r] = r2; }
    }
#pragma unroll
    for (int cc = 0; cc < 4; ++cc)
#pragma unroll
        for (int k = 0; k < 3; ++k) atomicAdd(&csum[k][tx + 16 * cc], cacc[cc][k]);
    __syncthreads();
    if (tid < SC_T) {
        const long long gi = i0 + tid;
        if (gi < T1) {
            const double a = l1 ? l1[gi] : 1.0, sgi = sg1 ? sg1[gi] : 1.0;
            if (g_sg1) atomicAdd(&g_sg1[gi], rsum[0][tid] / sgi);
            if (g_l1) atomicAdd(&g_l1[gi], rsum[0][tid] / (2.0 * a) - a * rsum[1][tid] + 2.0 * a * rsum[2][tid]);
        }
    } else if (tid < 2 * SC_T) {
        const int c = tid - SC_T;
        const long long gj = j0 + c;
        if (gj < T2) {
            const double b = l2 ? l2[gj] : 1.0, sgj = sg2 ? sg2[gj] : 1.0;
            if (g_sg2) atomicAdd(&g_sg2[gj], csum[0][c] / sgj);
            if (g_l2) atomicAdd(&g_l2[gj], csum[0][c] / (2.0 * b) - b * csum[1][c] + 2.0 * b * csum[2][c]);
        }
    }
}

// adjoint for a generic input dimension (dx > 1; no call site of the reference, which views x as (-1, 1) everywhere):
// one CTA per (row, 256 columns); row sums through a block reduction, column sums by one atomic per entry
__global__ void k_nonstat_cov_bwd_generic(const double* __restrict__ X1, const double* __restrict__ sg1,
                                          const double* __restrict__ l1, const double* __restrict__ X2,
                                          const double* __restrict__ sg2, const double* __restrict__ l2,
                                          const double* __restrict__ Kbar, double* __restrict__ g_sg1,
                                          double* __restrict__ g_l1, double* __restrict__ g_sg2, double* __restrict__ g_l2,
                                          long long T1, long long T2, int dx) {
    const long long j = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    for (long long i = blockIdx.y; i < T1; i += gridDim.y) {
        double w = 0.0, w1 = 0.0, w2 = 0.0;
        const double a = l1 ? l1[i] : 1.0, si = sg1 ? sg1[i] : 1.0;
        if (j < T2) {
            double xn = 0.0, yn = 0.0, xy = 0.0;
            for (int k = 0; k < dx; ++k) {
                const double u = X1[i * dx + k], v = X2[j * dx + k];
                xn = fma(u, u, xn);
                yn = fma(v, v, yn);
                xy = fma(u, v, xy);
            }
            const double dist = xn + yn - 2.0 * xy;
            const double b = l2 ? l2[j] : 1.0, sj = sg2 ? sg2[j] : 1.0;
            const double A = a * a + b * b;
            const double kv = si * sj * sqrt(2.0 * (a * b) / A) * exp(-dist / A);
            w = Kbar[i * T2 + j] * kv;
            w1 = w / A;
            w2 = w1 * dist / A;
            if (g_sg2) atomicAdd(&g_sg2[j], w / sj);
            if (g_l2) atomicAdd(&g_l2[j], w / (2.0 * b) - b * w1 + 2.0 * b * w2);
        }
        const double r0 = block_sum(w), r1 = block_sum(w1), r2 = block_sum(w2);
        if (threadIdx.x == 0) {
            if (g_sg1) atomicAdd(&g_sg1[i], r0 / si);
            if (g_l1) atomicAdd(&g_l1[i], r0 / (2.0 * a) - a * r1 + 2.0 * a * r2);
        }
    }
}

template <int KIND>
int launch_simcov(const double* X1, const double* sg1, const double* l1, const double* X2, const double* sg2,
                  const double* l2, double alpha, double beta, double jitter, double* K, long long T1, long long T2,
                  int dx, int self, cudaStream_t st) {
    if (dx != 1) {
        dim3 grid((unsigned)((T2 + 255) / 256), (unsigned)min(T1, 65535LL));
        k_simcov_generic<KIND><<<NMGP_L(grid), 256, 0, st>>>(X1, sg1, l1, X2, sg2, l2, alpha, beta, jitter, K, T1, T2, dx,
                                                            self);
        return 0;
    }
    const long long nti = (T1 + SC_T - 1) / SC_T, ntj = (T2 + SC_T - 1) / SC_T;
    const size_t strips = 2 * sizeof(Strip);
    if (self && T1 == T2) {
        const long long pairs = nti * (nti + 1) / 2;
        const size_t smem = 2 * SC_T * SC_LD * sizeof(double) + strips;
        if (int r = nmgp_opt_in_smem(k_simcov<KIND, true>, smem, "nmgp_simcov")) return r;
        k_simcov<KIND, true><<<NMGP_L((unsigned)pairs), SC_THREADS, smem, st>>>(X1, sg1, l1, X2, sg2, l2, alpha, beta,
                                                                                 jitter, K, T1, T2);
    } else {
        dim3 grid((unsigned)ntj, (unsigned)nti);
        const size_t smem = SC_T * SC_LD * sizeof(double) + strips;
        k_simcov<KIND, false><<<NMGP_L(grid), SC_THREADS, smem, st>>>(X1, sg1, l1, X2, sg2, l2, alpha, beta,
                                                                       self ? jitter : 0.0, K, T1, T2);
    }
    return 0;
}

}  // namespace

// `self` != 0: X2/sigma2/ell2 are X1/sigma1/ell1 (the reference's X2=None call): lower tiles mirrored, + jitter I.
NMGP_API int nmgp_nonstationary_cov(const double* X1, const double* sigma1, const double* ell1, const double* X2,
                                    const double* sigma2, const double* ell2, double jitter, double* K, long long T1,
                                    long long T2, int dx, int self, cudaStream_t st) {
    NMGP_REQUIRE(T1 >= 0 && T2 >= 0 && dx > 0 && T1 < (1LL << 31) && T2 < (1LL << 31), "nmgp_nonstationary_cov");
    if (T1 == 0 || T2 == 0) return 0;
    if (int r = launch_simcov<0>(X1, sigma1, ell1, X2, sigma2, ell2, 1.0, 1.0, jitter, K, T1, T2, dx, self, st)) return r;
    return nmgp_launch_status("nmgp_nonstationary_cov");
}

NMGP_API int nmgp_sim_rbf_cov(const double* X1, const double* X2, double alpha, double beta, double jitter, double* K,
                              long long T1, long long T2, int dx, int self, cudaStream_t st) {
    NMGP_REQUIRE(T1 >= 0 && T2 >= 0 && dx > 0 && T1 < (1LL << 31) && T2 < (1LL << 31), "nmgp_sim_rbf_cov");
    if (T1 == 0 || T2 == 0) return 0;
    if (int r = launch_simcov<1>(X1, nullptr, nullptr, X2, nullptr, nullptr, alpha, beta, jitter, K, T1, T2, dx, self, st))
        return r;
    return nmgp_launch_status("nmgp_sim_rbf_cov");
}

// g_* (+=, any of them may be NULL): adjoint of nmgp_nonstationary_cov w.r.t. the per-point sigma / ell (tiled kernel for
// dx == 1, the reference's case; a plain one for dx > 1).
// For a self-covariance the caller adds the row-side and column-side results.
NMGP_API int nmgp_nonstationary_cov_bwd(const double* X1, const double* sigma1, const double* ell1, const double* X2,
                                        const double* sigma2, const double* ell2, const double* Kbar, double* g_sigma1,
                                        double* g_ell1, double* g_sigma2, double* g_ell2, long long T1, long long T2,
                                        int dx, cudaStream_t st) {
    NMGP_REQUIRE(T1 >= 0 && T2 >= 0 && dx >= 1 && T1 < (1LL << 22) && T2 < (1LL << 31), "nmgp_nonstationary_cov_bwd");
    if (T1 == 0 || T2 == 0) return 0;
    if (dx != 1) {
        dim3 gg((unsigned)((T2 + 255) / 256), (unsigned)min(T1, 65535LL));
        k_nonstat_cov_bwd_generic<<<NMGP_L(gg), 256, 0, st>>>(X1, sigma1, ell1, X2, sigma2, ell2, Kbar, g_sigma1, g_ell1,
                                                               g_sigma2, g_ell2, T1, T2, dx);
        return nmgp_launch_status("nmgp_nonstationary_cov_bwd");
    }
    dim3 grid((unsigned)((T2 + SC_T - 1) / SC_T), (unsigned)((T1 + SC_T - 1) / SC_T));
    k_nonstat_cov_bwd<<<NMGP_L(grid), SC_THREADS, 0, st>>>(X1, sigma1, ell1, X2, sigma2, ell2, Kbar, g_sigma1, g_ell1,
                                                            g_sigma2, g_ell2, T1, T2);
    return nmgp_launch_status("nmgp_nonstationary_cov_bwd");
}
