// Row-wise kernels of the DSVI step: P = K12 (K22+eps I)^-1 through the Cholesky factor (replaces the
// LU solves of code/utils.py:119,142,154,230) with its adjoint, the tilde-ell / coefficient sampling
// (code/utils.py:32,120-124,231-235; code/nmgp_dsvi.py:215-238) and the expected log-likelihood
// (code/nmgp_dsvi.py:255-258) with all cotangents.
#include "common.cuh"
#include "philox.cuh"

// ------------------------------------------------------------------------------------------------
// solve_rows: per row k (Q), p = A^-1 k with A = R R^T: forward then backward substitution.
// Tile of TR rows per CTA; each thread owns one row, stored as a column of a [Q][TR+1] shared tile
// (conflict-free); R is read as a broadcast.
template <bool BWD>
__global__ void k_solve_rows(const double* __restrict__ K, const double* __restrict__ R, double* __restrict__ P,
                             double* __restrict__ c,
                             // backward extras
                             const double* __restrict__ Pbar, const double* __restrict__ cbar,
                             const double* __restrict__ Pin, double* __restrict__ Kbar, double* __restrict__ Abar,
                             long long B, int Q) {
    extern __shared__ double sm[];
    const int TR = blockDim.x, ldt = TR + 1;
    double* Rs = sm;                            // [Q*Q]
    double* Ys = Rs + (size_t)Q * Q;            // [Q][ldt]
    double* Ps = Ys + (size_t)Q * ldt;          // [Q][ldt] (BWD only)
    const int s = blockIdx.y;
    const long long row0 = (long long)blockIdx.x * TR;
    const int nrows = (int)min((long long)TR, B - row0);
    const double* Rg = R + (size_t)s * Q * Q;
    for (int e = threadIdx.x; e < Q * Q; e += blockDim.x) Rs[e] = Rg[e];
    const size_t base = ((size_t)s * B + row0) * Q;
    for (int e = threadIdx.x; e < nrows * Q; e += blockDim.x) {
        int r = e / Q, a = e - r * Q;
        double val;
        if (BWD) {
            val = fma(cbar[(size_t)s * B + row0 + r], K[base + e], Pbar[base + e]);
            Ps[a * ldt + r] = Pin[base + e];
        } else {
            val = K[base + e];
        }
        Ys[a * ldt + r] = val;
    }
    __syncthreads();
    const int r = threadIdx.x;
    if (r < nrows) {
        for (int a = 0; a < Q; ++a) {           // forward: R y = k
            double s0 = Ys[a * ldt + r], s1 = 0.0;
            int cidx = 0;
            for (; cidx + 1 < a; cidx += 2) {
                s0 = fma(-Rs[a * Q + cidx], Ys[cidx * ldt + r], s0);
                s1 = fma(-Rs[a * Q + cidx + 1], Ys[(cidx + 1) * ldt + r], s1);
            }
            if (cidx < a) s0 = fma(-Rs[a * Q + cidx], Ys[cidx * ldt + r], s0);
            Ys[a * ldt + r] = (s0 + s1) / Rs[a * Q + a];
        }
        for (int a = Q - 1; a >= 0; --a) {      // backward: R^T p = y
            double s0 = Ys[a * ldt + r], s1 = 0.0;
            int cidx = a + 1;
            for (; cidx + 1 < Q; cidx += 2) {
                s0 = fma(-Rs[cidx * Q + a], Ys[cidx * ldt + r], s0);
                s1 = fma(-Rs[(cidx + 1) * Q + a], Ys[(cidx + 1) * ldt + r], s1);
            }
            if (cidx < Q) s0 = fma(-Rs[cidx * Q + a], Ys[cidx * ldt + r], s0);
            Ys[a * ldt + r] = (s0 + s1) / Rs[a * Q + a];
        }
        if (!BWD) {
            const double* kr = K + base + (size_t)r * Q;
            double acc = 0.0;
            for (int a = 0; a < Q; ++a) acc = fma(Ys[a * ldt + r], kr[a], acc);
            c[(size_t)s * B + row0 + r] = acc;
        }
    }
    __syncthreads();
    if (!BWD) {
        for (int e = threadIdx.x; e < nrows * Q; e += blockDim.x) {
            int rr = e / Q, a = e - rr * Q;
            P[base + e] = Ys[a * ldt + rr];
        }
    } else {
        // Kbar = t + cbar p
        for (int e = threadIdx.x; e < nrows * Q; e += blockDim.x) {
            int rr = e / Q, a = e - rr * Q;
            Kbar[base + e] = fma(cbar[(size_t)s * B + row0 + rr], Ps[a * ldt + rr], Ys[a * ldt + rr]);
        }
        // Abar[s] -= t^T p  (partial over this tile)
        double* Ab = Abar + (size_t)s * Q * Q;
        for (int e = threadIdx.x; e < Q * Q; e += blockDim.x) {
            int a = e / Q, b = e - a * Q;
            double acc = 0.0;
            for (int rr = 0; rr < nrows; ++rr) acc = fma(Ys[a * ldt + rr], Ps[b * ldt + rr], acc);
            atomicAdd(&Ab[e], -acc);
        }
    }
}
static int solve_rows_tile(int Q, bool bwd) {
    // shared bytes = 8*(Q*Q + ntile*Q*(TR+1)); keep under ~200 KB
    int TR = 128;
    while (TR > 32 && 8.0 * ((double)Q * Q + (bwd ? 2.0 : 1.0) * Q * (TR + 1)) > 200.0 * 1024) TR >>= 1;
    return TR;
}
int nmgp_solve_rows_fwd_mma(const double* K, const double* R, double* P, double* c, int ns, long long B, int Q,
                            cudaStream_t st);
int nmgp_solve_rows_bwd_mma(const double* Pbar, const double* cbar, const double* K, const double* P, const double* R,
                            double* Kbar, double* Tout, int ns, long long B, int Q, cudaStream_t st);
int nmgp_atb_mma(const double* A, const double* Bm, double* C, double sign, int ns, long long B, int Q,
                 const double* cbar, double* Kbar, cudaStream_t st);

NMGP_API int nmgp_solve_rows_fwd(const double* K, const double* R, double* P, double* c, int ns, long long B, int Q,
                                 cudaStream_t st) {
    NMGP_REQUIRE(ns >= 0 && ns <= 65535 && B >= 0 && Q > 0 && Q <= 128, "nmgp_solve_rows_fwd");
    if (ns == 0 || B == 0) return 0;
    if (Q <= 128) return nmgp_solve_rows_fwd_mma(K, R, P, c, ns, B, Q, st);
    const int TR = solve_rows_tile(Q, false);
    size_t smem = sizeof(double) * ((size_t)Q * Q + (size_t)Q * (TR + 1));
    if (int r = nmgp_opt_in_smem(k_solve_rows<false>, smem, "nmgp_solve_rows_fwd")) return r;
    dim3 grid((unsigned)((B + TR - 1) / TR), ns);
    k_solve_rows<false><<<NMGP_L(grid), TR, smem, st>>>(K, R, P, c, nullptr, nullptr, nullptr, nullptr, nullptr, B, Q);
    return nmgp_launch_status("nmgp_solve_rows_fwd");
}
NMGP_API int nmgp_solve_rows_bwd(const double* Pbar, const double* cbar, const double* K, const double* P,
                                 const double* R, double* Kbar, double* Abar, double* work /* [ns,B,Q], Q <= 64 */,
                                 int ns, long long B, int Q, cudaStream_t st) {
    NMGP_REQUIRE(ns >= 0 && ns <= 65535 && B >= 0 && Q > 0 && Q <= 128, "nmgp_solve_rows_bwd");
    if (ns == 0 || B == 0) return 0;
    if (Q <= 128) {
        NMGP_REQUIRE(work != nullptr, "nmgp_solve_rows_bwd");
        if (int r = nmgp_solve_rows_bwd_mma(Pbar, cbar, K, P, R, Kbar, work, ns, B, Q, st)) return r;
        return nmgp_atb_mma(work, P, Abar, -1.0, ns, B, Q, cbar, Kbar, st);      // Abar -= T^T P ; Kbar = T + cbar P
    }
    const int TR = solve_rows_tile(Q, true);
    size_t smem = sizeof(double) * ((size_t)Q * Q + 2 * (size_t)Q * (TR + 1));
    if (int r = nmgp_opt_in_smem(k_solve_rows<true>, smem, "nmgp_solve_rows_bwd")) return r;
    dim3 grid((unsigned)((B + TR - 1) / TR), ns);
    k_solve_rows<true><<<NMGP_L(grid), TR, smem, st>>>(K, R, nullptr, nullptr, Pbar, cbar, P, Kbar, Abar, B, Q);
    return nmgp_launch_status("nmgp_solve_rows_bwd");
}

// ------------------------------------------------------------------------------------------------
// sd_ell = sqrt(s2_ell - c + eps)   (variance of tilde-ell | v, code/utils.py:233 + :32)
__global__ void k_ell_sd_fwd(const double* __restrict__ c, const double* __restrict__ hyp, double* __restrict__ sd,
                             long long B) {
    long long n = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (n < B) sd[n] = sqrt(hyp[H_S2_ELL] - c[n] + NMGP_EPS);
}
NMGP_API int nmgp_ell_sd_fwd(const double* c, const double* hyp, double* sd, long long B, cudaStream_t st) {
    if (B == 0) return 0;
    k_ell_sd_fwd<<<NMGP_L((unsigned)((B + 255) / 256)), 256, 0, st>>>(c, hyp, sd, B);
    return nmgp_launch_status("nmgp_ell_sd_fwd");
}
__global__ void k_ell_sd_bwd(const double* __restrict__ sdbar, const double* __restrict__ sd,
                             const double* __restrict__ hyp, double* __restrict__ ghyp, double* __restrict__ cbar,
                             long long B) {
    double acc = 0.0;
    for (long long n = (long long)blockIdx.x * blockDim.x + threadIdx.x; n < B; n += (long long)gridDim.x * blockDim.x) {
        double vb = sdbar[n] / (2.0 * sd[n]);
        cbar[n] = -vb;
        acc += vb;
    }
    acc = block_sum(acc);
    if (threadIdx.x == 0) atomicAdd(&ghyp[H_S2_ELL], acc * hyp[H_S2_ELL]);
}
NMGP_API int nmgp_ell_sd_bwd(const double* sdbar, const double* sd, const double* hyp, double* ghyp, double* cbar,
                             long long B, cudaStream_t st) {
    if (B == 0) return 0;
    long long blocks = (B + 255) / 256;
    if (blocks > 592) blocks = 592;
    k_ell_sd_bwd<<<NMGP_L((unsigned)blocks), 256, 0, st>>>(sdbar, sd, hyp, ghyp, cbar, B);
    return nmgp_launch_status("nmgp_ell_sd_bwd");
}

// ------------------------------------------------------------------------------------------------
// ellx[s,n] = exp(P_ell[n,:] . v[s,:] + z[s,n] sd[n]);  warp per row, lanes over q (Q <= 128), v in smem.
__global__ void k_ell_rows_fwd(const double* __restrict__ Pell, const double* __restrict__ v,
                               const double* __restrict__ zell, const double* __restrict__ sd,
                               double* __restrict__ ellx, int ns, long long B, int Q) {
    extern __shared__ double vs[];  // [ns*Q]
    for (int e = threadIdx.x; e < ns * Q; e += blockDim.x) vs[e] = v[e];
    __syncthreads();
    const int lane = threadIdx.x & 31;
    long long n = (long long)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (n >= B) return;
    double p[4];
#pragma unroll
    for (int u = 0; u < 4; ++u) {
        int q = lane + 32 * u;
        p[u] = q < Q ? Pell[(size_t)n * Q + q] : 0.0;
    }
    const double sdn = sd[n];
    for (int s = 0; s < ns; ++s) {
        double acc = 0.0;
#pragma unroll
        for (int u = 0; u < 4; ++u) {
            int q = lane + 32 * u;
            if (q < Q) acc = fma(p[u], vs[s * Q + q], acc);
        }
        acc = warp_sum(acc);
        if (lane == 0) ellx[(size_t)s * B + n] = exp(fma(zell[(size_t)s * B + n], sdn, acc));
    }
}
NMGP_API int nmgp_ell_rows_fwd(const double* Pell, const double* v, const double* zell, const double* sd, double* ellx,
                               int ns, long long B, int Q, cudaStream_t st) {
    NMGP_REQUIRE(ns > 0 && B >= 0 && Q > 0 && Q <= 128 && (size_t)ns * Q * 8 <= 200 * 1024, "nmgp_ell_rows_fwd");
    if (B == 0) return 0;
    size_t smem = (size_t)ns * Q * sizeof(double);
    if (int r = nmgp_opt_in_smem(k_ell_rows_fwd, smem, "nmgp_ell_rows_fwd")) return r;
    k_ell_rows_fwd<<<NMGP_L((unsigned)((B + 7) / 8)), 256, smem, st>>>(Pell, v, zell, sd, ellx, ns, B, Q);
    return nmgp_launch_status("nmgp_ell_rows_fwd");
}
// tb = ellxbar*ellx;  vbar[s,q] += sum_n tb P[n,q];  Pellbar[n,q] += sum_s tb v[s,q];  sdbar[n] += sum_s tb z
#define ER_TILE 64
__global__ void k_ell_rows_bwd(const double* __restrict__ ellxbar, const double* __restrict__ ellx,
                               const double* __restrict__ Pell, const double* __restrict__ v,
                               const double* __restrict__ zell, double* __restrict__ vbar, double* __restrict__ Pellbar,
                               double* __restrict__ sdbar, int ns, long long B, int Q) {
    extern __shared__ double sm[];
    double* vs = sm;                       // [ns*Q]
    double* tb = vs + (size_t)ns * Q;      // [ns][ER_TILE]
    double* Pt = tb + (size_t)ns * ER_TILE;  // [ER_TILE][Q]
    const long long row0 = (long long)blockIdx.x * ER_TILE;
    const int nrows = (int)min((long long)ER_TILE, B - row0);
    for (int e = threadIdx.x; e < ns * Q; e += blockDim.x) vs[e] = v[e];
    for (int e = threadIdx.x; e < ns * ER_TILE; e += blockDim.x) {
        int s = e / ER_TILE, r = e - s * ER_TILE;
        tb[e] = r < nrows ? ellxbar[(size_t)s * B + row0 + r] * ellx[(size_t)s * B + row0 + r] : 0.0;
    }
    for (int e = threadIdx.x; e < nrows * Q; e += blockDim.x) Pt[e] = Pell[(size_t)row0 * Q + e];
    __syncthreads();
    for (int e = threadIdx.x; e < nrows * Q; e += blockDim.x) {
        int r = e / Q, q = e - r * Q;
        double acc = 0.0;
        for (int s = 0; s < ns; ++s) acc = fma(tb[s * ER_TILE + r], vs[s * Q + q], acc);
        Pellbar[(size_t)row0 * Q + e] += acc;
    }
    for (int r = threadIdx.x; r < nrows; r += blockDim.x) {
        double acc = 0.0;
        for (int s = 0; s < ns; ++s) acc = fma(tb[s * ER_TILE + r], zell[(size_t)s * B + row0 + r], acc);
        sdbar[row0 + r] += acc;
    }
    for (int e = threadIdx.x; e < ns * Q; e += blockDim.x) {
        int s = e / Q, q = e - s * Q;
        double acc = 0.0;
        for (int r = 0; r < nrows; ++r) acc = fma(tb[s * ER_TILE + r], Pt[r * Q + q], acc);
        atomicAdd(&vbar[e], acc);
    }
}
NMGP_API int nmgp_ell_rows_bwd(const double* ellxbar, const double* ellx, const double* Pell, const double* v,
                               const double* zell, double* vbar, double* Pellbar, double* sdbar, int ns, long long B,
                               int Q, cudaStream_t st) {
    NMGP_REQUIRE(ns > 0 && B >= 0 && Q > 0, "nmgp_ell_rows_bwd");
    if (B == 0) return 0;
    size_t smem = sizeof(double) * ((size_t)ns * Q + (size_t)ns * ER_TILE + (size_t)ER_TILE * Q);
    if (int r = nmgp_opt_in_smem(k_ell_rows_bwd, smem, "nmgp_ell_rows_bwd")) return r;
    k_ell_rows_bwd<<<NMGP_L((unsigned)((B + ER_TILE - 1) / ER_TILE)), 256, smem, st>>>(ellxbar, ellx, Pell, v, zell, vbar,
                                                                               Pellbar, sdbar, ns, B, Q);
    return nmgp_launch_status("nmgp_ell_rows_bwd");
}

// ------------------------------------------------------------------------------------------------
// sd[n,j] = sqrt(s2_k - c_k[n] + q[n,j] + eps), k = L1 if j == I[n] else L0, for j <= I[n]; 0 otherwise.
__global__ void k_coef_sd_fwd(const double* __restrict__ q, const double* __restrict__ cL0,
                              const double* __restrict__ cL1, const int* __restrict__ I, const double* __restrict__ hyp,
                              double* __restrict__ sd, long long B, int D) {
    long long gid = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (gid >= B * D) return;
    long long n = gid / D;
    int j = (int)(gid - n * D), i = I[n];
    double out = 0.0;
    if (j <= i) {
        double base = (j == i) ? hyp[H_S2_L1] - cL1[n] : hyp[H_S2_L0] - cL0[n];
        out = sqrt(base + q[gid] + NMGP_EPS);
    }
    sd[gid] = out;
}
NMGP_API int nmgp_coef_sd_fwd(const double* q, const double* cL0, const double* cL1, const int* I, const double* hyp,
                              double* sd, long long B, int D, cudaStream_t st) {
    if (B == 0) return 0;
    long long n = B * D;
    k_coef_sd_fwd<<<NMGP_L((unsigned)((n + 255) / 256)), 256, 0, st>>>(q, cL0, cL1, I, hyp, sd, B, D);
    return nmgp_launch_status("nmgp_coef_sd_fwd");
}
// qbar = sdbar/(2 sd); cL0bar[n] = -sum_{j<i} qbar; cL1bar[n] = -qbar[n,i]; ghyp += s2 * sums.  Warp per row.
__global__ void k_coef_sd_bwd(const double* __restrict__ sdbar, const double* __restrict__ sd, const int* __restrict__ I,
                              const double* __restrict__ hyp, double* __restrict__ ghyp, double* __restrict__ qbar,
                              double* __restrict__ cL0bar, double* __restrict__ cL1bar, long long B, int D) {
    const int lane = threadIdx.x & 31;
    const int wpb = blockDim.x >> 5;
    double g0 = 0.0, g1 = 0.0;
    for (long long n = (long long)blockIdx.x * wpb + (threadIdx.x >> 5); n < B; n += (long long)gridDim.x * wpb) {
        const int i = I[n];
        double off = 0.0, dg = 0.0;
        for (int j = lane; j < D; j += 32) {
            double qb = 0.0;
            if (j <= i) {
                qb = sdbar[(size_t)n * D + j] / (2.0 * sd[(size_t)n * D + j]);
                if (j == i) dg += qb; else off += qb;
            }
            qbar[(size_t)n * D + j] = qb;
        }
        off = warp_sum(off);
        dg = warp_sum(dg);
        if (lane == 0) {
            cL0bar[n] = -off;
            cL1bar[n] = -dg;
            g0 += off;
            g1 += dg;
        }
    }
    g0 = block_sum(g0);
    g1 = block_sum(g1);
    if (threadIdx.x == 0) {
        atomicAdd(&ghyp[H_S2_L0], g0 * hyp[H_S2_L0]);
        atomicAdd(&ghyp[H_S2_L1], g1 * hyp[H_S2_L1]);
    }
}
NMGP_API int nmgp_coef_sd_bwd(const double* sdbar, const double* sd, const int* I, const double* hyp, double* ghyp,
                              double* qbar, double* cL0bar, double* cL1bar, long long B, int D, cudaStream_t st) {
    if (B == 0) return 0;
    long long blocks = (B + 7) / 8;
    if (blocks > 148 * 8) blocks = 148 * 8;
    k_coef_sd_bwd<<<NMGP_L((unsigned)blocks), 256, 0, st>>>(sdbar, sd, I, hyp, ghyp, qbar, cL0bar, cL1bar, B, D);
    return nmgp_launch_status("nmgp_coef_sd_bwd");
}

// l[s,n,j] = m + z sd  (exp on j == I[n]), 0 for j > I[n]    (code/nmgp_dsvi.py:228-238)
// zL == NULL: the noise is generated in the kernel (counter-based, philox.cuh) from (key, s0 + s, gid[n], j) and never
// touches HBM; the backward kernel regenerates the same values.
// One thread per (row, quad of 4 consecutive columns): a Philox call yields the 4 normals of the quad.
__device__ __forceinline__ void coef_noise4(const double* __restrict__ zL, size_t o, int jn, NoiseKey key, int sglob,
                                            long long gid, int q4, double z[4]) {
    if (zL) {
#pragma unroll
        for (int k = 0; k < 4; ++k) z[k] = k < jn ? zL[o + k] : 0.0;
    } else {
        float zf[4];
        philox_normal4(key, (unsigned)sglob, (unsigned long long)gid, (unsigned)q4, zf);
#pragma unroll
        for (int k = 0; k < 4; ++k) z[k] = (double)zf[k];
    }
}
__global__ void k_coef_sample_fwd(const double* __restrict__ m, const double* __restrict__ sd,
                                  const double* __restrict__ zL, const int* __restrict__ I, double* __restrict__ l,
                                  long long B, int D, NoiseKey key, int s0, const long long* __restrict__ gid) {
    const int s = blockIdx.y;
    const int nq = (D + 3) >> 2;
    long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= B * nq) return;
    long long n = e / nq;
    const int q4 = (int)(e - n * nq), j0 = 4 * q4, i = I[n];
    const int jn = min(4, D - j0);
    const size_t o = (size_t)n * D + j0, os = (size_t)s * B * D + o;
    double z[4] = {0.0, 0.0, 0.0, 0.0};
    if (j0 <= i) coef_noise4(zL, os, jn, key, s0 + s, gid ? gid[n] : n, q4, z);
#pragma unroll
    for (int k = 0; k < 4; ++k) {
        if (k < jn) {
            const int j = j0 + k;
            double out = 0.0;
            if (j <= i) {
                out = fma(z[k], sd[o + k], m[o + k]);
                if (j == i) out = exp(out);
            }
            l[os + k] = out;
        }
    }
}
NMGP_API int nmgp_coef_sample_fwd(const double* m, const double* sd, const double* zL, const int* I, double* l, int ns,
                                  long long B, int D, unsigned long long seed, unsigned long long stream_id, int s0,
                                  const long long* gid, const unsigned long long* step_dev, cudaStream_t st) {
    NMGP_REQUIRE(ns >= 0 && ns <= 65535, "nmgp_coef_sample_fwd");
    if (B == 0 || ns == 0) return 0;
    long long n = B * ((D + 3) / 4);
    dim3 grid((unsigned)((n + 255) / 256), ns);
    NoiseKey key{seed, stream_id, step_dev};
    k_coef_sample_fwd<<<NMGP_L(grid), 256, 0, st>>>(m, sd, zL, I, l, B, D, key, s0, gid);
    return nmgp_launch_status("nmgp_coef_sample_fwd");
}
__global__ void k_coef_sample_bwd(const double* __restrict__ lbar, const double* __restrict__ l,
                                  const double* __restrict__ zL, const int* __restrict__ I, double* __restrict__ mbar,
                                  double* __restrict__ sdbar, int ns, long long B, int D, NoiseKey key, int s0,
                                  const long long* __restrict__ gid) {
    const int nq = (D + 3) >> 2;
    long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= B * nq) return;
    long long n = e / nq;
    const int q4 = (int)(e - n * nq), j0 = 4 * q4, i = I[n];
    if (j0 > i) return;
    const int jn = min(4, D - j0);
    const long long g = gid ? gid[n] : n;
    const size_t o = (size_t)n * D + j0;
    double am[4] = {0.0, 0.0, 0.0, 0.0}, as[4] = {0.0, 0.0, 0.0, 0.0};
    for (int s = 0; s < ns; ++s) {
        const size_t os = (size_t)s * B * D + o;
        double z[4];
        coef_noise4(zL, os, jn, key, s0 + s, g, q4, z);
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            if (k < jn && j0 + k <= i) {
                double rb = lbar[os + k];
                if (j0 + k == i) rb *= l[os + k];
                am[k] += rb;
                as[k] = fma(rb, z[k], as[k]);
            }
        }
    }
#pragma unroll
    for (int k = 0; k < 4; ++k) {
        if (k < jn && j0 + k <= i) {
            mbar[o + k] += am[k];
            sdbar[o + k] += as[k];
        }
    }
}
NMGP_API int nmgp_coef_sample_bwd(const double* lbar, const double* l, const double* zL, const int* I, double* mbar,
                                  double* sdbar, int ns, long long B, int D, unsigned long long seed,
                                  unsigned long long stream_id, int s0, const long long* gid,
                                  const unsigned long long* step_dev, cudaStream_t st) {
    if (B == 0 || ns == 0) return 0;
    long long n = B * ((D + 3) / 4);
    NoiseKey key{seed, stream_id, step_dev};
    k_coef_sample_bwd<<<NMGP_L((unsigned)((n + 255) / 256)), 256, 0, st>>>(lbar, l, zL, I, mbar, sdbar, ns, B, D, key, s0, gid);
    return nmgp_launch_status("nmgp_coef_sample_bwd");
}

// out[s, n, c] = N(0,1) (float32 precision) for columns c < C: the explicit form of the same generator
__global__ void k_noise_fill(double* __restrict__ out, long long B, int C, NoiseKey key, int s0,
                             const long long* __restrict__ gid) {
    const int s = blockIdx.y;
    const int nq = (C + 3) / 4;
    long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= B * nq) return;
    long long n = e / nq;
    int q4 = (int)(e - n * nq);
    float z[4];
    philox_normal4(key, (unsigned)(s0 + s), (unsigned long long)(gid ? gid[n] : n), (unsigned)q4, z);
#pragma unroll
    for (int k = 0; k < 4; ++k) {
        int c = 4 * q4 + k;
        if (c < C) out[((size_t)s * B + n) * C + c] = (double)z[k];
    }
}
NMGP_API int nmgp_noise_fill(double* out, int ns, long long B, int C, unsigned long long seed,
                             unsigned long long stream_id, int s0, const long long* gid,
                             const unsigned long long* step_dev, cudaStream_t st) {
    NMGP_REQUIRE(ns >= 0 && ns <= 65535 && B >= 0 && C > 0, "nmgp_noise_fill");
    if (B == 0 || ns == 0) return 0;
    long long n = B * ((C + 3) / 4);
    dim3 grid((unsigned)((n + 255) / 256), ns);
    NoiseKey key{seed, stream_id, step_dev};
    k_noise_fill<<<NMGP_L(grid), 256, 0, st>>>(out, B, C, key, s0, gid);
    return nmgp_launch_status("nmgp_noise_fill");
}

// ------------------------------------------------------------------------------------------------
// Expected log-likelihood of one sample row (code/nmgp_dsvi.py:255-258, code/utils.py:268-272) + cotangents for
// loss = -scale * sum_s R_s.  Warp per (s, n), lanes over j <= I[n].
__global__ void k_lik_rows(const double* __restrict__ l, const double* __restrict__ mg, const double* __restrict__ qg,
                           const double* __restrict__ cG, const double* __restrict__ y, const int* __restrict__ I,
                           const double* __restrict__ hyp, double scale, double* __restrict__ Rsum,
                           double* __restrict__ ghyp, double* __restrict__ lbar, double* __restrict__ mgbar,
                           double* __restrict__ qgbar, double* __restrict__ cGbar, long long B, int D, long long ystride) {
    const int s = blockIdx.y;
    const int lane = threadIdx.x & 31, wpb = blockDim.x >> 5;
    const double s2e = hyp[H_S2_ERR];
    const double cst = -0.5 * log(s2e) - log(sqrt(2.0 * 3.14159265358979323846));
    double racc = 0.0, gacc = 0.0;
    for (long long n = (long long)blockIdx.x * wpb + (threadIdx.x >> 5); n < B; n += (long long)gridDim.x * wpb) {
        const int i = I[n];
        const size_t o = ((size_t)s * B + n) * D;
        const double one_m_c = 1.0 - cG[(size_t)s * B + n];
        double F = 0.0, pen = 0.0;
        for (int j = lane; j <= i; j += 32) {
            double lj = l[o + j];
            F = fma(lj, mg[o + j], F);
            pen = fma(lj * lj, one_m_c + qg[o + j], pen);
        }
        F = warp_sum(F);
        pen = warp_sum(pen);
        const double r = y[(size_t)s * ystride + n] - F;
        const double rr = r / s2e;
        double qsum = 0.0;
        for (int j = lane; j < D; j += 32) {
            double mb = 0.0, qb = 0.0, lb = 0.0;
            if (j <= i) {
                double lj = l[o + j];
                mb = -scale * rr * lj;
                qb = scale * (0.5 / s2e) * lj * lj;
                lb = -scale * (rr * mg[o + j] - (1.0 / s2e) * lj * (one_m_c + qg[o + j]));
                qsum += qb;
            }
            mgbar[o + j] = mb;
            qgbar[o + j] = qb;
            lbar[o + j] = lb;
        }
        qsum = warp_sum(qsum);
        if (lane == 0) {
            cGbar[(size_t)s * B + n] = -qsum;
            racc += -(r * r) / (2.0 * s2e) + cst - (0.5 / s2e) * pen;
            gacc += (r * r) / (2.0 * s2e) - 0.5 + (0.5 / s2e) * pen;
        }
    }
    racc = block_sum(racc);
    gacc = block_sum(gacc);
    if (threadIdx.x == 0) {
        atomicAdd(&Rsum[s], racc);
        atomicAdd(&ghyp[H_S2_ERR], -scale * gacc);
    }
}
NMGP_API int nmgp_lik_rows(const double* l, const double* mg, const double* qg, const double* cG, const double* y,
                           const int* I, const double* hyp, double scale, double* Rsum /* pre-zeroed */, double* ghyp,
                           double* lbar, double* mgbar, double* qgbar, double* cGbar, int ns, long long B, int D,
                           long long ystride, cudaStream_t st) {
    NMGP_REQUIRE(ns >= 0 && ns <= 65535 && (ystride == 0 || ystride >= B), "nmgp_lik_rows");
    if (B == 0 || ns == 0) return 0;
    long long blocks = (B + 7) / 8;
    if (blocks > 148 * 4) blocks = 148 * 4;
    dim3 grid((unsigned)blocks, ns);
    k_lik_rows<<<NMGP_L(grid), 256, 0, st>>>(l, mg, qg, cG, y, I, hyp, scale, Rsum, ghyp, lbar, mgbar, qgbar, cGbar, B, D, ystride);
    return nmgp_launch_status("nmgp_lik_rows");
}

// ------------------------------------------------------------------------------------------------
// Means only (no quadratic forms): m[s,n,j] = p . Mu[idx] for j <= I[n]   (code/utils.py:149-157 MGP_mu,
// used by predict_Y, code/nmgp_dsvi.py:698-718).  Warp per (s,n), lanes over q.
__global__ void k_pair_means(const double* __restrict__ Pa, const double* __restrict__ Pb, const int* __restrict__ I,
                             const double* __restrict__ Mu, double* __restrict__ m, long long B, int Q, int D,
                             int mode) {
    const int s = blockIdx.y;
    const int lane = threadIdx.x & 31;
    long long n = (long long)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (n >= B) return;
    const int i = I[n];
    const size_t prow = ((size_t)s * B + n) * Q;
    double pa[4], pb[4];
#pragma unroll
    for (int u = 0; u < 4; ++u) {
        int q = lane + 32 * u;
        pa[u] = q < Q ? Pa[prow + q] : 0.0;
        pb[u] = (mode == MODE_U && q < Q) ? Pb[prow + q] : pa[u];
    }
    double* mo = m + ((size_t)s * B + n) * D;
    for (int j = 0; j < D; ++j) {
        double acc = 0.0;
        if (j <= i) {
            const int idx = (mode == MODE_U) ? pair_slot(i, j, D) : j;
            const bool useB = (mode == MODE_U) && (j == i);
#pragma unroll
            for (int u = 0; u < 4; ++u) {
                int q = lane + 32 * u;
                if (q < Q) acc = fma(useB ? pb[u] : pa[u], Mu[(size_t)idx * Q + q], acc);
            }
            acc = warp_sum(acc);
        }
        if (lane == 0) mo[j] = acc;
    }
}
NMGP_API int nmgp_pair_means(const double* Pa, const double* Pb, const int* I, const double* Mu, double* m, int ns,
                             long long B, int Q, int D, int mode, cudaStream_t st) {
    NMGP_REQUIRE(ns >= 0 && ns <= 65535 && B >= 0 && Q > 0 && Q <= 128 && D > 0, "nmgp_pair_means");
    if (ns == 0 || B == 0) return 0;
    dim3 grid((unsigned)((B + 7) / 8), ns);
    k_pair_means<<<NMGP_L(grid), 256, 0, st>>>(Pa, Pb, I, Mu, m, B, Q, D, mode);
    return nmgp_launch_status("nmgp_pair_means");
}

// F[s,n] = sum_{j <= I[n]} l[s,n,j] g[s,n,j]     (code/nmgp_dsvi.py:255 and :721-722)
__global__ void k_rowdot_live(const double* __restrict__ l, const double* __restrict__ g, const int* __restrict__ I,
                              double* __restrict__ F, long long B, int D) {
    const int s = blockIdx.y;
    const int lane = threadIdx.x & 31;
    long long n = (long long)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (n >= B) return;
    const int i = I[n];
    const size_t o = ((size_t)s * B + n) * D;
    double acc = 0.0;
    for (int j = lane; j <= i; j += 32) acc = fma(l[o + j], g[o + j], acc);
    acc = warp_sum(acc);
    if (lane == 0) F[(size_t)s * B + n] = acc;
}
NMGP_API int nmgp_rowdot_live(const double* l, const double* g, const int* I, double* F, int ns, long long B, int D,
                              cudaStream_t st) {
    NMGP_REQUIRE(ns >= 0 && ns <= 65535 && B >= 0 && D > 0, "nmgp_rowdot_live");
    if (ns == 0 || B == 0) return 0;
    dim3 grid((unsigned)((B + 7) / 8), ns);
    k_rowdot_live<<<NMGP_L(grid), 256, 0, st>>>(l, g, I, F, B, D);
    return nmgp_launch_status("nmgp_rowdot_live");
}

// ------------------------------------------------------------------------------------------------
// element helpers behind the reference's small utilities (code/utils.py:15-33, 268-287)
// out = mean + z * sqrt(var + eps)
__global__ void k_reparam_diag(const double* __restrict__ mean, const double* __restrict__ var,
                               const double* __restrict__ z, double* __restrict__ out, long long n) {
    long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) out[i] = fma(z[i], sqrt(var[i] + NMGP_EPS), mean[i]);
}
NMGP_API int nmgp_reparam_diag(const double* mean, const double* var, const double* z, double* out, long long n,
                               cudaStream_t st) {
    if (n <= 0) return 0;
    k_reparam_diag<<<NMGP_L((unsigned)((n + 255) / 256)), 256, 0, st>>>(mean, var, z, out, n);
    return nmgp_launch_status("nmgp_reparam_diag");
}
// out += sum_i [ -(y-loc)^2/(2 scale^2) - log(scale) - log(sqrt(2 pi)) ]   (scale: scalar on device)
__global__ void k_normal_logprob(const double* __restrict__ loc, const double* __restrict__ scale,
                                 const double* __restrict__ y, double* __restrict__ out, long long n) {
    const double sc = scale[0], var = sc * sc;
    const double cst = -log(sc) - log(sqrt(2.0 * 3.14159265358979323846));
    double s = 0.0;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
        double r = y[i] - loc[i];
        s += -(r * r) / (2.0 * var) + cst;
    }
    s = block_sum(s);
    if (threadIdx.x == 0) atomicAdd(out, s);
}
NMGP_API int nmgp_normal_logprob_sum(const double* loc, const double* scale, const double* y, double* out /* += */,
                                     long long n, cudaStream_t st) {
    if (n <= 0) return 0;
    long long blocks = (n + 255) / 256;
    if (blocks > 592) blocks = 592;
    k_normal_logprob<<<NMGP_L((unsigned)blocks), 256, 0, st>>>(loc, scale, y, out, n);
    return nmgp_launch_status("nmgp_normal_logprob_sum");
}
// out[r] = sum_k x[r,k]^2   (warp per row)
__global__ void k_sumsq_rows(const double* __restrict__ x, double* __restrict__ out, long long rows, long long cols) {
    const int lane = threadIdx.x & 31;
    long long r = (long long)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (r >= rows) return;
    double s = 0.0;
    for (long long k = lane; k < cols; k += 32) {
        double v = x[r * cols + k];
        s = fma(v, v, s);
    }
    s = warp_sum(s);
    if (lane == 0) out[r] = s;
}
NMGP_API int nmgp_sumsq_rows(const double* x, double* out, long long rows, long long cols, cudaStream_t st) {
    if (rows <= 0) return 0;
    k_sumsq_rows<<<NMGP_L((unsigned)((rows + 7) / 8)), 256, 0, st>>>(x, out, rows, cols);
    return nmgp_launch_status("nmgp_sumsq_rows");
}

// corr[m] = diag(cov)^-1/2 cov diag(cov)^-1/2 with cov = L[m] L[m]^T for a batch of small D x D lower factors
// (code/nmgp_dsvi.py:567-569, the per-point output correlations of sample_FY).  One thread per (m, a, b).
__global__ void k_lcorr(const double* __restrict__ L, double* __restrict__ corr, long long nmat, int D) {
    long long gid = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (gid >= nmat * D * D) return;
    long long m = gid / (D * D);
    int r = (int)(gid - m * D * D), a = r / D, b = r - a * D;
    const double* Lm = L + m * D * D;
    double cab = 0.0, caa = 0.0, cbb = 0.0;
    for (int k = 0; k < D; ++k) {
        double la = Lm[a * D + k], lb = Lm[b * D + k];
        cab = fma(la, lb, cab);
        caa = fma(la, la, caa);
        cbb = fma(lb, lb, cbb);
    }
    corr[gid] = sqrt(1.0 / caa) * cab * sqrt(1.0 / cbb);
}
NMGP_API int nmgp_lcorr(const double* L, double* corr, long long nmat, int D, cudaStream_t st) {
    NMGP_REQUIRE(nmat >= 0 && D > 0, "nmgp_lcorr");
    if (nmat == 0) return 0;
    long long n = nmat * D * D;
    k_lcorr<<<NMGP_L((unsigned)((n + 255) / 256)), 256, 0, st>>>(L, corr, nmat, D);
    return nmgp_launch_status("nmgp_lcorr");
}
