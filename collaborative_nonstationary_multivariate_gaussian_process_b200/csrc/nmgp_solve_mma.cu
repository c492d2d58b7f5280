// Row solves P = K (R R^T)^-1 as a blocked triangular substitution on the FP64 tensor cores (Q <= 128; left-looking
// register kernel for Q <= 64, right-looking kernel k_solve_rows_rl above that).
//
// Each warp owns 16 rows.  With 8-column blocks b,
//   forward  (Y R^T = K):  Y_b = (K_b - sum_{c<b} Y_c R_{b,c}^T) inv(R_bb)^T
//   backward (P R   = Y):  P_b = (Y_b - sum_{c>b} P_c R_{c,b})   inv(R_bb)
// every product is a DMMA m8n8k4 with the running block as accumulator; the 8x8 diagonal blocks of the Cholesky factor
// are inverted once per CTA in shared memory (they are as well conditioned as R itself: cond(R) = sqrt(cond(K22+eps I))).
// Results of one block are re-used as A operands of later blocks after an in-quad shuffle (C-fragment -> A-fragment).
// Replaces the LU solves of code/utils.py:119,142,154,230 and their autograd; the Q x Q adjoint Abar -= T^T P is the
// separate DMMA reduction k_atb_mma (nmgp_atb_mma.cu).
#include "common.cuh"

__device__ __forceinline__ void dmma884m(double& d0, double& d1, double a, double b) {
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n"
                 : "+d"(d0), "+d"(d1)
                 : "d"(a), "d"(b));
}
// C-fragment (row g, cols 2t, 2t+1 of an 8x8 block) -> A-fragments of its two k-steps (row g, col 4ks + t)
__device__ __forceinline__ void c_to_a(double c0, double c1, int t, int lane, double& a0, double& a1) {
    const int base = lane & ~3;
    const int src0 = base + (t >> 1), src1 = base + 2 + (t >> 1);
    double x0 = __shfl_sync(0xffffffffu, c0, src0), x1 = __shfl_sync(0xffffffffu, c1, src0);
    a0 = (t & 1) ? x1 : x0;
    x0 = __shfl_sync(0xffffffffu, c0, src1);
    x1 = __shfl_sync(0xffffffffu, c1, src1);
    a1 = (t & 1) ? x1 : x0;
}

#define SM_ROWS 128
#define SM_THREADS 256
#define SM_TILES 4        // 128-row tiles per CTA: the factor staging / diagonal-block inversion is paid once per 512 rows

template <int NB, bool BWD>
__global__ void __launch_bounds__(SM_THREADS, 2)
k_solve_rows_mma(const double* __restrict__ K, const double* __restrict__ R, double* __restrict__ P,
                 double* __restrict__ cout, const double* __restrict__ Pbar, const double* __restrict__ cbar,
                 const double* __restrict__ Pin, double* __restrict__ Kbar, double* __restrict__ Tout, long long B,
                 int Q) {
    constexpr int QP = 8 * NB;
    constexpr int LDR = ((QP + 3) / 8) * 8 + 4;          // % 8 == 4: conflict-free B-fragment reads
    extern __shared__ __align__(16) double sm[];
    double* Rn = sm;                         // [QP][LDR]  -R (strictly below the diagonal blocks), row-major
    double* RTn = Rn + QP * LDR;             // [QP][LDR]  -R^T
    double* Ri = RTn + QP * LDR;             // [NB][8][8] inv(R_bb)        (Ri[b][r][c])
    double* RiT = Ri + NB * 64;              // [NB][8][8] inv(R_bb)^T
    const int s = blockIdx.y, tid = threadIdx.x, lane = tid & 31, w = tid >> 5, g = lane >> 2, t = lane & 3;
    const double* Rg = R + (size_t)s * Q * Q;
    for (int e = tid; e < QP * LDR; e += SM_THREADS) {
        int a = e / LDR, b = e - a * LDR;
        double v = (a < Q && b < Q && b < a) ? -Rg[(size_t)a * Q + b] : 0.0;     // strictly lower, negated
        Rn[e] = v;
    }
    __syncthreads();
    for (int e = tid; e < QP * LDR; e += SM_THREADS) {
        int a = e / LDR, b = e - a * LDR;
        RTn[e] = (b < QP) ? Rn[b * LDR + a] : 0.0;
    }
    // inverse of each 8x8 lower-triangular diagonal block: thread (b, col) solves R_bb x = e_col
    if (tid < NB * 8) {
        const int b = tid >> 3, col = tid & 7;
        double xcol[8];
#pragma unroll
        for (int r = 0; r < 8; ++r) {
            const int gr = 8 * b + r;
            double sacc = (r == col) ? 1.0 : 0.0;
#pragma unroll
            for (int k = 0; k < 8; ++k)
                if (k < r) {
                    const int gk = 8 * b + k;
                    double rv = (gr < Q && gk < Q) ? Rg[(size_t)gr * Q + gk] : 0.0;
                    sacc = fma(-rv, xcol[k], sacc);
                }
            const double dg = gr < Q ? Rg[(size_t)gr * Q + gr] : 1.0;
            xcol[r] = sacc / dg;
        }
#pragma unroll
        for (int r = 0; r < 8; ++r) {
            Ri[b * 64 + r * 8 + col] = xcol[r];
            RiT[b * 64 + col * 8 + r] = xcol[r];
        }
    }
    __syncthreads();

  for (int tile = 0; tile < SM_TILES; ++tile) {
    const long long row0 = ((long long)blockIdx.x * SM_TILES + tile) * SM_ROWS;
    if (row0 >= B) break;
    const int rl[2] = {16 * w + g, 16 * w + 8 + g};
    const long long gr_[2] = {row0 + rl[0], row0 + rl[1]};
    const bool ok[2] = {gr_[0] < B, gr_[1] < B};
    const size_t rb[2] = {((size_t)s * B + (ok[0] ? gr_[0] : 0)) * Q, ((size_t)s * B + (ok[1] ? gr_[1] : 0)) * Q};
    double cb[2] = {0.0, 0.0};
    if (BWD) {
        cb[0] = ok[0] ? cbar[(size_t)s * B + gr_[0]] : 0.0;
        cb[1] = ok[1] ? cbar[(size_t)s * B + gr_[1]] : 0.0;
    }
    auto load_rhs = [&](int mb, int nb, double& v0, double& v1) {
        const int c0 = 8 * nb + 2 * t;
        v0 = v1 = 0.0;
        if (ok[mb]) {
            if (c0 < Q) v0 = BWD ? fma(cb[mb], K[rb[mb] + c0], Pbar[rb[mb] + c0]) : K[rb[mb] + c0];
            if (c0 + 1 < Q) v1 = BWD ? fma(cb[mb], K[rb[mb] + c0 + 1], Pbar[rb[mb] + c0 + 1]) : K[rb[mb] + c0 + 1];
        }
    };

    double yA[2][NB][2];      // A-fragments of finished blocks (forward: Y, then reused for P in the backward sweep)
    double yC[2][NB][2];      // right-hand sides first (all global loads in flight at once), then C-fragments of Y
#pragma unroll
    for (int b = 0; b < NB; ++b) {
        load_rhs(0, b, yC[0][b][0], yC[0][b][1]);
        load_rhs(1, b, yC[1][b][0], yC[1][b][1]);
    }
    // ---- forward sweep ---------------------------------------------------------------------------------------
#pragma unroll
    for (int b = 0; b < NB; ++b) {
        double acc[2][2] = {{yC[0][b][0], yC[0][b][1]}, {yC[1][b][0], yC[1][b][1]}};
#pragma unroll
        for (int c = 0; c < b; ++c) {
#pragma unroll
            for (int ks = 0; ks < 2; ++ks) {
                const double bf = Rn[(8 * b + g) * LDR + 8 * c + 4 * ks + t];       // B[k][n] = -R[8b+n][8c+k]
                dmma884m(acc[0][0], acc[0][1], yA[0][c][ks], bf);
                dmma884m(acc[1][0], acc[1][1], yA[1][c][ks], bf);
            }
        }
#pragma unroll
        for (int mb = 0; mb < 2; ++mb) {
            double a0, a1;
            c_to_a(acc[mb][0], acc[mb][1], t, lane, a0, a1);
            double y0 = 0.0, y1 = 0.0;
            dmma884m(y0, y1, a0, Ri[b * 64 + g * 8 + t]);             // B[k][n] = inv(R_bb)^T[k][n] = inv(R_bb)[n][k]
            dmma884m(y0, y1, a1, Ri[b * 64 + g * 8 + 4 + t]);
            yC[mb][b][0] = y0;
            yC[mb][b][1] = y1;
            c_to_a(y0, y1, t, lane, yA[mb][b][0], yA[mb][b][1]);
        }
    }
    // ---- backward sweep ----------------------------------------------------------------------------------------
    double csum[2] = {0.0, 0.0};
#pragma unroll
    for (int b = NB - 1; b >= 0; --b) {
        double acc[2][2] = {{yC[0][b][0], yC[0][b][1]}, {yC[1][b][0], yC[1][b][1]}};
#pragma unroll
        for (int c = NB - 1; c > b; --c) {
#pragma unroll
            for (int ks = 0; ks < 2; ++ks) {
                const double bf = RTn[(8 * b + g) * LDR + 8 * c + 4 * ks + t];      // B[k][n] = -R[8c+k][8b+n]
                dmma884m(acc[0][0], acc[0][1], yA[0][c][ks], bf);
                dmma884m(acc[1][0], acc[1][1], yA[1][c][ks], bf);
            }
        }
#pragma unroll
        for (int mb = 0; mb < 2; ++mb) {
            double a0, a1;
            c_to_a(acc[mb][0], acc[mb][1], t, lane, a0, a1);
            double p0 = 0.0, p1 = 0.0;
            dmma884m(p0, p1, a0, RiT[b * 64 + g * 8 + t]);            // B[k][n] = inv(R_bb)[k][n] = RiT[n][k]
            dmma884m(p0, p1, a1, RiT[b * 64 + g * 8 + 4 + t]);
            c_to_a(p0, p1, t, lane, yA[mb][b][0], yA[mb][b][1]);      // P_b as A operand of the blocks to its left
            const int c0 = 8 * b + 2 * t;
            if (ok[mb]) {
                if (!BWD) {
                    if (c0 < Q) { P[rb[mb] + c0] = p0; csum[mb] = fma(p0, K[rb[mb] + c0], csum[mb]); }
                    if (c0 + 1 < Q) { P[rb[mb] + c0 + 1] = p1; csum[mb] = fma(p1, K[rb[mb] + c0 + 1], csum[mb]); }
                } else {
                    if (c0 < Q) Tout[rb[mb] + c0] = p0;          // Kbar = T + cbar P is written (coalesced) by k_atb_mma
                    if (c0 + 1 < Q) Tout[rb[mb] + c0 + 1] = p1;
                }
            }
        }
    }
    if (!BWD) {
#pragma unroll
        for (int mb = 0; mb < 2; ++mb) {
            double v = csum[mb];
            v += __shfl_xor_sync(0xffffffffu, v, 1);
            v += __shfl_xor_sync(0xffffffffu, v, 2);
            if (t == 0 && ok[mb]) cout[(size_t)s * B + gr_[mb]] = v;
        }
    }
  }
}

template <int NB, bool BWD>
static int launch_solve_mma(const double* K, const double* R, double* P, double* c, const double* Pbar,
                            const double* cbar, const double* Pin, double* Kbar, double* Tout, int ns, long long B,
                            int Q, cudaStream_t st, const char* what) {
    constexpr int QP = 8 * NB, LDR = ((QP + 3) / 8) * 8 + 4;
    size_t smem = sizeof(double) * (2 * QP * LDR + 2 * NB * 64);
    if (int r = nmgp_opt_in_smem(k_solve_rows_mma<NB, BWD>, smem, what)) return r;
    dim3 grid((unsigned)((B + SM_ROWS * SM_TILES - 1) / (SM_ROWS * SM_TILES)), ns);
    k_solve_rows_mma<NB, BWD><<<NMGP_L(grid), SM_THREADS, smem, st>>>(K, R, P, c, Pbar, cbar, Pin, Kbar, Tout, B, Q);
    return nmgp_launch_status(what);
}

// ------------------------------------------------------------------------------------------------------------
// 8 < NB <= 16 (64 < Q <= 128): right-looking form of the same substitution.  As soon as a block Y_b (P_b) is final it
// is applied to ALL remaining right-hand-side blocks (independent DMMAs: the dependency chain is only the NB diagonal
// steps), so the only per-row state is the NB running blocks -- no A-fragment copy of every finished block, which is
// what limits the left-looking kernel above to NB <= 8.  One copy of -R serves both sweeps: with a row stride
// = 4 or 12 (mod 16) doubles the direct (rows by g, columns by t) and the transposed (rows by t, columns by g)
// B-fragment reads are both bank-conflict free.  One CTA of 8 warps per SM, RL_TILES 128-row tiles per CTA.
#define RL_TILES 4

template <int NB, bool BWD>
__global__ void __launch_bounds__(SM_THREADS, 1)
k_solve_rows_rl(const double* __restrict__ K, const double* __restrict__ R, double* __restrict__ P,
                double* __restrict__ cout, const double* __restrict__ Pbar, const double* __restrict__ cbar,
                double* __restrict__ Tout, long long B, int Q, int right_only) {
    constexpr int QP = 8 * NB;
    constexpr int LDR = QP + 4;
    extern __shared__ __align__(16) double sm[];
    double* Rn = sm;                         // [QP][LDR]  -R strictly below the diagonal
    double* Ri = Rn + QP * LDR;              // [NB][8][8] inv(R_bb)
    double* RiT = Ri + NB * 64;              // [NB][8][8] inv(R_bb)^T
    const int s = blockIdx.y, tid = threadIdx.x, lane = tid & 31, w = tid >> 5, g = lane >> 2, t = lane & 3;
    const double* Rg = R + (size_t)s * Q * Q;
    for (int e = tid; e < QP * LDR; e += SM_THREADS) {
        const int a = e / LDR, b = e - a * LDR;
        Rn[e] = (a < Q && b < a) ? -Rg[(size_t)a * Q + b] : 0.0;
    }
    if (tid < NB * 8) {
        const int b = tid >> 3, col = tid & 7;
        double xcol[8];
#pragma unroll
        for (int r = 0; r < 8; ++r) {
            const int gr = 8 * b + r;
            double sacc = (r == col) ? 1.0 : 0.0;
#pragma unroll
            for (int k = 0; k < 8; ++k)
                if (k < r) {
                    const int gk = 8 * b + k;
                    const double rv = (gr < Q && gk < Q) ? Rg[(size_t)gr * Q + gk] : 0.0;
                    sacc = fma(-rv, xcol[k], sacc);
                }
            const double dg = gr < Q ? Rg[(size_t)gr * Q + gr] : 1.0;
            xcol[r] = sacc / dg;
        }
#pragma unroll
        for (int r = 0; r < 8; ++r) {
            Ri[b * 64 + r * 8 + col] = xcol[r];
            RiT[b * 64 + col * 8 + r] = xcol[r];
        }
    }
    __syncthreads();

    for (int tile = 0; tile < RL_TILES; ++tile) {
        const long long row0 = ((long long)blockIdx.x * RL_TILES + tile) * SM_ROWS;
        if (row0 >= B) break;
        const long long gr_[2] = {row0 + 16 * w + g, row0 + 16 * w + 8 + g};
        const bool ok[2] = {gr_[0] < B, gr_[1] < B};
        const size_t rb[2] = {((size_t)s * B + (ok[0] ? gr_[0] : 0)) * Q, ((size_t)s * B + (ok[1] ? gr_[1] : 0)) * Q};
        double cb[2] = {0.0, 0.0};
        if (BWD) {
            cb[0] = ok[0] ? cbar[(size_t)s * B + gr_[0]] : 0.0;
            cb[1] = ok[1] ? cbar[(size_t)s * B + gr_[1]] : 0.0;
        }
        double yC[2][NB][2];                 // running right-hand side -> Y -> P, C-fragment layout
#pragma unroll
        for (int b = 0; b < NB; ++b)
#pragma unroll
            for (int mb = 0; mb < 2; ++mb) {
                const int c0 = 8 * b + 2 * t;
                double v0 = 0.0, v1 = 0.0;
                if (ok[mb]) {
                    if (c0 < Q) v0 = BWD ? fma(cb[mb], K[rb[mb] + c0], Pbar[rb[mb] + c0]) : K[rb[mb] + c0];
                    if (c0 + 1 < Q) v1 = BWD ? fma(cb[mb], K[rb[mb] + c0 + 1], Pbar[rb[mb] + c0 + 1]) : K[rb[mb] + c0 + 1];
                }
                yC[mb][b][0] = v0;
                yC[mb][b][1] = v1;
            }
        // ---- forward sweep: Y R^T = rhs  (skipped for the plain right solve X = rhs R^-1) ------------------------
        if (!right_only)
#pragma unroll
        for (int b = 0; b < NB; ++b) {
            double ya[2][2];
#pragma unroll
            for (int mb = 0; mb < 2; ++mb) {
                double a0, a1;
                c_to_a(yC[mb][b][0], yC[mb][b][1], t, lane, a0, a1);
                double y0 = 0.0, y1 = 0.0;
                dmma884m(y0, y1, a0, Ri[b * 64 + g * 8 + t]);             // B[k][n] = inv(R_bb)[n][k]
                dmma884m(y0, y1, a1, Ri[b * 64 + g * 8 + 4 + t]);
                yC[mb][b][0] = y0;
                yC[mb][b][1] = y1;
                c_to_a(y0, y1, t, lane, ya[mb][0], ya[mb][1]);
            }
#pragma unroll
            for (int b2 = b + 1; b2 < NB; ++b2)
#pragma unroll
                for (int ks = 0; ks < 2; ++ks) {
                    const double bf = Rn[(8 * b2 + g) * LDR + 8 * b + 4 * ks + t];      // B[k][n] = -R[8 b2 + n][8 b + k]
                    dmma884m(yC[0][b2][0], yC[0][b2][1], ya[0][ks], bf);
                    dmma884m(yC[1][b2][0], yC[1][b2][1], ya[1][ks], bf);
                }
        }
        // ---- backward sweep: P R = Y ------------------------------------------------------------------------------
#pragma unroll
        for (int b = NB - 1; b >= 0; --b) {
            double pa[2][2];
#pragma unroll
            for (int mb = 0; mb < 2; ++mb) {
                double a0, a1;
                c_to_a(yC[mb][b][0], yC[mb][b][1], t, lane, a0, a1);
                double p0 = 0.0, p1 = 0.0;
                dmma884m(p0, p1, a0, RiT[b * 64 + g * 8 + t]);            // B[k][n] = inv(R_bb)[k][n]
                dmma884m(p0, p1, a1, RiT[b * 64 + g * 8 + 4 + t]);
                yC[mb][b][0] = p0;
                yC[mb][b][1] = p1;
                c_to_a(p0, p1, t, lane, pa[mb][0], pa[mb][1]);
            }
#pragma unroll
            for (int b2 = 0; b2 < b; ++b2)
#pragma unroll
                for (int ks = 0; ks < 2; ++ks) {
                    const double bf = Rn[(8 * b + 4 * ks + t) * LDR + 8 * b2 + g];      // B[k][n] = -R[8 b + k][8 b2 + n]
                    dmma884m(yC[0][b2][0], yC[0][b2][1], pa[0][ks], bf);
                    dmma884m(yC[1][b2][0], yC[1][b2][1], pa[1][ks], bf);
                }
        }
        // ---- outputs ----------------------------------------------------------------------------------------------
#pragma unroll
        for (int mb = 0; mb < 2; ++mb) {
            double csum = 0.0;
            if (ok[mb]) {
#pragma unroll
                for (int b = 0; b < NB; ++b) {
                    const int c0 = 8 * b + 2 * t;
                    if (!BWD) {
                        if (c0 < Q) { P[rb[mb] + c0] = yC[mb][b][0]; csum = fma(yC[mb][b][0], K[rb[mb] + c0], csum); }
                        if (c0 + 1 < Q) { P[rb[mb] + c0 + 1] = yC[mb][b][1]; csum = fma(yC[mb][b][1], K[rb[mb] + c0 + 1], csum); }
                    } else {
                        if (c0 < Q) Tout[rb[mb] + c0] = yC[mb][b][0];     // Kbar = T + cbar P is written by k_atb_mma
                        if (c0 + 1 < Q) Tout[rb[mb] + c0 + 1] = yC[mb][b][1];
                    }
                }
            }
            if (!BWD) {
                csum += __shfl_xor_sync(0xffffffffu, csum, 1);
                csum += __shfl_xor_sync(0xffffffffu, csum, 2);
                if (t == 0 && ok[mb] && cout) cout[(size_t)s * B + gr_[mb]] = csum;
            }
        }
    }
}

static thread_local int tl_right_only = 0;      // set by nmgp_right_solve_rows around its dispatch
template <int NB, bool BWD>
static int launch_solve_rl(const double* K, const double* R, double* P, double* c, const double* Pbar,
                           const double* cbar, const double* Pin, double* Kbar, double* Tout, int ns, long long B,
                           int Q, cudaStream_t st, const char* what) {
    constexpr int QP = 8 * NB, LDR = QP + 4;
    size_t smem = sizeof(double) * (QP * LDR + 2 * NB * 64);
    if (int r = nmgp_opt_in_smem(k_solve_rows_rl<NB, BWD>, smem, what)) return r;
    dim3 grid((unsigned)((B + SM_ROWS * RL_TILES - 1) / (SM_ROWS * RL_TILES)), ns);
    k_solve_rows_rl<NB, BWD><<<NMGP_L(grid), SM_THREADS, smem, st>>>(K, R, P, c, Pbar, cbar, Tout, B, Q, tl_right_only);
    return nmgp_launch_status(what);
}
#define SMM_DISPATCH(BWDFLAG, ...)                                           \
    switch ((Q + 7) / 8) {                                                   \
        case 1: return launch_solve_mma<1, BWDFLAG>(__VA_ARGS__);            \
        case 2: return launch_solve_mma<2, BWDFLAG>(__VA_ARGS__);            \
        case 3: return launch_solve_mma<3, BWDFLAG>(__VA_ARGS__);            \
        case 4: return launch_solve_mma<4, BWDFLAG>(__VA_ARGS__);            \
        case 5: return launch_solve_mma<5, BWDFLAG>(__VA_ARGS__);            \
        case 6: return launch_solve_mma<6, BWDFLAG>(__VA_ARGS__);            \
        case 7: return launch_solve_mma<7, BWDFLAG>(__VA_ARGS__);            \
        case 8: return launch_solve_mma<8, BWDFLAG>(__VA_ARGS__);            \
        case 9: return launch_solve_rl<9, BWDFLAG>(__VA_ARGS__);             \
        case 10: return launch_solve_rl<10, BWDFLAG>(__VA_ARGS__);           \
        case 11: return launch_solve_rl<11, BWDFLAG>(__VA_ARGS__);           \
        case 12: return launch_solve_rl<12, BWDFLAG>(__VA_ARGS__);           \
        case 13: return launch_solve_rl<13, BWDFLAG>(__VA_ARGS__);           \
        case 14: return launch_solve_rl<14, BWDFLAG>(__VA_ARGS__);           \
        case 15: return launch_solve_rl<15, BWDFLAG>(__VA_ARGS__);           \
        case 16: return launch_solve_rl<16, BWDFLAG>(__VA_ARGS__);           \
        default: return 1;                                                   \
    }
int nmgp_solve_rows_fwd_mma(const double* K, const double* R, double* P, double* c, int ns, long long B, int Q,
                            cudaStream_t st) {
    SMM_DISPATCH(false, K, R, P, c, nullptr, nullptr, nullptr, nullptr, nullptr, ns, B, Q, st, "nmgp_solve_rows_fwd(mma)")
}
int nmgp_solve_rows_bwd_mma(const double* Pbar, const double* cbar, const double* K, const double* P, const double* R,
                            double* Kbar, double* Tout, int ns, long long B, int Q, cudaStream_t st) {
    SMM_DISPATCH(true, K, R, nullptr, nullptr, Pbar, cbar, P, Kbar, Tout, ns, B, Q, st, "nmgp_solve_rows_bwd(mma)")
}

// X[s] = K[s] R[s]^-1 (rows of K solved against the lower-triangular R from the right; 64 < Q <= 128): the second sweep
// of the row solve alone.  Building block of the DMMA Cholesky adjoint for large Q (nmgp_potrf_bwd_batched).
int nmgp_right_solve_rows(const double* K, const double* R, double* X, int ns, long long B, int Q, cudaStream_t st) {
    if (Q <= 64 || Q > 128) return 1;
    tl_right_only = 1;
    int r = nmgp_solve_rows_fwd_mma(K, R, X, nullptr, ns, B, Q, st);
    tl_right_only = 0;
    return r;
}
