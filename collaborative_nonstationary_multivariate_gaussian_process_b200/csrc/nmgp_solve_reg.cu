// Register-resident row solves P = K (R R^T)^-1 for Q <= 64 and the DMMA reduction Abar -= T^T P of their adjoint.
//
// k_solve_rows_reg: one thread per row, the row (QP doubles) lives in registers through both substitutions; the
//   Cholesky factor R (and R^T, so that both sweeps read contiguous, 16-byte aligned pairs) is read from shared
//   memory as warp-wide broadcasts; rows enter and leave through a shared tile so that global traffic is coalesced.
//   Replaces the LU solves of code/utils.py:119,142,154,230 (torch.solve of K22 + eps I) and their autograd.
// k_atb_mma: C[s] += sign * A[s]^T B[s] over the rows of a chunk (DMMA m8n8k4), one atomicAdd pass per chunk.
#include "common.cuh"

__device__ __forceinline__ void dmma884s(double& d0, double& d1, double a, double b) {
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n"
                 : "+d"(d0), "+d"(d1)
                 : "d"(a), "d"(b));
}
__device__ __forceinline__ void cpa8(double* smem_dst, const double* gsrc) {
    unsigned sa = (unsigned)__cvta_generic_to_shared(smem_dst);
    asm volatile("cp.async.ca.shared.global [%0], [%1], 8;\n" ::"r"(sa), "l"(gsrc));
}
__device__ __forceinline__ void cpa_commit() { asm volatile("cp.async.commit_group;\n" ::); }
__device__ __forceinline__ void cpa_wait0() { asm volatile("cp.async.wait_group 0;\n" ::); }

template <int QP, bool BWD>
__global__ void __launch_bounds__(128)
k_solve_rows_reg(const double* __restrict__ K, const double* __restrict__ R, double* __restrict__ P,
                 double* __restrict__ c, const double* __restrict__ Pbar, const double* __restrict__ cbar,
                 const double* __restrict__ Pin, double* __restrict__ Kbar, double* __restrict__ Tout, long long B,
                 int Q) {
    extern __shared__ __align__(16) double sm[];
    constexpr int LDT = QP + 1;
    const int TR = blockDim.x;
    double* Rs = sm;                       // [QP][QP]   R (lower), identity padding
    double* RTs = Rs + QP * QP;            // [QP][QP]   R^T
    double* rinv = RTs + QP * QP;          // [QP]
    double* tile = rinv + QP;              // [TR][LDT]
    const int s = blockIdx.y, tid = threadIdx.x;
    const long long row0 = (long long)blockIdx.x * TR;
    const int nrows = (int)min((long long)TR, B - row0);
    const double* Rg = R + (size_t)s * Q * Q;
    for (int e = tid; e < QP * QP; e += TR) {
        int a = e / QP, b = e - a * QP;
        double v = (a < Q && b < Q) ? ((b <= a) ? Rg[(size_t)a * Q + b] : 0.0) : (a == b ? 1.0 : 0.0);
        Rs[e] = v;
        RTs[b * QP + a] = v;
    }
    for (int a = tid; a < QP; a += TR) rinv[a] = a < Q ? 1.0 / Rg[(size_t)a * Q + a] : 1.0;
    const size_t base = ((size_t)s * B + row0) * Q;
    for (int e = tid; e < TR * QP; e += TR) {
        int r = e / QP, a = e - r * QP;
        double v = 0.0;
        if (r < nrows && a < Q) {
            size_t o = base + (size_t)r * Q + a;
            v = BWD ? fma(cbar[(size_t)s * B + row0 + r], K[o], Pbar[o]) : K[o];
        }
        tile[r * LDT + a] = v;
    }
    __syncthreads();
    double yv[QP];
#pragma unroll
    for (int a = 0; a < QP; ++a) yv[a] = tile[tid * LDT + a];
    // forward: R y = k
#pragma unroll
    for (int a = 0; a < QP; ++a) {
        double s0 = yv[a], s1 = 0.0, s2 = 0.0, s3 = 0.0;
#pragma unroll
        for (int cidx = 0; cidx < a; ++cidx) {
            const double rv = Rs[a * QP + cidx];
            if ((cidx & 3) == 0) s0 = fma(-rv, yv[cidx], s0);
            else if ((cidx & 3) == 1) s1 = fma(-rv, yv[cidx], s1);
            else if ((cidx & 3) == 2) s2 = fma(-rv, yv[cidx], s2);
            else s3 = fma(-rv, yv[cidx], s3);
        }
        yv[a] = ((s0 + s1) + (s2 + s3)) * rinv[a];
    }
    // backward: R^T p = y
#pragma unroll
    for (int a = QP - 1; a >= 0; --a) {
        double s0 = yv[a], s1 = 0.0, s2 = 0.0, s3 = 0.0;
#pragma unroll
        for (int cidx = a + 1; cidx < QP; ++cidx) {
            const double rv = RTs[a * QP + cidx];
            if ((cidx & 3) == 0) s0 = fma(-rv, yv[cidx], s0);
            else if ((cidx & 3) == 1) s1 = fma(-rv, yv[cidx], s1);
            else if ((cidx & 3) == 2) s2 = fma(-rv, yv[cidx], s2);
            else s3 = fma(-rv, yv[cidx], s3);
        }
        yv[a] = ((s0 + s1) + (s2 + s3)) * rinv[a];
    }
    if (!BWD) {
        asm volatile("" ::: "memory");   // keep the re-read of k below from being hoisted above the sweeps
        double acc0 = 0.0, acc1 = 0.0;
#pragma unroll
        for (int a = 0; a < QP; a += 2) {
            acc0 = fma(yv[a], tile[tid * LDT + a], acc0);
            acc1 = fma(yv[a + 1], tile[tid * LDT + a + 1], acc1);
        }
        if (tid < nrows) c[(size_t)s * B + row0 + tid] = acc0 + acc1;
    }
#pragma unroll
    for (int a = 0; a < QP; ++a) tile[tid * LDT + a] = yv[a];
    __syncthreads();
    for (int e = tid; e < nrows * Q; e += TR) {
        int r = e / Q, a = e - r * Q;
        double v = tile[r * LDT + a];
        if (!BWD) {
            P[base + e] = v;
        } else {
            Tout[base + e] = v;
            Kbar[base + e] = fma(cbar[(size_t)s * B + row0 + r], Pin[base + e], v);
        }
    }
}

template <int QP, bool BWD>
static int launch_solve_reg(const double* K, const double* R, double* P, double* c, const double* Pbar,
                            const double* cbar, const double* Pin, double* Kbar, double* Tout, int ns, long long B,
                            int Q, cudaStream_t st, const char* what) {
    int TR = 128;
    size_t smem = sizeof(double) * (2 * QP * QP + QP + (size_t)TR * (QP + 1));
    if (smem > 113 * 1024) {               // keep two CTAs per SM
        TR = 64;
        smem = sizeof(double) * (2 * QP * QP + QP + (size_t)TR * (QP + 1));
    }
    if (int r = nmgp_opt_in_smem(k_solve_rows_reg<QP, BWD>, smem, what)) return r;
    dim3 grid((unsigned)((B + TR - 1) / TR), ns);
    k_solve_rows_reg<QP, BWD><<<grid, TR, smem, st>>>(K, R, P, c, Pbar, cbar, Pin, Kbar, Tout, B, Q);
    return nmgp_launch_status(what);
}

#define SR_DISPATCH(BWDFLAG, ...)                                                        \
    switch ((Q + 7) / 8) {                                                               \
        case 1: return launch_solve_reg<8, BWDFLAG>(__VA_ARGS__);                        \
        case 2: return launch_solve_reg<16, BWDFLAG>(__VA_ARGS__);                       \
        case 3: return launch_solve_reg<24, BWDFLAG>(__VA_ARGS__);                       \
        case 4: return launch_solve_reg<32, BWDFLAG>(__VA_ARGS__);                       \
        case 5: return launch_solve_reg<40, BWDFLAG>(__VA_ARGS__);                       \
        case 6: return launch_solve_reg<48, BWDFLAG>(__VA_ARGS__);                       \
        case 7: return launch_solve_reg<56, BWDFLAG>(__VA_ARGS__);                       \
        case 8: return launch_solve_reg<64, BWDFLAG>(__VA_ARGS__);                       \
        default: return 1;                                                               \
    }

int nmgp_solve_rows_fwd_reg(const double* K, const double* R, double* P, double* c, int ns, long long B, int Q,
                            cudaStream_t st) {
    SR_DISPATCH(false, K, R, P, c, nullptr, nullptr, nullptr, nullptr, nullptr, ns, B, Q, st, "nmgp_solve_rows_fwd(reg)")
}
int nmgp_solve_rows_bwd_reg(const double* Pbar, const double* cbar, const double* K, const double* P, const double* R,
                            double* Kbar, double* Tout, int ns, long long B, int Q, cudaStream_t st) {
    SR_DISPATCH(true, K, R, nullptr, nullptr, Pbar, cbar, P, Kbar, Tout, ns, B, Q, st, "nmgp_solve_rows_bwd(reg)")
}

// ------------------------------------------------------------------------------------------------------------
// C[s] += sign * sum_n A[s,n,:]^T B[s,n,:]   (Q x Q, rows n of one chunk per CTA), DMMA.
#define ATB_TROWS 32
#define ATB_CHUNK 2048
#define ATB_THREADS 128
__host__ __device__ constexpr int atb_pad(int n) { return ((n + 3) / 8) * 8 + 4; }

template <int NB>
__global__ void __launch_bounds__(ATB_THREADS)
k_atb_mma(const double* __restrict__ A, const double* __restrict__ Bm, double* __restrict__ C, double sign,
          long long B, int Q) {
    constexpr int LDP = atb_pad(8 * NB);
    extern __shared__ __align__(16) double sm[];
    double* At = sm;                                   // [2][ATB_TROWS][LDP]
    double* Bt = At + 2 * ATB_TROWS * LDP;             // [2][ATB_TROWS][LDP]
    const int s = blockIdx.y, tid = threadIdx.x, lane = tid & 31, w = tid >> 5, g = lane >> 2, t = lane & 3;
    const long long rbeg = (long long)blockIdx.x * ATB_CHUNK, rend = min(B, rbeg + ATB_CHUNK);
    for (int e = tid; e < 4 * ATB_TROWS * LDP; e += ATB_THREADS) sm[e] = 0.0;
    __syncthreads();
    const int a1 = w, a2 = w + 4;                      // strips of 8 rows of C handled by this warp
    const bool on1 = a1 < NB, on2 = a2 < NB;
    double acc[2][NB][2];
#pragma unroll
    for (int h = 0; h < 2; ++h)
#pragma unroll
        for (int nb = 0; nb < NB; ++nb) acc[h][nb][0] = acc[h][nb][1] = 0.0;
    const long long ntiles = (rend - rbeg + ATB_TROWS - 1) / ATB_TROWS;
    auto stage = [&](long long tile, int buf) {
        const long long r0 = rbeg + tile * ATB_TROWS;
        const int nr = (int)min((long long)ATB_TROWS, rend - r0);
        double* Ad = At + buf * ATB_TROWS * LDP;
        double* Bd = Bt + buf * ATB_TROWS * LDP;
        for (int e = tid; e < ATB_TROWS * Q; e += ATB_THREADS) {
            int r = e / Q, cc = e - r * Q;
            if (r < nr) {
                size_t o = ((size_t)s * B + r0 + r) * Q + cc;
                cpa8(&Ad[r * LDP + cc], &A[o]);
                cpa8(&Bd[r * LDP + cc], &Bm[o]);
            } else {
                Ad[r * LDP + cc] = 0.0;
                Bd[r * LDP + cc] = 0.0;
            }
        }
    };
    stage(0, 0);
    cpa_commit();
    for (long long tile = 0; tile < ntiles; ++tile) {
        const int buf = (int)(tile & 1);
        cpa_wait0();
        __syncthreads();
        if (tile + 1 < ntiles) stage(tile + 1, buf ^ 1);
        cpa_commit();
        if (!on1) continue;
        const double* Ad = At + buf * ATB_TROWS * LDP;
        const double* Bd = Bt + buf * ATB_TROWS * LDP;
#pragma unroll
        for (int kk = 0; kk < ATB_TROWS / 4; ++kk) {
            const int n = 4 * kk + t;
            const double fa1 = Ad[n * LDP + 8 * a1 + g];
            const double fa2 = on2 ? Ad[n * LDP + 8 * a2 + g] : 0.0;
#pragma unroll
            for (int nb = 0; nb < NB; ++nb) {
                const double fb = Bd[n * LDP + 8 * nb + g];
                dmma884s(acc[0][nb][0], acc[0][nb][1], fa1, fb);
                if (on2) dmma884s(acc[1][nb][0], acc[1][nb][1], fa2, fb);
            }
        }
    }
    cpa_wait0();
    if (!on1) return;
    double* Cs = C + (size_t)s * Q * Q;
#pragma unroll
    for (int h = 0; h < 2; ++h) {
        if (h == 1 && !on2) break;
        const int r = 8 * (h == 0 ? a1 : a2) + g;
#pragma unroll
        for (int nb = 0; nb < NB; ++nb)
#pragma unroll
            for (int e = 0; e < 2; ++e) {
                const int cc = 8 * nb + 2 * t + e;
                if (r < Q && cc < Q) atomicAdd(&Cs[(size_t)r * Q + cc], sign * acc[h][nb][e]);
            }
    }
}

template <int NB>
static int launch_atb(const double* A, const double* Bm, double* C, double sign, int ns, long long B, int Q,
                      cudaStream_t st) {
    size_t smem = sizeof(double) * 4 * ATB_TROWS * atb_pad(8 * NB);
    if (int r = nmgp_opt_in_smem(k_atb_mma<NB>, smem, "nmgp_atb")) return r;
    dim3 grid((unsigned)((B + ATB_CHUNK - 1) / ATB_CHUNK), ns);
    k_atb_mma<NB><<<grid, ATB_THREADS, smem, st>>>(A, Bm, C, sign, B, Q);
    return nmgp_launch_status("nmgp_atb");
}
int nmgp_atb_mma(const double* A, const double* Bm, double* C, double sign, int ns, long long B, int Q, cudaStream_t st) {
    switch ((Q + 7) / 8) {
        case 1: return launch_atb<1>(A, Bm, C, sign, ns, B, Q, st);
        case 2: return launch_atb<2>(A, Bm, C, sign, ns, B, Q, st);
        case 3: return launch_atb<3>(A, Bm, C, sign, ns, B, Q, st);
        case 4: return launch_atb<4>(A, Bm, C, sign, ns, B, Q, st);
        case 5: return launch_atb<5>(A, Bm, C, sign, ns, B, Q, st);
        case 6: return launch_atb<6>(A, Bm, C, sign, ns, B, Q, st);
        case 7: return launch_atb<7>(A, Bm, C, sign, ns, B, Q, st);
        case 8: return launch_atb<8>(A, Bm, C, sign, ns, B, Q, st);
        default: return 1;
    }
}
