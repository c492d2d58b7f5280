// DMMA reduction Abar -= T^T P of the row-solve adjoint (the solves themselves: nmgp_solve_mma.cu).
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
__device__ __forceinline__ void cpa16(double* smem_dst, const double* gsrc) {
    unsigned sa = (unsigned)__cvta_generic_to_shared(smem_dst);
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;\n" ::"r"(sa), "l"(gsrc));
}
__device__ __forceinline__ void cpa_commit() { asm volatile("cp.async.commit_group;\n" ::); }
__device__ __forceinline__ void cpa_wait0() { asm volatile("cp.async.wait_group 0;\n" ::); }

// ------------------------------------------------------------------------------------------------------------
// C[s] += sign * sum_n A[s,n,:]^T B[s,n,:]   (Q x Q, rows n of one chunk per CTA), DMMA.
#define ATB_TROWS 32
#define ATB_CHUNK 2048
#define ATB_THREADS 256   // 8 warps: warp & 3 = strips of C it owns, warp >> 2 = which half of a tile's rows it sums
__host__ __device__ constexpr int atb_pad(int n) { return ((n + 3) / 8) * 8 + 4; }

// WIDE (8 < NB <= 16): the 8 warps are 8 strip roles (strips w and w + 8 of C) and every warp sums all rows of a tile;
// otherwise 4 strip roles x 2 row halves.
template <int NB, bool WIDE>
__global__ void __launch_bounds__(ATB_THREADS, WIDE ? 1 : 3)
k_atb_mma(const double* __restrict__ A, const double* __restrict__ Bm, double* __restrict__ C, double sign,
          long long B, int Q, long long chunk, const double* __restrict__ cbar, double* __restrict__ Kbar) {
    constexpr int LDP = atb_pad(8 * NB);
    extern __shared__ __align__(16) double sm[];
    double* At = sm;                                   // [2][ATB_TROWS][LDP]
    double* Bt = At + 2 * ATB_TROWS * LDP;             // [2][ATB_TROWS][LDP]
    const int s = blockIdx.y, tid = threadIdx.x, lane = tid & 31, w = tid >> 5, g = lane >> 2, t = lane & 3;
    const int wr = WIDE ? w : (w & 3), half = WIDE ? 0 : (w >> 2);
    constexpr int NSTRIP = WIDE ? 8 : 4, KQ = WIDE ? ATB_TROWS / 4 : ATB_TROWS / 8;
    const long long rbeg = (long long)blockIdx.x * chunk, rend = min(B, rbeg + chunk);
    if (rbeg >= rend) return;
    for (int e = tid; e < 4 * ATB_TROWS * LDP; e += ATB_THREADS) sm[e] = 0.0;
    __syncthreads();
    const int a1 = wr, a2 = wr + NSTRIP;               // strips of 8 rows of C handled by this warp
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
        for (int r = w; r < ATB_TROWS; r += ATB_THREADS / 32) {
            if (r < nr) {
                const size_t o = ((size_t)s * B + r0 + r) * Q;
                if ((Q & 1) == 0) {
                    for (int cc = 2 * lane; cc < Q; cc += 64) {
                        cpa16(&Ad[r * LDP + cc], &A[o + cc]);
                        cpa16(&Bd[r * LDP + cc], &Bm[o + cc]);
                    }
                } else {
                    for (int cc = lane; cc < Q; cc += 32) {
                        cpa8(&Ad[r * LDP + cc], &A[o + cc]);
                        cpa8(&Bd[r * LDP + cc], &Bm[o + cc]);
                    }
                }
            } else {
                for (int cc = lane; cc < Q; cc += 32) {
                    Ad[r * LDP + cc] = 0.0;
                    Bd[r * LDP + cc] = 0.0;
                }
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
        const double* Ad = At + buf * ATB_TROWS * LDP;
        const double* Bd = Bt + buf * ATB_TROWS * LDP;
        if (Kbar) {   // adjoint of the solve's right-hand side, Kbar = T + cbar P, written coalesced from the staged tiles
            const long long r0 = rbeg + tile * ATB_TROWS;
            const int nr = (int)min((long long)ATB_TROWS, rend - r0);
            for (int r = w; r < nr; r += ATB_THREADS / 32) {
                const double cb = cbar[(size_t)s * B + r0 + r];
                const size_t o = ((size_t)s * B + r0 + r) * Q;
                for (int cc = lane; cc < Q; cc += 32) Kbar[o + cc] = fma(cb, Bd[r * LDP + cc], Ad[r * LDP + cc]);
            }
        }
        if (!on1) continue;
#pragma unroll
        for (int kq = 0; kq < KQ; ++kq) {
            const int n = 4 * (kq + half * KQ) + t;
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

template <int NB, bool WIDE = (NB > 8)>
static int launch_atb(const double* A, const double* Bm, double* C, double sign, int ns, long long B, int Q,
                      const double* cbar, double* Kbar, cudaStream_t st) {
    size_t smem = sizeof(double) * 4 * ATB_TROWS * atb_pad(8 * NB);
    if (int r = nmgp_opt_in_smem(k_atb_mma<NB, WIDE>, smem, "nmgp_atb")) return r;
    // one wave: 148 SMs x 3 resident CTAs (61 KB of shared memory each), split evenly over the ns samples
    long long nchunks = ((WIDE ? 296 : 444) + ns - 1) / ns;
    long long chunk = ((B + nchunks - 1) / nchunks + ATB_TROWS - 1) / ATB_TROWS * ATB_TROWS;
    if (chunk < 4 * ATB_TROWS) chunk = 4 * ATB_TROWS;
    dim3 grid((unsigned)((B + chunk - 1) / chunk), ns);
    k_atb_mma<NB, WIDE><<<NMGP_L(grid), ATB_THREADS, smem, st>>>(A, Bm, C, sign, B, Q, chunk, cbar, Kbar);
    return nmgp_launch_status("nmgp_atb");
}
int nmgp_atb_mma(const double* A, const double* Bm, double* C, double sign, int ns, long long B, int Q,
                 const double* cbar, double* Kbar, cudaStream_t st) {
    switch ((Q + 7) / 8) {
        case 1: return launch_atb<1>(A, Bm, C, sign, ns, B, Q, cbar, Kbar, st);
        case 2: return launch_atb<2>(A, Bm, C, sign, ns, B, Q, cbar, Kbar, st);
        case 3: return launch_atb<3>(A, Bm, C, sign, ns, B, Q, cbar, Kbar, st);
        case 4: return launch_atb<4>(A, Bm, C, sign, ns, B, Q, cbar, Kbar, st);
        case 5: return launch_atb<5>(A, Bm, C, sign, ns, B, Q, cbar, Kbar, st);
        case 6: return launch_atb<6>(A, Bm, C, sign, ns, B, Q, cbar, Kbar, st);
        case 7: return launch_atb<7>(A, Bm, C, sign, ns, B, Q, cbar, Kbar, st);
        case 8: return launch_atb<8>(A, Bm, C, sign, ns, B, Q, cbar, Kbar, st);
        case 9: return launch_atb<9>(A, Bm, C, sign, ns, B, Q, cbar, Kbar, st);
        case 10: return launch_atb<10>(A, Bm, C, sign, ns, B, Q, cbar, Kbar, st);
        case 11: return launch_atb<11>(A, Bm, C, sign, ns, B, Q, cbar, Kbar, st);
        case 12: return launch_atb<12>(A, Bm, C, sign, ns, B, Q, cbar, Kbar, st);
        case 13: return launch_atb<13>(A, Bm, C, sign, ns, B, Q, cbar, Kbar, st);
        case 14: return launch_atb<14>(A, Bm, C, sign, ns, B, Q, cbar, Kbar, st);
        case 15: return launch_atb<15>(A, Bm, C, sign, ns, B, Q, cbar, Kbar, st);
        case 16: return launch_atb<16>(A, Bm, C, sign, ns, B, Q, cbar, Kbar, st);
        default: return 1;
    }
}
