// TMA-fed FP64 tensor-core GEMM:  C[M,N] = alpha A[M,K] B[N,K]^T + beta C   (row-major; the trailing updates and panel
// products of the blocked Cholesky, kron_mv, the triangular-inverse build of the log-density adjoints).
//
// Same DMMA (mma.sync.m8n8k4.f64) warp tiling as k_gemm_nt<1> in nmgp_dense.cu -- 128 x 64 CTA tile, four consumer warps
// with 32 x 64 warp tiles, two CTAs per SM so that the read-modify-write epilogue of one overlaps the MMAs of the other
// -- but the operand tiles are no longer copied by the compute warps with cp.async:
//   * one elected lane (warp 0, lane 0) is the PRODUCER: it issues cp.async.bulk.tensor.2d (TMA, SASS UTMALDG) loads
//     of [rows x 16 doubles] boxes (128-byte rows, SWIZZLE_128B) into a 2-stage shared-memory ring one k-stage ahead
//     of the MMAs and arms the stage's `full` mbarrier with the expected byte count; every warp releases a stage
//     through its `empty` mbarrier (no CTA-wide barrier in the K loop; a dedicated producer warp would cap the
//     compute warps at 168 registers and spill the 64 accumulators);
//   * out-of-range rows / k columns are zero-filled by the TMA unit itself: no boundary code in the load path;
//   * the swizzled box layout (16-byte chunk j of row r at chunk j ^ (r & 7)) replaces the padded rows of the cp.async
//     kernel.  A-fragment rows are visited in the order 0,2,4,6,1,3,5,7 within each 8-row group, which makes the A loads
//     bank-conflict free under that swizzle; B-fragment loads keep the natural order (C columns must stay adjacent for
//     the 16-byte stores) and see 2-way conflicts, far below the shared-memory budget of a DMMA-bound loop.
// The tensor maps are encoded per call on the host (cuTensorMapEncodeTiled through cudaGetDriverEntryPoint: no link-time
// dependency on libcuda) and passed as __grid_constant__ kernel parameters.
#include <cuda.h>

#include "common.cuh"

namespace {

#define TG_M 128
#define TG_N 64
#define TG_KBOX 16                 // doubles per TMA box row (128 bytes: the SWIZZLE_128B span)
#define TG_K (2 * TG_KBOX)         // k extent of one pipeline stage
#define TG_STAGES 2
#define TG_CONSUMERS 4
#define TG_THREADS (32 * TG_CONSUMERS)
#define TG_STAGE_DOUBLES ((TG_M + TG_N) * TG_K)
#define TG_SMEM (TG_STAGES * TG_STAGE_DOUBLES * 8 + 1024 + 64)

__device__ __forceinline__ void dmma884t(double& d0, double& d1, double a, double b) {
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n"
                 : "+d"(d0), "+d"(d1)
                 : "d"(a), "d"(b));
}
__device__ __forceinline__ unsigned su32(const void* p) { return (unsigned)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mb_init(unsigned long long* b, int count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;\n" ::"r"(su32(b)), "r"(count) : "memory");
}
__device__ __forceinline__ void mb_expect_tx(unsigned long long* b, unsigned bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;\n" ::"r"(su32(b)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mb_arrive(unsigned long long* b) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];\n" ::"r"(su32(b)) : "memory");
}
__device__ __forceinline__ void mb_wait(unsigned long long* b, unsigned parity) {
    unsigned done = 0;
    const unsigned a = su32(b);
    while (!done) {
        asm volatile(
            "{\n\t.reg .pred P1;\n\tmbarrier.try_wait.parity.shared::cta.b64 P1, [%1], %2;\n\tselp.b32 %0, 1, 0, P1;\n\t}"
            : "=r"(done)
            : "r"(a), "r"(parity)
            : "memory");
    }
}
// 2-D tiled TMA load: box at (c0 = innermost coordinate (k), c1 = row) of the tensor described by `map`
__device__ __forceinline__ void tma_load_2d(void* dst, const CUtensorMap* map, int c0, int c1, unsigned long long* bar) {
    asm volatile(
        "cp.async.bulk.tensor.2d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];\n" ::"r"(
            su32(dst)),
        "l"(map), "r"(su32(bar)), "r"(c0), "r"(c1)
        : "memory");
}
// element (r, c) of a [rows][16 doubles] box written with SWIZZLE_128B (box base 1024-byte aligned)
__device__ __forceinline__ int swz(int r, int c) { return r * TG_KBOX + ((((c >> 1) ^ (r & 7)) << 1) | (c & 1)); }
// fragment row g -> row inside its 8-row group (A operand): 0,2,4,6,1,3,5,7
__device__ __forceinline__ int prow(int g) { return g < 4 ? 2 * g : 2 * (g - 4) + 1; }

__global__ void __launch_bounds__(TG_THREADS, 2)
k_gemm_nt_tma(const __grid_constant__ CUtensorMap mapA, const __grid_constant__ CUtensorMap mapB, double* __restrict__ C,
              long long M, long long N, long long K, long long ldc, double alpha, double beta, int lower_only) {
    extern __shared__ unsigned char tg_raw[];
    // 1024-byte aligned stage buffers (required by the 128-byte swizzle), barriers behind them
    double* stage0 = reinterpret_cast<double*>((reinterpret_cast<uintptr_t>(tg_raw) + 1023) & ~uintptr_t(1023));
    unsigned long long* full = reinterpret_cast<unsigned long long*>(stage0 + TG_STAGES * TG_STAGE_DOUBLES);
    unsigned long long* empty = full + TG_STAGES;
    const long long m0 = (long long)blockIdx.y * TG_M, n0 = (long long)blockIdx.x * TG_N;
    if ((lower_only & 1) && n0 > m0 + TG_M - 1) return;
    const int tid = threadIdx.x, lane = tid & 31, w = tid >> 5, g = lane >> 2, t = lane & 3;
    if (tid == 0) {
        for (int s = 0; s < TG_STAGES; ++s) {
            mb_init(&full[s], 1);
            mb_init(&empty[s], TG_CONSUMERS);
        }
        asm volatile("fence.mbarrier_init.release.cluster;\n" ::: "memory");
    }
    __syncthreads();
    const int nk = (int)((K + TG_K - 1) / TG_K);

    auto issue = [&](int kt) {                      // lane 0 of warp 0: TMA loads of k-stage kt
        const int s = kt % TG_STAGES;
        if (kt >= TG_STAGES) mb_wait(&empty[s], (unsigned)(((kt / TG_STAGES) - 1) & 1));
        double* As = stage0 + (size_t)s * TG_STAGE_DOUBLES;          // [2 boxes][128 rows][16]
        double* Bs = As + 2 * TG_M * TG_KBOX;                        // [2 boxes][64 rows][16]
        mb_expect_tx(&full[s], (unsigned)(TG_STAGE_DOUBLES * sizeof(double)));
        const int k0 = kt * TG_K;
        tma_load_2d(As, &mapA, k0, (int)m0, &full[s]);
        tma_load_2d(As + TG_M * TG_KBOX, &mapA, k0 + TG_KBOX, (int)m0, &full[s]);
        tma_load_2d(Bs, &mapB, k0, (int)n0, &full[s]);
        tma_load_2d(Bs + TG_N * TG_KBOX, &mapB, k0 + TG_KBOX, (int)n0, &full[s]);
    };
    if (tid == 0) {
        if (lower_only >= 0) {      // (bit 1 of the flags switches the descriptor prefetch off: experiment knob)
            if (!(lower_only & 2)) {
                asm volatile("prefetch.tensormap [%0];\n" ::"l"(&mapA) : "memory");
                asm volatile("prefetch.tensormap [%0];\n" ::"l"(&mapB) : "memory");
            }
        }
        issue(0);
    }

    // ===== 32 x 64 warp tiles (rows 32 w .. 32 w + 31 of the CTA tile) =====
    if (beta != 0.0) {   // pull the C tile towards L2 while the K loop runs
        for (int e = tid; e < TG_M * (TG_N / 16); e += 32 * TG_CONSUMERS) {
            const long long r = m0 + e / (TG_N / 16), c = n0 + (e % (TG_N / 16)) * 16;
            if (r < M && c < N) asm volatile("prefetch.global.L2 [%0];" ::"l"(&C[r * ldc + c]));
        }
    }
    double acc[4][8][2];
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 8; ++j) acc[i][j][0] = acc[i][j][1] = 0.0;
    const int pg = prow(g);
    for (int kt = 0; kt < nk; ++kt) {
        const int s = kt % TG_STAGES;
        if (tid == 0 && kt + 1 < nk) issue(kt + 1);       // the other stage: free once every warp released k-stage kt - 1
        mb_wait(&full[s], (unsigned)((kt / TG_STAGES) & 1));
        const double* As = stage0 + (size_t)s * TG_STAGE_DOUBLES;
        const double* Bs = As + 2 * TG_M * TG_KBOX;
#pragma unroll
        for (int box = 0; box < 2; ++box) {
            const double* Ab = As + box * TG_M * TG_KBOX;
            const double* Bb = Bs + box * TG_N * TG_KBOX;
#pragma unroll
            for (int ks = 0; ks < TG_KBOX / 4; ++ks) {
                double af[4], bf[8];
#pragma unroll
                for (int i = 0; i < 4; ++i) af[i] = Ab[swz(32 * w + 8 * i + pg, 4 * ks + t)];   // A[row][k]
#pragma unroll
                for (int j = 0; j < 8; ++j) bf[j] = Bb[swz(8 * j + g, 4 * ks + t)];             // B^T[k][col] = B[col][k]
#pragma unroll
                for (int i = 0; i < 4; ++i)
#pragma unroll
                    for (int j = 0; j < 8; ++j) dmma884t(acc[i][j][0], acc[i][j][1], af[i], bf[j]);
            }
        }
        __syncwarp();
        if (lane == 0) mb_arrive(&empty[s]);
    }
    // epilogue: all of the thread's C values are fetched as independent 16-byte loads before the first store
    const bool c_vec = ((ldc & 1) == 0) && ((((size_t)C) & 15) == 0);
#pragma unroll
    for (int ih = 0; ih < 4; ih += 2) {
        double2 cv[2][8];
        if (beta != 0.0) {
#pragma unroll
            for (int i = 0; i < 2; ++i) {
                const long long r = m0 + 32 * w + 8 * (ih + i) + pg;
#pragma unroll
                for (int j = 0; j < 8; ++j) {
                    const long long c = n0 + 8 * j + 2 * t;
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
            const long long r = m0 + 32 * w + 8 * (ih + i) + pg;
            if (r >= M) continue;
#pragma unroll
            for (int j = 0; j < 8; ++j) {
                const long long c = n0 + 8 * j + 2 * t;
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

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
EncodeTiledFn g_encode = nullptr;
int g_encode_state = 0;          // 0: not looked up, 1: available, -1: unavailable

bool encode_map(CUtensorMap* map, const double* base, long long rows, long long cols, long long ld, int box_rows) {
    cuuint64_t dims[2] = {(cuuint64_t)cols, (cuuint64_t)rows};
    cuuint64_t strides[1] = {(cuuint64_t)ld * sizeof(double)};
    cuuint32_t box[2] = {TG_KBOX, (cuuint32_t)box_rows};
    cuuint32_t estr[2] = {1, 1};
    static const int promo = [] { const char* e = getenv("NMGP_TMA_L2PROMO"); return e ? atoi(e) : 256; }();
    return g_encode(map, CU_TENSOR_MAP_DATA_TYPE_FLOAT64, 2, const_cast<double*>(base), dims, strides, box, estr,
                    CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
                    promo == 0 ? CU_TENSOR_MAP_L2_PROMOTION_NONE : (promo == 128 ? CU_TENSOR_MAP_L2_PROMOTION_L2_128B
                                                                                 : CU_TENSOR_MAP_L2_PROMOTION_L2_256B),
                    CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
}

bool ensure_encode() {
    if (g_encode_state == 0) {
        void* fn = nullptr;
        cudaDriverEntryPointQueryResult qres;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &qres) == cudaSuccess && fn &&
            qres == cudaDriverEntryPointSuccess) {
            g_encode = reinterpret_cast<EncodeTiledFn>(fn);
            g_encode_state = 1;
        } else {
            g_encode_state = -1;
        }
    }
    return g_encode_state > 0;
}

}  // namespace

// returns 0 when the GEMM was launched, 1 when this path does not apply (caller falls back to the cp.async kernel),
// < 0 on error.  Requirements of the TMA descriptors: 16-byte aligned base pointers and row strides.
int nmgp_gemm_nt_tma(const double* A, const double* Bm, double* C, long long M, long long N, long long K, long long lda,
                     long long ldb, long long ldc, double alpha, double beta, int lower_only, cudaStream_t st) {
    static const bool off = [] { const char* e = getenv("NMGP_GEMM_TMA"); return e && e[0] == '0'; }();
    if (off || !ensure_encode()) return 1;
    if ((lda & 1) || (ldb & 1) || (((size_t)A) & 15) || (((size_t)Bm) & 15) || K < 1 || M >= (1LL << 31) ||
        N >= (1LL << 31) || K >= (1LL << 31))
        return 1;
    CUtensorMap mapA, mapB;
    if (!encode_map(&mapA, A, M, K, lda, TG_M) || !encode_map(&mapB, Bm, N, K, ldb, TG_N)) return 1;
    if (int r = nmgp_opt_in_smem(k_gemm_nt_tma, TG_SMEM, "nmgp_gemm_nt(tma)")) return r;
    dim3 grid((unsigned)((N + TG_N - 1) / TG_N), (unsigned)((M + TG_M - 1) / TG_M));
    static const int noprefetch = [] { const char* e = getenv("NMGP_TMA_NOPREFETCH"); return e && e[0] == '1' ? 2 : 0; }();
    k_gemm_nt_tma<<<NMGP_L(grid), TG_THREADS, TG_SMEM, st>>>(mapA, mapB, C, M, N, K, ldc, alpha, beta,
                                                              (lower_only ? 1 : 0) | noprefetch);
    return nmgp_launch_status("nmgp_gemm_nt(tma)");
}
