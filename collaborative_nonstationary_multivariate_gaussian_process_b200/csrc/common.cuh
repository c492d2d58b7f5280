// Shared helpers for the NMGP B200 kernels (sm_100a, FP64).
#pragma once
#include <cuda_runtime.h>
#include <math.h>
#include <stdint.h>
#include <stdio.h>

#define NMGP_API extern "C" __attribute__((visibility("default")))
#define NMGP_EPS 1e-4 /* tridiagonal_jitter, reference code/utils.py:7 */

// hyper-parameter slots (exp of the reference's *_log parameters, code/nmgp_dsvi.py:180-188)
enum { H_S2_ELL = 0, H_LEN_ELL, H_S2_L0, H_LEN_L0, H_S2_L1, H_LEN_L1, H_S2_ERR, H_COUNT };
enum { MODE_W = 0, MODE_U = 1 };

void nmgp_set_error(const char* fmt, ...);
// every kernel launch of the library is counted (bench.py reports the count as gpu_launches): the first launch-
// configuration argument is wrapped as <<<NMGP_L(grid), ...>>>
extern unsigned long long g_nmgp_launches;
#define NMGP_L(grid) (++g_nmgp_launches, (grid))
int nmgp_launch_status(const char* what);

#define NMGP_REQUIRE(cond, what)                                  \
    do {                                                          \
        if (!(cond)) {                                            \
            nmgp_set_error("%s: invalid argument (%s)", what, #cond); \
            return -1;                                            \
        }                                                         \
    } while (0)

template <typename K>
static inline int nmgp_opt_in_smem(K kernel, size_t bytes, const char* what) {
    if (bytes > 227 * 1024) {
        nmgp_set_error("%s: needs %zu bytes of shared memory (> 227 KB); Q too large", what, bytes);
        return -2;
    }
    if (bytes > 48 * 1024) {
        cudaError_t e = cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes);
        if (e != cudaSuccess) {
            nmgp_set_error("%s: cudaFuncSetAttribute: %s", what, cudaGetErrorString(e));
            return -3;
        }
    }
    return 0;
}

__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

// Sum over the block; result valid in thread 0 (and broadcast to all).  blockDim.x multiple of 32, <= 1024.
__device__ __forceinline__ double block_sum(double v) {
    __shared__ double red[32];
    __shared__ double total;
    int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    v = warp_sum(v);
    __syncthreads();
    if (lane == 0) red[w] = v;
    __syncthreads();
    if (w == 0) {
        int nw = (blockDim.x + 31) >> 5;
        double t = lane < nw ? red[lane] : 0.0;
        t = warp_sum(t);
        if (lane == 0) total = t;
    }
    __syncthreads();
    return total;
}

// Division of a 32-bit index by a launch-invariant divisor in three instructions (Granlund-Montgomery): the staging
// loops turn flat element indices into (row, column) and a hardware-emulated integer division there costs more issue
// slots than the copy itself.
struct FastDiv {
    unsigned int d, m, s;
    __host__ explicit FastDiv(unsigned int div = 1) : d(div) {
        s = 0;
        while ((1ull << s) < div) ++s;
        m = (unsigned int)(((1ull << 32) * ((1ull << s) - div)) / div + 1);
    }
    __device__ __forceinline__ unsigned int div(unsigned int n) const {
        const unsigned int t = __umulhi(m, n);
        return (t + n) >> s;   // valid while t + n does not overflow: n < 2^31 (all uses are tile-local indices)
    }
};

// slot of the packed coefficient pair (i, j<=i): diagonal pairs first, then strictly-lower row-major
__device__ __forceinline__ int pair_slot(int i, int j, int D) { return i == j ? i : D + (i * (i - 1)) / 2 + j; }
