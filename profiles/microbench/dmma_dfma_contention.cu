// Does a warp's FP64 FMA/ADD stream make progress while another warp of the same scheduler streams DMMA?
// Warps 0-3 (one per scheduler) run a DMMA loop; warps 4-7 run a dependent DFMA chain of fixed length.
// Reported: cycles the DFMA warps need alone vs next to the DMMA stream, and the DMMA warps' slowdown.
#include <cstdio>
#include <cuda_runtime.h>
__device__ __forceinline__ void mma884(double& d0, double& d1, double a, double b) {
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n" : "+d"(d0), "+d"(d1) : "d"(a), "d"(b));
}
__global__ void k(double* out, long long* clk, int mma_iters, int fma_iters, int mode, int ilp) {
    const int w = threadIdx.x >> 5;
    double s = threadIdx.x * 1e-9;
    long long t0 = clock64();
    if (w < 4) {
        if (mode & 1) {
            double acc[8][2];
            for (int i = 0; i < 8; ++i) acc[i][0] = acc[i][1] = s + i;
            for (int it = 0; it < mma_iters; ++it)
#pragma unroll
                for (int i = 0; i < 8; ++i) mma884(acc[i][0], acc[i][1], 1.0 + s, 0.5);
            for (int i = 0; i < 8; ++i) s += acc[i][0] + acc[i][1];
        }
    } else {
        if (mode & 2) {
            double a0 = s, a1 = s + 1, a2 = s + 2, a3 = s + 3;
            for (int it = 0; it < fma_iters; ++it) {
                a0 = fma(a0, 1.0000001, 0.5);
                if (ilp > 1) a1 = fma(a1, 1.0000001, 0.5);
                if (ilp > 2) { a2 = fma(a2, 1.0000001, 0.5); a3 = fma(a3, 1.0000001, 0.5); }
            }
            s += a0 + a1 + a2 + a3;
        }
    }
    long long t1 = clock64();
    if ((threadIdx.x & 31) == 0 && blockIdx.x == 0) clk[w] = t1 - t0;
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
int main() {
    double* out; long long* clk; cudaMalloc(&out, 148 * 256 * 8); cudaMalloc(&clk, 64);
    long long h[8];
    const int MI = 20000, FI = 20000;
    for (int ilp = 1; ilp <= 4; ilp *= 2)
    for (int mode = 1; mode <= 3; ++mode) {
        k<<<148, 256>>>(out, clk, MI, FI, mode, ilp); cudaDeviceSynchronize();
        k<<<148, 256>>>(out, clk, MI, FI, mode, ilp); cudaDeviceSynchronize();
        cudaMemcpy(h, clk, 64, cudaMemcpyDeviceToHost);
        printf("ilp=%d mode=%d (%s): dmma warp %lld cycles (%.1f / dmma), dfma warp %lld cycles (%.1f / dfma-step)\n", ilp, mode,
               mode == 1 ? "dmma only" : mode == 2 ? "dfma only" : "both", h[0], (double)h[0] / (MI * 8.0), h[4], (double)h[4] / FI);
    }
    return 0;
}
