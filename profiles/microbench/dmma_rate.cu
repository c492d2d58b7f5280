#include <cstdio>
#include <cuda_runtime.h>
__device__ __forceinline__ void mma884(double& d0, double& d1, double a, double b) {
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n" : "+d"(d0), "+d"(d1) : "d"(a), "d"(b));
}
__device__ __forceinline__ void mma16816(double* d, const double* a, const double* b) {
    asm volatile("mma.sync.aligned.m16n8k16.row.col.f64.f64.f64.f64 {%0,%1,%2,%3}, {%4,%5,%6,%7,%8,%9,%10,%11}, {%12,%13,%14,%15}, {%0,%1,%2,%3};\n"
        : "+d"(d[0]), "+d"(d[1]), "+d"(d[2]), "+d"(d[3])
        : "d"(a[0]), "d"(a[1]), "d"(a[2]), "d"(a[3]), "d"(a[4]), "d"(a[5]), "d"(a[6]), "d"(a[7]), "d"(b[0]), "d"(b[1]), "d"(b[2]), "d"(b[3]));
}
__device__ __forceinline__ void mma1688(double* d, const double* a, const double* b) {
    asm volatile("mma.sync.aligned.m16n8k8.row.col.f64.f64.f64.f64 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};\n"
        : "+d"(d[0]), "+d"(d[1]), "+d"(d[2]), "+d"(d[3])
        : "d"(a[0]), "d"(a[1]), "d"(a[2]), "d"(a[3]), "d"(b[0]), "d"(b[1]));
}
__device__ __forceinline__ void mma1684(double* d, const double* a, const double* b) {
    asm volatile("mma.sync.aligned.m16n8k4.row.col.f64.f64.f64.f64 {%0,%1,%2,%3}, {%4,%5}, {%6}, {%0,%1,%2,%3};\n"
        : "+d"(d[0]), "+d"(d[1]), "+d"(d[2]), "+d"(d[3])
        : "d"(a[0]), "d"(a[1]), "d"(b[0]));
}
template<int MODE, int NACC>
__global__ void k(double* out, int iters, double seed) {
    double acc[NACC][4];
    for (int i=0;i<NACC;i++) for (int j=0;j<4;j++) acc[i][j]=threadIdx.x*1e-9+i;
    double a[8], b[4];
    for (int i=0;i<8;i++) a[i]=seed+i*1e-3; for (int i=0;i<4;i++) b[i]=seed*0.5+i*1e-3;
    for (int it=0; it<iters; ++it) {
#pragma unroll
        for (int i=0;i<NACC;i++) {
            if (MODE==0) { acc[i][0]=fma(a[0],b[0],acc[i][0]); acc[i][1]=fma(a[1],b[1],acc[i][1]); acc[i][2]=fma(a[2],b[2],acc[i][2]); acc[i][3]=fma(a[3],b[3],acc[i][3]); }
            if (MODE==1) { mma884(acc[i][0],acc[i][1],a[0],b[0]); mma884(acc[i][2],acc[i][3],a[1],b[1]); }
            if (MODE==2) mma1684(acc[i],a,b);
            if (MODE==3) mma1688(acc[i],a,b);
            if (MODE==4) mma16816(acc[i],a,b);
        }
    }
    double s=0; for (int i=0;i<NACC;i++) for (int j=0;j<4;j++) s+=acc[i][j];
    out[blockIdx.x*blockDim.x+threadIdx.x]=s;
}
template<int MODE,int NACC> void run(const char* name, double flop_per_iter_per_warp, int warps_per_block, int blocks_per_sm) {
    double* out; cudaMalloc(&out, 148*8*1024*8);
    int iters=20000; int threads=warps_per_block*32; int blocks=148*blocks_per_sm;
    k<MODE,NACC><<<blocks,threads>>>(out,100,1.0); cudaDeviceSynchronize();
    cudaEvent_t e0,e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    cudaEventRecord(e0); k<MODE,NACC><<<blocks,threads>>>(out,iters,1.0); cudaEventRecord(e1); cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms,e0,e1);
    double flops=(double)iters*NACC*flop_per_iter_per_warp*warps_per_block*blocks;
    printf("%-12s nacc=%d warps/blk=%d blk/sm=%d : %.2f TFLOP/s (%.3f ms) err=%s\n", name, NACC, warps_per_block, blocks_per_sm, flops/ms/1e9, ms, cudaGetErrorString(cudaGetLastError()));
    cudaFree(out);
}
int main(){
    // flops per inner op per warp: DFMA: 4 fma *32 lanes*2 ; m8n8k4 x2: 2*512 ; m16n8k4: 2*16*8*4 ; m16n8k8: 2*16*8*8; m16n8k16: 2*16*8*16
    run<0,8>("dfma",4*32*2,8,1); run<0,8>("dfma",4*32*2,16,2); run<0,4>("dfma",4*32*2,8,4);
    run<1,8>("m8n8k4",2*512,4,1); run<1,8>("m8n8k4",2*512,8,1); run<1,8>("m8n8k4",2*512,16,2); run<1,2>("m8n8k4",2*512,16,2);
    run<2,8>("m16n8k4",1024,8,1); run<2,8>("m16n8k4",1024,16,2);
    run<3,8>("m16n8k8",2048,8,1); run<3,8>("m16n8k8",2048,16,2);
    run<4,8>("m16n8k16",4096,4,1); run<4,8>("m16n8k16",4096,8,1); run<4,8>("m16n8k16",4096,16,2); run<4,2>("m16n8k16",4096,16,2);
    return 0;
}
