// Microbenchmark: mma.sync.m16n8k32 u8*u8->s32 issue rate on sm_100a (MACs per clk per SM)
#include <cstdio>
#include <cuda_runtime.h>
#include <cstdint>
#define CK(x) do{cudaError_t e=(x); if(e!=cudaSuccess){printf("ERR %s line %d\n",cudaGetErrorString(e),__LINE__);return 1;}}while(0)
__device__ __forceinline__ void mma_u8(int (&c)[4], const unsigned (&a)[4], const unsigned (&b)[2]) {
    asm volatile("mma.sync.aligned.m16n8k32.row.col.s32.u8.u8.s32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};\n"
                 : "+r"(c[0]), "+r"(c[1]), "+r"(c[2]), "+r"(c[3])
                 : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b[0]), "r"(b[1]));
}
template<int NACC>
__global__ void __launch_bounds__(256) k(int iters, int* out, int seed) {
    int c[NACC][4]; unsigned a[4], b[2];
    for (int i=0;i<4;i++) a[i]=seed*(i+3)+threadIdx.x;
    for (int i=0;i<2;i++) b[i]=seed*(i+7)+threadIdx.x;
    for (int n=0;n<NACC;n++) for (int i=0;i<4;i++) c[n][i]=0;
    for (int it=0; it<iters; ++it) {
#pragma unroll
        for (int n=0;n<NACC;n++) mma_u8(c[n], a, b);
    }
    int s=0; for (int n=0;n<NACC;n++) for (int i=0;i<4;i++) s+=c[n][i];
    if (s==0x12345) out[0]=s;
}
template<int NACC> int run(int warps_per_sm_blocks, int nsm, double hz) {
    int* d; CK(cudaMalloc(&d,4)); int iters=4000;
    cudaEvent_t a,b; cudaEventCreate(&a); cudaEventCreate(&b);
    k<NACC><<<nsm*warps_per_sm_blocks,256>>>(10,d,1); CK(cudaDeviceSynchronize());
    float best=1e9;
    for(int r=0;r<3;r++){ cudaEventRecord(a); k<NACC><<<nsm*warps_per_sm_blocks,256>>>(iters,d,1); cudaEventRecord(b); CK(cudaEventSynchronize(b)); float ms; cudaEventElapsedTime(&ms,a,b); if(ms<best)best=ms; }
    double mmas=(double)nsm*warps_per_sm_blocks*8.0*iters*NACC;   // warp-level mma count
    double macs=mmas*16*8*32;
    printf("NACC=%2d blocks/SM=%d: %.3f ms  %.1f mma/clk/SM  %.0f MAC/clk/SM  (%.1f TMAC/s)\n", NACC, warps_per_sm_blocks, best, mmas/(best*1e-3)/nsm/hz, macs/(best*1e-3)/nsm/hz, macs/(best*1e-3)/1e12);
    cudaFree(d); return 0;
}
int main(){ cudaDeviceProp p; CK(cudaGetDeviceProperties(&p,0)); int khz; cudaDeviceGetAttribute(&khz,cudaDevAttrClockRate,0);
    printf("%s SMs=%d clock=%d kHz\n",p.name,p.multiProcessorCount,khz); double hz=khz*1e3;
    for (int b=1;b<=4;b*=2){ run<1>(b,p.multiProcessorCount,hz); run<4>(b,p.multiProcessorCount,hz); run<8>(b,p.multiProcessorCount,hz); run<16>(b,p.multiProcessorCount,hz);} return 0; }
