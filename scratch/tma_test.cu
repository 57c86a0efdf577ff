// Standalone check of the TMA window load used by pm_points_kernel.
#include <cuda.h>
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdint>
#include <cstring>
#include <vector>
#define CK(x) do{cudaError_t e=(x); if(e!=cudaSuccess){printf("ERR %s at line %d\n",cudaGetErrorString(e),__LINE__);return 1;}}while(0)
__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }

struct Args { int x0, y0, box_w, box_h; unsigned char* out; int mode; };

__global__ void __launch_bounds__(192) k(const Args a, const __grid_constant__ CUtensorMap tmap) {
    extern __shared__ __align__(128) unsigned char sm[];
    __shared__ __align__(8) unsigned long long bar;
    const int tid = threadIdx.x;
    if (tid == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(&bar)), "r"(1));
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    if (tid == 0) {
        asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(&bar)), "r"(a.box_w * a.box_h) : "memory");
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
        asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];"
                     ::"r"(smem_u32(sm)), "l"(&tmap), "r"(a.x0), "r"(a.y0), "r"(smem_u32(&bar)) : "memory");
    }
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "WAIT_%=:\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
        "@p bra DONE_%=;\n\t"
        "bra WAIT_%=;\n\t"
        "DONE_%=:\n\t}" ::"r"(smem_u32(&bar)), "r"(0) : "memory");
    for (int i = tid; i < a.box_w * a.box_h; i += blockDim.x) a.out[i] = sm[i];
}
typedef CUresult (*EncodeTiledFn)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *, const cuuint64_t *, const cuuint32_t *, const cuuint32_t *, CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
int main() {
    const int rows = 1000, cols = 1200; const long long pitch = 1232;
    std::vector<unsigned char> h((size_t)pitch * rows);
    for (int y = 0; y < rows; ++y) for (int x = 0; x < pitch; ++x) h[(size_t)y * pitch + x] = (unsigned char)((x * 7 + y * 13) & 255);
    unsigned char *d, *dout; CK(cudaMalloc(&d, h.size() + 4096)); CK(cudaMemcpy(d, h.data(), h.size(), cudaMemcpyHostToDevice));
    const int box_w = 96, box_h = 75; CK(cudaMalloc(&dout, box_w * box_h));
    void *fn = nullptr; cudaDriverEntryPointQueryResult q;
    CK(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &q));
    printf("entry point %p q=%d\n", fn, (int)q);
    alignas(64) CUtensorMap map; memset(&map, 0, sizeof map);
    const cuuint64_t gdim[2] = {(cuuint64_t)cols, (cuuint64_t)rows}; const cuuint64_t gstride[1] = {(cuuint64_t)pitch};
    const cuuint32_t box[2] = {(cuuint32_t)box_w, (cuuint32_t)box_h}; const cuuint32_t estr[2] = {1, 1};
    CUresult r = ((EncodeTiledFn)fn)(&map, CU_TENSOR_MAP_DATA_TYPE_UINT8, 2, d, gdim, gstride, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    printf("encode result %d\n", (int)r);
    int tests[9][2] = {{0, 0}, {16, 11}, {32, 7}, {48, 900}, {1152, 960}, {8, 3}, {4, 3}, {2, 3}, {37, 11}};
    for (auto &t : tests) {
        Args a{t[0], t[1], box_w, box_h, dout, 0};
        CK(cudaMemset(dout, 0xEE, box_w * box_h));
        void *params[] = {&a, &map};
        CK(cudaLaunchKernel((const void *)k, dim3(1), dim3(192), params, box_w * box_h + 256, 0));
        cudaError_t e = cudaDeviceSynchronize();
        printf("x0=%d y0=%d: sync -> %s\n", t[0], t[1], cudaGetErrorString(e));
        if (e != cudaSuccess) { printf("   (stopping: context is dead)\n"); return 2; }
        std::vector<unsigned char> o(box_w * box_h); CK(cudaMemcpy(o.data(), dout, o.size(), cudaMemcpyDeviceToHost));
        int bad = 0;
        for (int y = 0; y < box_h; ++y) for (int x = 0; x < box_w; ++x) {
            int gx = t[0] + x, gy = t[1] + y; unsigned char want = (gx < cols && gy < rows) ? h[(size_t)gy * pitch + gx] : 0;
            if (o[y * box_w + x] != want) ++bad;
        }
        printf("   mismatches %d of %d\n", bad, box_w * box_h);
    }
    return 0;
}
