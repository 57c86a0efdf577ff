// sid_knn_kernel.cuh -- brute-force Hamming 2-nearest-neighbour matcher for 256-bit (ORB) descriptors.
//
// Replaces cv2.BFMatcher(cv2.NORM_HAMMING).knnMatch(d1, d2, k=2) as called by the reference's feature
// tracking (reference sea_ice_drift/ftlib.py:92-99, the `matcher` plug-in kwarg of get_match_coords).
// Exact integers; ordering pinned against OpenCV 4.13: smaller distance first, equal distances by lower
// train index.  A strict '<' insertion while candidates arrive in increasing train index reproduces it.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace sid {

constexpr int KNN_THREADS = 128;      // queries per CTA (one per thread)
constexpr int KNN_TILE = 128;         // train descriptors staged per step (4 KB)

struct Top2 { int d0, i0, d1, i1; };

__device__ __forceinline__ void top2_insert(Top2 &t, int d, int i) {
    if (d < t.d0) { t.d1 = t.d0; t.i1 = t.i0; t.d0 = d; t.i0 = i; }
    else if (d < t.d1) { t.d1 = d; t.i1 = i; }
}

// grid = (ceil(n1 / KNN_THREADS), nseg): thread = one query, blockIdx.y = one contiguous train segment.
__global__ void __launch_bounds__(KNN_THREADS) knn_hamming2_kernel(const uint4 *__restrict__ q, int n1,
                                                                   const uint4 *__restrict__ t, int n2,
                                                                   int seg_len, Top2 *__restrict__ part) {
    __shared__ uint4 tile[KNN_TILE * 2];
    const int qi = blockIdx.x * KNN_THREADS + threadIdx.x;
    const int t0 = blockIdx.y * seg_len, t1 = min(n2, t0 + seg_len);
    uint4 qa = make_uint4(0, 0, 0, 0), qb = qa;
    if (qi < n1) { qa = __ldg(q + 2 * (size_t)qi); qb = __ldg(q + 2 * (size_t)qi + 1); }
    Top2 best;
    best.d0 = best.d1 = 0x7fffffff; best.i0 = best.i1 = -1;
    for (int base = t0; base < t1; base += KNN_TILE) {
        const int cnt = min(KNN_TILE, t1 - base);
        __syncthreads();
        for (int k = threadIdx.x; k < 2 * cnt; k += KNN_THREADS) tile[k] = __ldg(t + 2 * (size_t)base + k);
        __syncthreads();
#pragma unroll 4
        for (int j = 0; j < cnt; ++j) {
            const uint4 a = tile[2 * j], b = tile[2 * j + 1];      // same address for every lane: broadcast
            const int d = __popc(qa.x ^ a.x) + __popc(qa.y ^ a.y) + __popc(qa.z ^ a.z) + __popc(qa.w ^ a.w) +
                          __popc(qb.x ^ b.x) + __popc(qb.y ^ b.y) + __popc(qb.z ^ b.z) + __popc(qb.w ^ b.w);
            top2_insert(best, d, base + j);
        }
    }
    if (qi < n1) part[(size_t)blockIdx.y * n1 + qi] = best;
}

// Merge the per-segment candidates in segment (= train index) order; out: idx[n1][2], dist[n1][2] (-1 = none).
__global__ void knn_merge_kernel(const Top2 *__restrict__ part, int n1, int nseg, int *__restrict__ idx, int *__restrict__ dist) {
    const int qi = blockIdx.x * blockDim.x + threadIdx.x;
    if (qi >= n1) return;
    Top2 best;
    best.d0 = best.d1 = 0x7fffffff; best.i0 = best.i1 = -1;
    for (int sgm = 0; sgm < nseg; ++sgm) {
        const Top2 c = part[(size_t)sgm * n1 + qi];
        if (c.i0 >= 0) top2_insert(best, c.d0, c.i0);
        if (c.i1 >= 0) top2_insert(best, c.d1, c.i1);
    }
    idx[2 * qi] = best.i0; idx[2 * qi + 1] = best.i1;
    dist[2 * qi] = best.i0 >= 0 ? best.d0 : -1;
    dist[2 * qi + 1] = best.i1 >= 0 ? best.d1 : -1;
}

}  // namespace sid
