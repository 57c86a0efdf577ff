// Standalone checks + rates of tcgen05.mma kind::i8 (u8 x u8 -> s32) in the shapes the pattern-matching
// kernel needs.  Everything is compared with a CPU computation; the program prints PASS/FAIL per case.
//
//   check modes (one CTA):
//     0  SS, both operands K-major SWIZZLE_128B (rows 128 B apart), B start row shifted by `rs` rows
//        (start address + rs*128, base_offset field = 0 or (rs & 7))
//     1  SS, both operands K-major no-swizzle "panel" layout (16-byte column panels, rows 16 B apart:
//        SBO = 128, LBO = rows*16), B shifted by rs rows (start address + rs*16)
//     2  TS, A written to TMEM with tcgen05.st.32x32b (lane = row, 4 K bytes per 32-bit column), B panel
//     3  TS, B SWIZZLE_128B
//   corr mode: the full correlation of one grid point (window 75x75, 3 templates 35x35):
//        D[(x,a)][y] = sum_i sum_k Toep_a,i[x][k] * W[y+i][k], Toeplitz rows generated into TMEM from 4 byte-shifted
//        copies of the template rows, window rows selected by the descriptor start address
//   rate mode: cycles per MMA for N = 48..256, SS and TS, one CTA per SM on every SM
#include <cuda.h>
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdint>
#include <cstring>
#include <cstdlib>
#include <vector>
#define CK(x) do{cudaError_t e_=(x); if(e_!=cudaSuccess){printf("ERR %s at line %d\n",cudaGetErrorString(e_),__LINE__);exit(1);}}while(0)

__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(unsigned long long *bar, unsigned count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_wait(unsigned long long *bar, unsigned parity) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "WAIT_%=:\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
        "@p bra DONE_%=;\n\t"
        "bra WAIT_%=;\n\t"
        "DONE_%=:\n\t}" ::"r"(smem_u32(bar)), "r"(parity) : "memory");
}
__device__ __forceinline__ void tmem_alloc(uint32_t *dst_smem, uint32_t ncols) {   // one full warp
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(dst_smem)), "r"(ncols) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc(uint32_t taddr, uint32_t ncols) {     // same warp
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_commit(unsigned long long *bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mma_i8_ss(uint32_t d, uint64_t da, uint64_t db, uint32_t idesc, uint32_t acc) {
    asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
                 "tcgen05.mma.cta_group::1.kind::i8 [%0], %1, %2, %3, p;\n\t}"
                 ::"r"(d), "l"(da), "l"(db), "r"(idesc), "r"(acc) : "memory");
}
__device__ __forceinline__ void mma_i8_ts(uint32_t d, uint32_t ta, uint64_t db, uint32_t idesc, uint32_t acc) {
    asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
                 "tcgen05.mma.cta_group::1.kind::i8 [%0], [%1], %2, %3, p;\n\t}"
                 ::"r"(d), "r"(ta), "l"(db), "r"(idesc), "r"(acc) : "memory");
}
__device__ __forceinline__ void tmem_ld16(uint32_t taddr, uint32_t (&v)[16]) {
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
                 : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]),
                   "=r"(v[8]), "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15])
                 : "r"(taddr));
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_ld16_nowait(uint32_t taddr, uint32_t (&v)[16]) {
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
                 : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]),
                   "=r"(v[8]), "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15])
                 : "r"(taddr));
}
__device__ __forceinline__ void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }
__device__ __forceinline__ void tmem_st8(uint32_t taddr, const uint32_t (&v)[8]) {
    asm volatile("tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8};"
                 ::"r"(taddr), "r"(v[0]), "r"(v[1]), "r"(v[2]), "r"(v[3]), "r"(v[4]), "r"(v[5]), "r"(v[6]), "r"(v[7]) : "memory");
}
__device__ __forceinline__ void tmem_st16(uint32_t taddr, const uint32_t (&v)[16]) {
    asm volatile("tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16};"
                 ::"r"(taddr), "r"(v[0]), "r"(v[1]), "r"(v[2]), "r"(v[3]), "r"(v[4]), "r"(v[5]), "r"(v[6]), "r"(v[7]),
                   "r"(v[8]), "r"(v[9]), "r"(v[10]), "r"(v[11]), "r"(v[12]), "r"(v[13]), "r"(v[14]), "r"(v[15]) : "memory");
}
__device__ __forceinline__ void tmem_st_wait() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }

// shared-memory matrix descriptor (sm_100: version 1)
__host__ __device__ inline uint64_t make_desc(uint32_t saddr, uint32_t lbo, uint32_t sbo, uint32_t layout, uint32_t base_off) {
    uint64_t d = 0;
    d |= (uint64_t)((saddr >> 4) & 0x3fffu);
    d |= (uint64_t)((lbo >> 4) & 0x3fffu) << 16;
    d |= (uint64_t)((sbo >> 4) & 0x3fffu) << 32;
    d |= (uint64_t)1 << 46;
    d |= (uint64_t)(base_off & 7u) << 49;
    d |= (uint64_t)(layout & 7u) << 61;
    return d;
}
__host__ __device__ inline uint32_t make_idesc_u8(int M, int N) {
    return (2u << 4) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);   // c = s32, a = b = u8, K-major both
}

constexpr int KB = 96;          // K bytes per row used everywhere (3 MMAs of K = 32)
constexpr int BROWS = 96;       // rows of the B operand resident (N + max shift)

struct CheckArgs {
    const uint8_t *A;   // [128][KB]
    const uint8_t *B;   // [BROWS][KB]
    int *D;             // [128][N]
    int mode, N, rs, bo_mode;
};

__global__ void __launch_bounds__(128) k_check(CheckArgs a) {
    extern __shared__ __align__(1024) uint8_t sm[];
    __shared__ __align__(8) unsigned long long bar;
    __shared__ uint32_t tbase_s;
    uint8_t *sA = sm;                    // 128 rows x 128 B (either layout fits in 16 KB)
    uint8_t *sB = sm + 16384;            // BROWS rows x 128 B
    const int tid = threadIdx.x, warp = tid >> 5;
    if (tid == 0 && (smem_u32(sm) & 1023u)) printf("dynamic smem base %u not 1024-aligned\n", smem_u32(sm));
    const bool panelA = a.mode == 1, panelB = a.mode == 1 || a.mode == 2;
    for (int e = tid; e < 128 * 128; e += 128) sA[e] = 0;
    for (int e = tid; e < BROWS * 128; e += 128) sB[e] = 0;
    __syncthreads();
    for (int e = tid; e < 128 * KB; e += 128) {
        const int r = e / KB, k = e % KB;
        const int off = panelA ? (k >> 4) * (128 * 16) + r * 16 + (k & 15) : r * 128 + ((((k >> 4) ^ (r & 7))) << 4) + (k & 15);
        sA[off] = a.A[e];
    }
    for (int e = tid; e < BROWS * KB; e += 128) {
        const int r = e / KB, k = e % KB;
        const int off = panelB ? (k >> 4) * (BROWS * 16) + r * 16 + (k & 15) : r * 128 + ((((k >> 4) ^ (r & 7))) << 4) + (k & 15);
        sB[off] = a.B[e];
    }
    if (tid == 0) { mbar_init(&bar, 1); asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
    if (warp == 0) tmem_alloc(&tbase_s, 128);
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");      // generic-proxy smem writes -> visible to the tensor core
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tb = tbase_s;
    const uint32_t tD = tb, tA = tb + 64;
    if (a.mode >= 2) {          // A -> TMEM: thread = row
        const uint32_t lane_base = (uint32_t)(warp * 32) << 16;
        for (int ks = 0; ks < 3; ++ks) {
            uint32_t v[8];
            for (int c = 0; c < 8; ++c) {
                const uint8_t *p = a.A + tid * KB + ks * 32 + c * 4;
                v[c] = p[0] | (p[1] << 8) | (p[2] << 16) | ((uint32_t)p[3] << 24);
            }
            tmem_st8(tA + lane_base + ks * 8, v);
        }
        tmem_st_wait();
        tc_fence_before();
        __syncthreads();
        tc_fence_after();
    }
    if (tid == 0) {
        const uint32_t idesc = make_idesc_u8(128, a.N);
        for (int ks = 0; ks < 3; ++ks) {
            uint64_t da, db;
            if (panelA) da = make_desc(smem_u32(sA) + ks * 2 * (128 * 16), 128 * 16, 128, 0, 0);
            else da = make_desc(smem_u32(sA) + ks * 32, 16, 1024, 2, 0);
            if (panelB) db = make_desc(smem_u32(sB) + a.rs * 16 + ks * 2 * (BROWS * 16), BROWS * 16, 128, 0, 0);
            else db = make_desc(smem_u32(sB) + a.rs * 128 + ks * 32, 16, 1024, 2, a.bo_mode ? (a.rs & 7) : 0);
            if (a.mode >= 2) mma_i8_ts(tD, tA + ks * 8, db, idesc, ks > 0);
            else mma_i8_ss(tD, da, db, idesc, ks > 0);
        }
        tc_commit(&bar);
    }
    mbar_wait(&bar, 0);
    tc_fence_after();
    for (int c0 = 0; c0 < a.N; c0 += 16) {
        uint32_t v[16];
        tmem_ld16(tD + ((uint32_t)(warp * 32) << 16) + c0, v);
        for (int c = 0; c < 16; ++c) a.D[tid * a.N + c0 + c] = (int)v[c];
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 0) tmem_dealloc(tb, 128);
}

// ---------------------------------------------------------------- rate
struct RateArgs { long long *cycles; int N, ts, nmma, iters, nacc; };
__global__ void __launch_bounds__(128) k_rate(RateArgs a) {
    extern __shared__ __align__(1024) uint8_t sm[];
    __shared__ __align__(8) unsigned long long bar;
    __shared__ uint32_t tbase_s;
    const int tid = threadIdx.x, warp = tid >> 5;
    for (int e = tid; e < 16384 + 320 * 128; e += 128) sm[e] = (uint8_t)(e * 7 + 1);
    if (tid == 0) { mbar_init(&bar, 1); asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
    if (warp == 0) tmem_alloc(&tbase_s, 512);
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tb = tbase_s;
    {   // deterministic TMEM A contents
        uint32_t v[16];
        for (int c = 0; c < 16; ++c) v[c] = 0x01020304u * (c + 1);
        tmem_st16(tb + 480 + ((uint32_t)(warp * 32) << 16), v);
        tmem_st16(tb + 496 + ((uint32_t)(warp * 32) << 16), v);
        tmem_st_wait();
        tc_fence_before();
        __syncthreads();
        tc_fence_after();
    }
    long long t0 = 0, t1 = 0;
    if (tid == 0) {
        const uint32_t idesc = make_idesc_u8(128, a.N);
        unsigned ph = 0;
        t0 = clock64();
        for (int it = 0; it < a.iters; ++it) {
            const uint64_t da0 = make_desc(smem_u32(sm), 16, 1024, 2, 0);
            const uint64_t db0 = make_desc(smem_u32(sm + 16384), 16, 1024, 2, 0);
            int acc_i = 0;
            for (int rs = 0; rs < 35; ++rs) {
#pragma unroll
                for (int ks = 0; ks < 3; ++ks) {
                    const uint32_t td = tb + (uint32_t)acc_i * (uint32_t)a.N;          // nacc independent accumulator chains
                    if (a.ts) mma_i8_ts(td, tb + 480 + ks * 8, db0 + (uint64_t)(rs * 8 + ks * 2), idesc, (rs * 3 + ks) >= a.nacc);
                    else mma_i8_ss(td, da0 + (uint64_t)(ks * 2), db0 + (uint64_t)(rs * 8 + ks * 2), idesc, (rs * 3 + ks) >= a.nacc);
                    acc_i = acc_i + 1 == a.nacc ? 0 : acc_i + 1;
                }
            }
            tc_commit(&bar);
            mbar_wait(&bar, ph); ph ^= 1u;
        }
        t1 = clock64();
        a.cycles[blockIdx.x] = t1 - t0;
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 0) tmem_dealloc(tb, 512);
}

// ---------------------------------------------------------------- full correlation of one point
constexpr int CS = 35, CR = 41, CWIN = 75, CNA = 3;         // template, result, window sizes, angles
constexpr int CWROWS = 96;                                  // window rows resident per panel (>= CS - 1 + 48)
constexpr int CTPITCH = 96;                                 // bytes per byte-shifted template-row copy
constexpr int CTL = 24;                                     // byte offset of template column 0 in copy 0
constexpr int CTANG = CS * 4 * CTPITCH + 16;                // bytes per angle (+16: angles land on different banks)
struct CorrArgs {
    const uint8_t *win;     // [CWIN][CWIN]
    const uint8_t *tpl;     // [CNA][CS][CS]
    int *D;                 // [128][48]  lane m = 3*x + a, column y
    long long *cycles;
    int off;                // byte offset of window column 0 inside the staged rows (0..15)
    int reps;
};
__global__ void __launch_bounds__(128) k_corr(CorrArgs a) {
    extern __shared__ __align__(1024) uint8_t sm[];
    __shared__ __align__(8) unsigned long long bar[3];
    __shared__ uint32_t tbase_s;
    uint8_t *sW = sm;                                    // 6 panels x CWROWS rows x 16 B
    uint8_t *sT = sm + 6 * CWROWS * 16;                  // [CNA] x CTANG: [CS][4 copies][CTPITCH]
    const int tid = threadIdx.x, warp = tid >> 5;
    for (int e = tid; e < 6 * CWROWS * 16 + CNA * CTANG; e += 128) sm[e] = 0;
    __syncthreads();
    for (int e = tid; e < CWIN * CWIN; e += 128) {
        const int r = e / CWIN, k = e % CWIN + a.off;
        sW[(k >> 4) * (CWROWS * 16) + r * 16 + (k & 15)] = a.win[e];
    }
    for (int e = tid; e < CNA * CS * CS; e += 128) {
        const int j = e % CS, ai = e / CS, ang = ai / CS, i = ai % CS;
        for (int s = 0; s < 4; ++s) sT[ang * CTANG + (i * 4 + s) * CTPITCH + CTL + s + j] = a.tpl[e];
    }
    if (tid == 0) { mbar_init(&bar[0], 1); mbar_init(&bar[1], 1); mbar_init(&bar[2], 1); asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
    if (warp == 0) tmem_alloc(&tbase_s, 256);
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tb = tbase_s;
    const uint32_t lane_base = (uint32_t)(warp * 32) << 16;
    const uint32_t tD = tb, tA = tb + 144;                // D: one accumulator per K step (3 x 48 columns); A slots: 3 x 24 columns
    // lane m = 3*x + a (x-major so that a warp's rows share a narrow band of non-zero K words)
    const int m = tid, x = m / 3, ang = m - 3 * x;
    const bool live = x < CR;
    const int q = x + a.off;
    const int xmin = (warp * 32) / 3;
    const int cw0 = min((xmin + a.off) >> 2, 8);          // first of the 16 K words this warp writes (warp-uniform)
    const uint32_t *trow = reinterpret_cast<const uint32_t *>(sT + (size_t)ang * CTANG + (q & 3) * CTPITCH) + (CTL >> 2) - (q >> 2) + cw0;
    const uint32_t idesc = make_idesc_u8(128, 48);
    unsigned ph[3] = {0, 0, 0};
    long long t0 = clock64();
    for (int rep = 0; rep < a.reps; ++rep) {
        {   // clear both A slots (columns outside this point's band must be zero)
            uint32_t z[8];
            for (int c = 0; c < 8; ++c) z[c] = 0;
            for (int c = 0; c < 72; c += 8) tmem_st8(tA + lane_base + c, z);
            tmem_st_wait();
        }
        for (int i = 0; i < CS; ++i) {
            const int slot = i % 3;
            if (i >= 3) {
                mbar_wait(&bar[slot], ph[slot]); ph[slot] ^= 1u;
                tc_fence_after();
            }
            uint32_t v[16];
            const uint32_t *p = trow + i * (4 * CTPITCH / 4);
#pragma unroll
            for (int c = 0; c < 16; ++c) v[c] = live ? p[c] : 0u;
            tmem_st16(tA + lane_base + slot * 24 + cw0, v);
            tmem_st_wait();
            tc_fence_before();
            __syncthreads();
            if (tid == 0) {
                tc_fence_after();
#pragma unroll
                for (int ks = 0; ks < 3; ++ks) {
                    const uint64_t db = make_desc(smem_u32(sW) + i * 16 + ks * 2 * (CWROWS * 16), CWROWS * 16, 128, 0, 0);
                    mma_i8_ts(tD + ks * 48, tA + slot * 24 + ks * 8, db, idesc, i > 0 ? 1u : 0u);
                }
                tc_commit(&bar[slot]);
            }
        }
        // rows 32, 33, 34 are the last users of slots 2, 0, 1: every slot has one completion outstanding
        for (int sl = 0; sl < 3; ++sl) { mbar_wait(&bar[sl], ph[sl]); ph[sl] ^= 1u; }
        tc_fence_after();
    }
    long long t1 = clock64();
    for (int c0 = 0; c0 < 48; c0 += 16) {
        uint32_t v[16], w[16], u[16];
        tmem_ld16(tD + lane_base + c0, v);
        tmem_ld16(tD + lane_base + 48 + c0, w);
        tmem_ld16(tD + lane_base + 96 + c0, u);
        for (int c = 0; c < 16; ++c) a.D[tid * 48 + c0 + c] = (int)(v[c] + w[c] + u[c]);
    }
    if (tid == 0) a.cycles[blockIdx.x] = t1 - t0;
    tc_fence_before();
    __syncthreads();
    if (warp == 0) tmem_dealloc(tb, 256);
}


// ---------------------------------------------------------------- warp-specialised MAC phase prototype
// Warp 0 lane 0 issues the MMAs; gen warp groups (4 warps each = the four TMEM lane quarters) write the Toeplitz rows
// of RPS template rows per step into one of NSLOT A slots; full[slot] (gen -> issuer, 128 arrivals) and empty[slot]
// (tcgen05.commit -> gen) mbarriers; NACC accumulators (K step ks goes to accumulator ks % NACC) summed in the epilogue.
__device__ __forceinline__ void mbar_arrive(unsigned long long *bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
template <int RPS, int NWG, int NACC, int NSLOT>
__global__ void __launch_bounds__(128 * (1 + NWG)) k_mac(CorrArgs a) {
    extern __shared__ __align__(1024) uint8_t sm[];
    __shared__ __align__(8) unsigned long long full[NSLOT], empty[NSLOT], done;
    __shared__ uint32_t tbase_s;
    constexpr int NT = 128 * (1 + NWG);
    constexpr int SLOTC = 32 * RPS;
    constexpr int NSTEP = (CS + RPS - 1) / RPS;
    static_assert(NACC * 48 + NSLOT * SLOTC <= 256, "TMEM budget");
    uint8_t *sW = sm;
    uint8_t *sT = sm + 6 * CWROWS * 16;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    for (int e = tid; e < 6 * CWROWS * 16 + CNA * CTANG + 4 * CTPITCH; e += NT) sm[e] = 0;
    __syncthreads();
    for (int e = tid; e < CWIN * CWIN; e += NT) {
        const int r = e / CWIN, k = e % CWIN + a.off;
        sW[(k >> 4) * (CWROWS * 16) + r * 16 + (k & 15)] = a.win[e];
    }
    for (int e = tid; e < CNA * CS * CS; e += NT) {
        const int j = e % CS, ai = e / CS, ang = ai / CS, i = ai % CS;
        for (int s = 0; s < 4; ++s) sT[ang * CTANG + (i * 4 + s) * CTPITCH + CTL + s + j] = a.tpl[e];
    }
    if (tid == 0) {
        for (int k = 0; k < NSLOT; ++k) { mbar_init(&full[k], 128); mbar_init(&empty[k], 1); }
        mbar_init(&done, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 0) tmem_alloc(&tbase_s, 256);
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tb = tbase_s;
    const uint32_t tD = tb, tA = tb + NACC * 48;
    const uint32_t lane_base = (uint32_t)((warp & 3) * 32) << 16;
    const int wg = warp >> 2;                 // 0: issuer group, 1..NWG: gen groups
    const int m = tid & 127, x = m / 3, ang = m - 3 * x;
    const bool live = x < CR;
    const int q = x + a.off;
    const int xmin = ((warp & 3) * 32) / 3;
    const int cw0 = (xmin + a.off) >> 2;
    const uint32_t *trow = reinterpret_cast<const uint32_t *>(sT + (size_t)ang * CTANG + (q & 3) * CTPITCH) + (CTL >> 2) - (q >> 2) + cw0;
    const uint32_t idesc = make_idesc_u8(128, 48);
    unsigned nfull[NSLOT], nempty[NSLOT];     // completions consumed so far (same sequence in every thread)
    for (int k = 0; k < NSLOT; ++k) { nfull[k] = 0; nempty[k] = 0; }
    long long t0 = clock64();
    for (int rep = 0; rep < a.reps; ++rep) {
        if (wg >= 1) {   // clear the A slots (columns outside this point's band must be zero); every gen group clears all
            uint32_t z[8];
            for (int c = 0; c < 8; ++c) z[c] = 0;
            for (int c = 0; c < NSLOT * SLOTC; c += 8) tmem_st8(tA + lane_base + c, z);
            tmem_st_wait();
            tc_fence_before();
        }
        __syncthreads();
        tc_fence_after();
        if (wg == 0) {
            if (warp == 0 && lane == 0) {
                uint64_t dbase[3];
                for (int ks = 0; ks < 3; ++ks) dbase[ks] = make_desc(smem_u32(sW) + ks * 2 * (CWROWS * 16), CWROWS * 16, 128, 0, 0);
                for (int st = 0; st < NSTEP; ++st) {
                    const int slot = st % NSLOT;
                    mbar_wait(&full[slot], (nfull[slot] + (unsigned)(st / NSLOT)) & 1u);
                    tc_fence_after();
#pragma unroll
                    for (int rr = 0; rr < RPS; ++rr) {
                        const int i = st * RPS + rr;
                        if (i < CS) {
#pragma unroll
                            for (int ks = 0; ks < 3; ++ks)
                                mma_i8_ts(tD + (ks % NACC) * 48, tA + slot * SLOTC + rr * 32 + ks * 8, dbase[ks] + (uint64_t)i, idesc,
                                          (i * 3 + ks) >= NACC ? 1u : 0u);
                        }
                    }
                    tc_commit(&empty[slot]);
                }
                tc_commit(&done);            // all MMAs of this point
            }
            // every thread of the issuer group keeps the same counters
            for (int st = 0; st < NSTEP; ++st) nfull[st % NSLOT]++;
        } else {
            for (int st = wg - 1; st < NSTEP; st += NWG) {
                const int slot = st % NSLOT;
                // completions of empty[slot] needed before this write = number of earlier steps of this point in that slot
                const unsigned need = nempty[slot] + (unsigned)(st / NSLOT);
                if (st >= NSLOT) { mbar_wait(&empty[slot], (need - 1u) & 1u); tc_fence_after(); }
#pragma unroll
                for (int rr = 0; rr < RPS; ++rr) {
                    const int i = st * RPS + rr;
                    uint32_t v[16];
                    const uint32_t *p = trow + (i < CS ? i : CS) * (4 * CTPITCH / 4);      // row CS = zeros (odd tail)
#pragma unroll
                    for (int c = 0; c < 16; ++c) v[c] = live ? p[c] : 0u;
                    tmem_st16(tA + lane_base + slot * SLOTC + rr * 32 + cw0, v);
                }
                tmem_st_wait();
                tc_fence_before();
                mbar_arrive(&full[slot]);
            }
        }
        for (int k = 0; k < NSLOT; ++k) {       // all steps of this point, per slot
            unsigned cnt = 0;
            for (int st = k; st < NSTEP; st += NSLOT) ++cnt;
            nempty[k] += cnt;
            if (wg != 0) nfull[k] += cnt;
        }
        // drain: one completion of `done` per point, observed by every thread (a parity wait can only tell adjacent
        // phases apart, so threads that skipped intermediate completions of empty[] must not wait on those)
        mbar_wait(&done, (unsigned)rep & 1u);
        tc_fence_after();
    }
    long long t1 = clock64();
    if (wg == 1) {
        for (int c0 = 0; c0 < 48; c0 += 16) {
            uint32_t acc[16];
            for (int c = 0; c < 16; ++c) acc[c] = 0;
            for (int k = 0; k < NACC; ++k) {
                uint32_t v[16];
                tmem_ld16(tD + lane_base + k * 48 + c0, v);
                for (int c = 0; c < 16; ++c) acc[c] += v[c];
            }
            for (int c = 0; c < 16; ++c) a.D[m * 48 + c0 + c] = (int)acc[c];
        }
    }
    if (tid == 0) a.cycles[blockIdx.x] = t1 - t0;
    tc_fence_before();
    __syncthreads();
    if (warp == 0) tmem_dealloc(tb, 256);
}

// TMEM load throughput: every warp reads `cols` columns of its lane quarter, `iters` times
__global__ void __launch_bounds__(128) k_ldtm(long long *cycles, int *sink, int cols, int iters) {
    __shared__ uint32_t tbase_s;
    const int warp = threadIdx.x >> 5;
    if (warp == 0) tmem_alloc(&tbase_s, 256);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tb = tbase_s + ((uint32_t)(warp * 32) << 16);
    uint32_t acc = 0;
    long long t0 = clock64();
    for (int it = 0; it < iters; ++it)
        for (int c = 0; c < cols; c += 48) {
            uint32_t v[16], w[16], u[16];
            tmem_ld16_nowait(tb + c, v);
            tmem_ld16_nowait(tb + c + 16, w);
            tmem_ld16_nowait(tb + c + 32, u);
            tmem_ld_wait();
            for (int k = 0; k < 16; ++k) acc += v[k] + w[k] + u[k];
        }
    long long t1 = clock64();
    if (threadIdx.x == 0) cycles[blockIdx.x] = t1 - t0;
    if (acc == 0x12345u) sink[0] = 1;
    tc_fence_before();
    __syncthreads();
    if (warp == 0) tmem_dealloc(tbase_s, 256);
}

template <int RPS, int NWG, int NACC, int NSLOT>
static void run_mac(const char *name, const uint8_t *dw, const uint8_t *dt, int *dD, long long *dcyc, int nsm,
                    const std::vector<uint8_t> &hw, const std::vector<uint8_t> &ht) {
    const int smem = 6 * CWROWS * 16 + CNA * CTANG + 4 * CTPITCH;
    auto kfn = k_mac<RPS, NWG, NACC, NSLOT>;
    CK(cudaFuncSetAttribute(kfn, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    const int nt = 128 * (1 + NWG);
    CK(cudaMemset(dD, 0xff, 128 * 48 * 4));
    CorrArgs ca{dw, dt, dD, dcyc, 7, 1};
    kfn<<<1, nt, smem>>>(ca);
    cudaError_t e = cudaDeviceSynchronize();
    if (e != cudaSuccess) { printf("mac %s: CUDA error %s (stopping)\n", name, cudaGetErrorString(e)); exit(2); }
    std::vector<int> hD(128 * 48);
    CK(cudaMemcpy(hD.data(), dD, hD.size() * 4, cudaMemcpyDeviceToHost));
    int bad = 0;
    for (int ang = 0; ang < CNA; ++ang) for (int y = 0; y < CR; ++y) for (int x = 0; x < CR; ++x) {
        int ref = 0;
        for (int i = 0; i < CS; ++i) for (int j = 0; j < CS; ++j) ref += (int)hw[(y + i) * CWIN + x + j] * (int)ht[(ang * CS + i) * CS + j];
        if (ref != hD[(3 * x + ang) * 48 + y]) ++bad;
    }
    printf("mac %s: %s (%d mismatches)", name, bad ? "FAIL" : "PASS", bad);
    for (int cps = 1; cps <= 2; ++cps) {
        const int reps = 200;
        CorrArgs cb{dw, dt, dD, dcyc, 7, reps};
        cudaEvent_t e0, e1; CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
        kfn<<<nsm * cps, nt, smem>>>(cb);
        CK(cudaDeviceSynchronize());
        CK(cudaEventRecord(e0));
        kfn<<<nsm * cps, nt, smem>>>(cb);
        CK(cudaEventRecord(e1));
        CK(cudaDeviceSynchronize());
        float ms; CK(cudaEventElapsedTime(&ms, e0, e1));
        printf(" | %d CTA/SM: %.0f clk/point/SM, %.2f M points/s", cps, ms * 1e-3 * 1.965e9 / reps / cps, nsm * cps * reps / (ms * 1e-3) / 1e6);
    }
    printf("\n");
}


// several issuing threads (one per warp) each with its own accumulator; M selectable
struct Rate2Args { long long *cycles; int N, M, nissue, iters; };
__global__ void __launch_bounds__(128) k_rate2(Rate2Args a) {
    extern __shared__ __align__(1024) uint8_t sm[];
    __shared__ __align__(8) unsigned long long bar[4];
    __shared__ uint32_t tbase_s;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    for (int e = tid; e < 16384 + 320 * 128; e += 128) sm[e] = (uint8_t)(e * 7 + 1);
    if (tid == 0) { for (int k = 0; k < 4; ++k) mbar_init(&bar[k], 1); asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
    if (warp == 0) tmem_alloc(&tbase_s, 512);
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tb = tbase_s;
    {
        uint32_t v[16];
        for (int c = 0; c < 16; ++c) v[c] = 0x01020304u * (c + 1);
        tmem_st16(tb + 480 + ((uint32_t)(warp * 32) << 16), v);
        tmem_st16(tb + 496 + ((uint32_t)(warp * 32) << 16), v);
        tmem_st_wait();
        tc_fence_before();
        __syncthreads();
        tc_fence_after();
    }
    long long t0 = clock64();
    if (lane == 0 && warp < a.nissue) {
        const uint32_t idesc = make_idesc_u8(a.M, a.N);
        unsigned ph = 0;
        const uint64_t db0 = make_desc(smem_u32(sm + 16384), 16, 1024, 2, 0);
        const uint32_t td = tb + (uint32_t)warp * (uint32_t)a.N;
        for (int it = 0; it < a.iters; ++it) {
            for (int rs = 0; rs < 35; ++rs) {
#pragma unroll
                for (int ks = 0; ks < 3; ++ks)
                    mma_i8_ts(td, tb + 480 + ks * 8, db0 + (uint64_t)(rs * 8 + ks * 2), idesc, (rs * 3 + ks) >= 1);
            }
            tc_commit(&bar[warp]);
            mbar_wait(&bar[warp], ph); ph ^= 1u;
        }
    }
    __syncthreads();
    long long t1 = clock64();
    if (tid == 0) a.cycles[blockIdx.x] = t1 - t0;
    tc_fence_before();
    __syncthreads();
    if (warp == 0) tmem_dealloc(tb, 512);
}

// cycle breakdown of one row iteration of the simple (4 gen warps, thread 0 issues) pipeline, thread 0's view
struct TimingArgs { const uint8_t *win; const uint8_t *tpl; long long *acc; int off; int reps; };
__global__ void __launch_bounds__(128) k_rowtime(TimingArgs a) {
    extern __shared__ __align__(1024) uint8_t sm[];
    __shared__ __align__(8) unsigned long long bar[3];
    __shared__ uint32_t tbase_s;
    uint8_t *sW = sm;
    uint8_t *sT = sm + 6 * CWROWS * 16;
    const int tid = threadIdx.x, warp = tid >> 5;
    for (int e = tid; e < 6 * CWROWS * 16 + CNA * CTANG; e += 128) sm[e] = (uint8_t)(e * 13 + 1);
    if (tid == 0) { mbar_init(&bar[0], 1); mbar_init(&bar[1], 1); mbar_init(&bar[2], 1); asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
    if (warp == 0) tmem_alloc(&tbase_s, 256);
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tb = tbase_s;
    const uint32_t lane_base = (uint32_t)(warp * 32) << 16;
    const uint32_t tD = tb, tA = tb + 144;
    const int m = tid, x = m / 3, ang = m - 3 * x;
    const int q = x + a.off;
    const int cw0 = min((((warp * 32) / 3) + a.off) >> 2, 8);
    const uint32_t *trow = reinterpret_cast<const uint32_t *>(sT + (size_t)ang * CTANG + (q & 3) * CTPITCH) + (CTL >> 2) - (q >> 2) + cw0;
    const uint32_t idesc = make_idesc_u8(128, 48);
    unsigned ph[3] = {0, 0, 0};
    long long t_wait = 0, t_lds = 0, t_st = 0, t_stwait = 0, t_sync = 0, t_issue = 0;
    for (int rep = 0; rep < a.reps; ++rep) {
        for (int i = 0; i < CS; ++i) {
            const int slot = i % 3;
            long long c0 = clock64();
            if (i >= 3) { mbar_wait(&bar[slot], ph[slot]); ph[slot] ^= 1u; tc_fence_after(); }
            long long c1 = clock64();
            uint32_t v[16];
            const uint32_t *p = trow + i * (4 * CTPITCH / 4);
#pragma unroll
            for (int c = 0; c < 16; ++c) v[c] = p[c];
            uint32_t sink = 0;
#pragma unroll
            for (int c = 0; c < 16; ++c) sink ^= v[c];
            if (sink == 0x12345u) a.acc[100] = 1;             // force the loads to complete here
            long long c2 = clock64();
            tmem_st16(tA + lane_base + slot * 24 + cw0, v);
            long long c3 = clock64();
            tmem_st_wait();
            long long c4 = clock64();
            tc_fence_before();
            __syncthreads();
            long long c5 = clock64();
            if (tid == 0) {
                tc_fence_after();
#pragma unroll
                for (int ks = 0; ks < 3; ++ks) {
                    const uint64_t db = make_desc(smem_u32(sW) + i * 16 + ks * 2 * (CWROWS * 16), CWROWS * 16, 128, 0, 0);
                    mma_i8_ts(tD + ks * 48, tA + slot * 24 + ks * 8, db, idesc, i > 0 ? 1u : 0u);
                }
                tc_commit(&bar[slot]);
            }
            long long c6 = clock64();
            t_wait += c1 - c0; t_lds += c2 - c1; t_st += c3 - c2; t_stwait += c4 - c3; t_sync += c5 - c4; t_issue += c6 - c5;
        }
        for (int sl = 0; sl < 3; ++sl) { mbar_wait(&bar[sl], ph[sl]); ph[sl] ^= 1u; }
        tc_fence_after();
    }
    if (tid == 0 && blockIdx.x == 0) {
        a.acc[0] = t_wait; a.acc[1] = t_lds; a.acc[2] = t_st; a.acc[3] = t_stwait; a.acc[4] = t_sync; a.acc[5] = t_issue;
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 0) tmem_dealloc(tb, 256);
}
static uint32_t rng_state = 12345u;
static uint32_t rnd() { rng_state = rng_state * 1664525u + 1013904223u; return rng_state >> 8; }

int main(int argc, char **argv) {
    int dev = 0; cudaDeviceProp prop; CK(cudaGetDeviceProperties(&prop, dev));
    printf("device %s, %d SMs, clock %d kHz\n", prop.name, prop.multiProcessorCount, prop.clockRate);
    const int nsm = prop.multiProcessorCount;

    if (argc > 1 && !strcmp(argv[1], "rowtime")) {
        long long *dacc; CK(cudaMalloc(&dacc, 1024 * 8)); CK(cudaMemset(dacc, 0, 1024 * 8));
        const int smem = 6 * CWROWS * 16 + CNA * CTANG + 4 * CTPITCH;
        CK(cudaFuncSetAttribute(k_rowtime, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
        for (int cps = 1; cps <= 2; ++cps) {
            TimingArgs ta{nullptr, nullptr, dacc, 5, 100};
            k_rowtime<<<nsm * cps, 128, smem>>>(ta);
            CK(cudaDeviceSynchronize());
            long long h[6]; CK(cudaMemcpy(h, dacc, 48, cudaMemcpyDeviceToHost));
            const double rows = 100.0 * CS;
            printf("rowtime %d CTA/SM (clk per row, thread 0): slot wait %.0f | 16 LDS %.0f | st issue %.0f | wait::st %.0f | fence+syncthreads %.0f | 3 MMA + commit %.0f | total %.0f\n",
                   cps, h[0] / rows, h[1] / rows, h[2] / rows, h[3] / rows, h[4] / rows, h[5] / rows, (h[0] + h[1] + h[2] + h[3] + h[4] + h[5]) / rows);
        }
        return 0;
    }
    // ---------------- checks
    std::vector<uint8_t> hA(128 * KB), hB(BROWS * KB);
    for (auto &v : hA) v = (uint8_t)(rnd() & 255);
    for (auto &v : hB) v = (uint8_t)(rnd() & 255);
    uint8_t *dA, *dB; int *dD;
    CK(cudaMalloc(&dA, hA.size())); CK(cudaMalloc(&dB, hB.size())); CK(cudaMalloc(&dD, 128 * 256 * 4));
    CK(cudaMemcpy(dA, hA.data(), hA.size(), cudaMemcpyHostToDevice));
    CK(cudaMemcpy(dB, hB.data(), hB.size(), cudaMemcpyHostToDevice));
    CK(cudaFuncSetAttribute(k_check, cudaFuncAttributeMaxDynamicSharedMemorySize, 65536));
    struct Case { int mode, N, rs, bo; };
    const Case cases[] = {{0, 48, 0, 0}, {0, 64, 0, 0}, {0, 48, 8, 0}, {0, 48, 1, 0}, {0, 48, 1, 1}, {0, 48, 5, 0}, {0, 48, 5, 1}, {0, 48, 34, 0}, {0, 48, 34, 1},
                          {1, 48, 0, 0}, {1, 48, 1, 0}, {1, 48, 5, 0}, {1, 48, 34, 0},
                          {2, 48, 0, 0}, {2, 48, 7, 0}, {2, 48, 34, 0},
                          {3, 48, 0, 0}, {3, 48, 7, 0}, {3, 48, 7, 1}, {3, 48, 34, 1}, {3, 48, 34, 0}};
    for (const Case &c : cases) {
        CK(cudaMemset(dD, 0xff, 128 * 256 * 4));
        CheckArgs ca{dA, dB, dD, c.mode, c.N, c.rs, c.bo};
        k_check<<<1, 128, 65536>>>(ca);
        cudaError_t e = cudaDeviceSynchronize();
        if (e != cudaSuccess) { printf("check mode %d N %d rs %d bo %d: CUDA error %s (stopping)\n", c.mode, c.N, c.rs, c.bo, cudaGetErrorString(e)); return 2; }
        std::vector<int> hD(128 * c.N);
        CK(cudaMemcpy(hD.data(), dD, hD.size() * 4, cudaMemcpyDeviceToHost));
        int bad = 0, first = -1;
        for (int m = 0; m < 128; ++m) for (int n = 0; n < c.N; ++n) {
            int ref = 0;
            for (int k = 0; k < KB; ++k) ref += (int)hA[m * KB + k] * (int)hB[(n + c.rs) * KB + k];
            if (ref != hD[m * c.N + n]) { if (first < 0) first = m * c.N + n; ++bad; }
        }
        printf("check mode %d N %3d rowshift %2d base_offset_mode %d: %s (%d mismatches", c.mode, c.N, c.rs, c.bo, bad ? "FAIL" : "PASS", bad);
        if (bad) printf(", first at m=%d n=%d got %d", first / c.N, first % c.N, hD[first]);
        printf(")\n");
    }
    // ---------------- full correlation
    {
        std::vector<uint8_t> hw(CWIN * CWIN), ht(CNA * CS * CS);
        for (auto &v : hw) v = (uint8_t)(1 + rnd() % 255);
        for (auto &v : ht) v = (uint8_t)(1 + rnd() % 255);
        uint8_t *dw, *dt; long long *dcyc;
        CK(cudaMalloc(&dw, hw.size())); CK(cudaMalloc(&dt, ht.size())); CK(cudaMalloc(&dcyc, 8 * 1024));
        CK(cudaMemcpy(dw, hw.data(), hw.size(), cudaMemcpyHostToDevice));
        CK(cudaMemcpy(dt, ht.data(), ht.size(), cudaMemcpyHostToDevice));
        const int smem = 6 * CWROWS * 16 + CNA * CTANG;
        CK(cudaFuncSetAttribute(k_corr, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
        for (int off : {0, 5, 15}) {
            CK(cudaMemset(dD, 0xff, 128 * 48 * 4));
            CorrArgs ca{dw, dt, dD, dcyc, off, 1};
            k_corr<<<1, 128, smem>>>(ca);
            cudaError_t e = cudaDeviceSynchronize();
            if (e != cudaSuccess) { printf("corr off %d: CUDA error %s (stopping)\n", off, cudaGetErrorString(e)); return 2; }
            std::vector<int> hD(128 * 48);
            CK(cudaMemcpy(hD.data(), dD, hD.size() * 4, cudaMemcpyDeviceToHost));
            int bad = 0;
            for (int ang = 0; ang < CNA; ++ang) for (int y = 0; y < CR; ++y) for (int x = 0; x < CR; ++x) {
                int ref = 0;
                for (int i = 0; i < CS; ++i) for (int j = 0; j < CS; ++j) ref += (int)hw[(y + i) * CWIN + x + j] * (int)ht[(ang * CS + i) * CS + j];
                if (ref != hD[(3 * x + ang) * 48 + y]) ++bad;
            }
            printf("corr off %2d: %s (%d of %d mismatches)\n", off, bad ? "FAIL" : "PASS", bad, CNA * CR * CR);
        }
        // steady-state cycles per point-batch with 1..3 CTAs per SM
        for (int cps = 1; cps <= 2; ++cps) {
            const int reps = 200;
            CorrArgs ca{dw, dt, dD, dcyc, 5, reps};
            cudaEvent_t e0, e1; CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
            k_corr<<<nsm * cps, 128, smem>>>(ca);      // warm
            CK(cudaDeviceSynchronize());
            CK(cudaEventRecord(e0));
            k_corr<<<nsm * cps, 128, smem>>>(ca);
            CK(cudaEventRecord(e1));
            CK(cudaDeviceSynchronize());
            float ms; CK(cudaEventElapsedTime(&ms, e0, e1));
            std::vector<long long> hc(nsm * cps);
            CK(cudaMemcpy(hc.data(), dcyc, hc.size() * 8, cudaMemcpyDeviceToHost));
            double avg = 0; for (auto v : hc) avg += (double)v; avg /= hc.size();
            printf("corr rate: %d CTA/SM: %.0f clk per point (3 angles) per CTA, %.1f us per point per SM, %.3f ms total -> %.2f M points/s/GPU, %.1f useful TMAC/s\n",
                   cps, avg / reps, ms * 1e3 / reps / cps, ms, nsm * cps * reps / (ms * 1e-3) / 1e6,
                   nsm * cps * reps / (ms * 1e-3) * 3.0 * CS * CS * CR * CR / 1e12);
        }

        run_mac<1, 1, 1, 4>("RPS1 NWG1 NACC1 NSLOT4", dw, dt, dD, dcyc, nsm, hw, ht);
        run_mac<1, 1, 3, 3>("RPS1 NWG1 NACC3 NSLOT3", dw, dt, dD, dcyc, nsm, hw, ht);
        run_mac<1, 2, 1, 4>("RPS1 NWG2 NACC1 NSLOT4", dw, dt, dD, dcyc, nsm, hw, ht);
        run_mac<1, 2, 3, 3>("RPS1 NWG2 NACC3 NSLOT3", dw, dt, dD, dcyc, nsm, hw, ht);
        run_mac<1, 3, 1, 6>("RPS1 NWG3 NACC1 NSLOT6", dw, dt, dD, dcyc, nsm, hw, ht);
        run_mac<2, 1, 1, 3>("RPS2 NWG1 NACC1 NSLOT3", dw, dt, dD, dcyc, nsm, hw, ht);
        run_mac<2, 1, 2, 2>("RPS2 NWG1 NACC2 NSLOT2", dw, dt, dD, dcyc, nsm, hw, ht);
        run_mac<2, 2, 1, 3>("RPS2 NWG2 NACC1 NSLOT3", dw, dt, dD, dcyc, nsm, hw, ht);
        run_mac<2, 2, 2, 2>("RPS2 NWG2 NACC2 NSLOT2", dw, dt, dD, dcyc, nsm, hw, ht);
        {
            int *dsink; CK(cudaMalloc(&dsink, 4));
            for (int cols : {48, 144, 240}) {
                k_ldtm<<<nsm, 128>>>(dcyc, dsink, cols, 1000);
                CK(cudaDeviceSynchronize());
                long long c0; CK(cudaMemcpy(&c0, dcyc, 8, cudaMemcpyDeviceToHost));
                printf("ldtm: 4 warps x %d columns: %.1f clk per pass, %.1f B/clk/SM\n", cols, c0 / 1000.0, 128.0 * cols * 4 / (c0 / 1000.0));
            }
        }
    }

    {
        long long *dcyc; CK(cudaMalloc(&dcyc, nsm * 8));
        CK(cudaFuncSetAttribute(k_rate2, cudaFuncAttributeMaxDynamicSharedMemorySize, 16384 + 320 * 128));
        struct R2 { int N, M, ni; };
        const R2 r2s[] = {{48, 128, 1}, {48, 128, 2}, {48, 128, 4}, {48, 64, 1}, {48, 64, 2}, {48, 64, 4}, {96, 128, 2}, {96, 128, 4}, {16, 128, 4}};
        for (const R2 &r : r2s) {
            Rate2Args ra{dcyc, r.N, r.M, r.ni, 200};
            k_rate2<<<nsm, 128, 16384 + 320 * 128>>>(ra);
            cudaError_t e = cudaDeviceSynchronize();
            if (e != cudaSuccess) { printf("rate2 N %d M %d: CUDA error %s (stopping)\n", r.N, r.M, cudaGetErrorString(e)); return 2; }
            std::vector<long long> hc(nsm);
            CK(cudaMemcpy(hc.data(), dcyc, nsm * 8, cudaMemcpyDeviceToHost));
            double avg = 0; for (auto v : hc) avg += (double)v; avg /= nsm;
            const double per = avg / (105.0 * 200.0 * r.ni);
            printf("rate2 TS M=%3d N=%3d, %d issuing warps: %.1f clk per MMA per SM, %.0f MAC/clk/SM\n", r.M, r.N, r.ni, per, (double)r.M * r.N * 32 / per);
        }
    }
    // ---------------- MMA rates
    {
        long long *dcyc; CK(cudaMalloc(&dcyc, nsm * 8));
        CK(cudaFuncSetAttribute(k_rate, cudaFuncAttributeMaxDynamicSharedMemorySize, 16384 + 320 * 128));
        struct RC { int N, nacc; };
        const RC rcs[] = {{48, 1}, {48, 2}, {48, 3}, {48, 4}, {48, 6}, {48, 8}, {16, 8}, {96, 4}, {128, 3}, {208, 2}, {256, 1}};
        for (int ts = 0; ts < 2; ++ts) for (const RC &rc : rcs) {
            const int N = rc.N;
            RateArgs ra{dcyc, N, ts, 105, 200, rc.nacc};
            k_rate<<<nsm, 128, 16384 + 320 * 128>>>(ra);
            CK(cudaDeviceSynchronize());
            cudaEvent_t e0, e1; CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
            CK(cudaEventRecord(e0));
            k_rate<<<nsm, 128, 16384 + 320 * 128>>>(ra);
            CK(cudaEventRecord(e1));
            cudaError_t e = cudaDeviceSynchronize();
            if (e != cudaSuccess) { printf("rate N %d: CUDA error %s (stopping)\n", N, cudaGetErrorString(e)); return 2; }
            float ms; CK(cudaEventElapsedTime(&ms, e0, e1));
            std::vector<long long> hc(nsm);
            CK(cudaMemcpy(hc.data(), dcyc, nsm * 8, cudaMemcpyDeviceToHost));
            double avg = 0; for (auto v : hc) avg += (double)v; avg /= nsm;
            const double per = avg / (105.0 * 200.0);
            printf("rate %s M=128 N=%3d K=32 u8, %d accumulator chains: %.1f clk per MMA (batches of 105 + commit/wait), %.0f MAC/clk/SM, chip %.1f TMAC/s by events\n",
                   ts ? "TS" : "SS", N, rc.nacc, per, 128.0 * N * 32 / per, (double)nsm * 105 * 200 * 128.0 * N * 32 / (ms * 1e-3) / 1e12);
        }
    }
    return 0;
}
