// sid_api.cu -- C ABI (include/sid_b200.h) over the sm_100a kernels.
// No exceptions cross the boundary; every entry point returns a SID_* code.
#include <cuda.h>
#include <cuda_runtime.h>
#include <algorithm>
#include <cmath>
#include <cstdio>
#include <cstring>
#include <cstdlib>
#include <new>
#include <string>
#include <condition_variable>
#include <functional>
#include <mutex>
#include <thread>
#include <vector>
#include <emmintrin.h>

#include "../../include/sid_b200.h"
#include "sid_common.cuh"
#include "sid_pm_kernel.cuh"
#include "sid_pm_tc_kernel.cuh"
#include "sid_pm_ws_kernel.cuh"
#include "sid_single_kernels.cuh"
#include "sid_knn_kernel.cuh"
#include "sid_defor_kernel.cuh"
#include "sid_fg_kernel.cuh"

using namespace sid;

namespace {
#ifndef SID_WS_DEFAULT
#define SID_WS_DEFAULT 1
#endif

constexpr size_t IMG_TAIL_SLACK = 4096;   // bytes readable past the last image row

struct DevBuf {
    void *p = nullptr;
    size_t cap = 0;
    bool owned = true;      // false: caller-owned memory adopted by sid_adopt_pair_device (never freed / grown here)
};

}  // namespace

// Persistent host threads for the staged upload of pageable images (creating 7 threads per band and image cost ~25 % of a
// band's copy time).  Owned by a context; the workers sleep on a condition variable between jobs.
class HostPool {
public:
    explicit HostPool(int workers) : nworkers_(workers) {
        for (int i = 0; i < workers; ++i) threads_.emplace_back([this, i] { loop(i); });
    }
    ~HostPool() {
        { std::lock_guard<std::mutex> lk(m_); stop_ = true; }
        cv_.notify_all();
        for (std::thread &t : threads_) t.join();
    }
    int size() const { return nworkers_ + 1; }                 // the caller works too
    // fn(part) for part = 0 .. size() - 1, in parallel; returns when all parts are done
    void run(const std::function<void(int)> &fn) {
        { std::lock_guard<std::mutex> lk(m_); job_ = &fn; pending_ = nworkers_; ++gen_; }
        cv_.notify_all();
        fn(nworkers_);
        std::unique_lock<std::mutex> lk(m_);
        done_.wait(lk, [this] { return pending_ == 0; });
        job_ = nullptr;
    }
private:
    void loop(int id) {
        unsigned seen = 0;
        for (;;) {
            const std::function<void(int)> *job;
            {
                std::unique_lock<std::mutex> lk(m_);
                cv_.wait(lk, [&] { return stop_ || gen_ != seen; });
                if (stop_) return;
                seen = gen_;
                job = job_;
            }
            (*job)(id);
            { std::lock_guard<std::mutex> lk(m_); if (--pending_ == 0) done_.notify_one(); }
        }
    }
    int nworkers_;
    std::vector<std::thread> threads_;
    std::mutex m_;
    std::condition_variable cv_, done_;
    const std::function<void(int)> *job_ = nullptr;
    unsigned gen_ = 0;
    int pending_ = 0;
    bool stop_ = false;
};

struct sid_ctx {
    int device = 0;
    int sm_count = 0;
    int max_smem_optin = 0;
    cudaStream_t own_stream = nullptr;
    cudaStream_t stream = nullptr;
    cudaStream_t copy_stream = nullptr;      // image upload, overlapped with compute (sid_run_pair)
    cudaStream_t band_stream[2] = {nullptr, nullptr};   // band launches alternate here so their tails overlap
    cudaEvent_t slot_event[3] = {};          // points uploaded, band stream 0 done, band stream 1 done
    cudaEvent_t band_event[17] = {};
    std::string err;
    long long launches = 0;
    // resident image pair (padded pitch, 16-byte multiple)
    DevBuf img1, img2;
    int rows1 = 0, cols1 = 0, rows2 = 0, cols2 = 0;
    long long pitch1 = 0, pitch2 = 0;
    bool have_pair = false;
    // per-call buffers
    DevBuf pts, order, out, status, angles, scratch, counter, misc, tail_maps, tail_recs, epi, fg;
    void *pin = nullptr;
    size_t pin_cap = 0;
    cudaEvent_t k_ev[2] = {};                // bracket the last fused-kernel launch (sid_last_kernel_ms)
    bool k_ev_valid = false;
    const char *k_name = "";
    bool k_ev_hold = false;                  // second launch of a two-class call: keep the start event of the first
    HostPool *pool = nullptr;                // staged upload of pageable images
    size_t tail_region_stride = 0;           // floats per point of the tail-map regions of the current host call (0: the launch's own)
    long long table_n = -1;                  // rows of the result table the last sid_run / sid_run_pair left in `out`
    long long tail_hint_n = 0;               // total points of the current host call (sizes the tail hand-off once)
    // staged upload of pageable host images (sid_run_pair): pinned double buffer + "slot free again" events
    void *stage = nullptr;
    size_t stage_cap = 0;
    cudaEvent_t stage_event[2] = {};
    void *encode_tiled = nullptr;            // cuTensorMapEncodeTiled, resolved through the runtime (no libcuda link)
    bool encode_tried = false;
};

namespace {

int fail(sid_ctx *c, int code, const std::string &msg) {
    if (c) c->err = msg;
    return code;
}
int cuda_fail(sid_ctx *c, cudaError_t e, const char *what) {
    char buf[512];
    snprintf(buf, sizeof buf, "%s: %s", what, cudaGetErrorString(e));
    cudaGetLastError();
    return fail(c, SID_ECUDA, buf);
}
#define CU(call)                                                   \
    do {                                                           \
        cudaError_t e_ = (call);                                   \
        if (e_ != cudaSuccess) return cuda_fail(ctx, e_, #call);   \
    } while (0)

// cudaFuncAttributeMaxDynamicSharedMemorySize is a per-process, per-device property of the kernel (not of a
// context): raise it to the device's opt-in maximum, and re-assert it whenever some other context may have
// changed it -- setting it is cheap and idempotent.
int allow_max_smem(sid_ctx *ctx, const void *kernel) {
    cudaFuncAttributes fa;
    CU(cudaFuncGetAttributes(&fa, kernel));
    CU(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                            ctx->max_smem_optin - (int)fa.sharedSizeBytes));
    return SID_OK;
}

int reserve(sid_ctx *ctx, DevBuf &b, size_t bytes) {
    if (bytes <= b.cap && b.owned) return SID_OK;
    if (b.p && b.owned) cudaFree(b.p);
    b.p = nullptr; b.cap = 0; b.owned = true;
    size_t want = bytes + bytes / 4 + 256;
    cudaError_t e = cudaMalloc(&b.p, want);
    if (e != cudaSuccess) { cudaGetLastError(); b.p = nullptr; return fail(ctx, SID_ENOMEM, "device allocation failed"); }
    b.cap = want;
    return SID_OK;
}
int reserve_pinned(sid_ctx *ctx, size_t bytes) {
    if (bytes <= ctx->pin_cap) return SID_OK;
    if (ctx->pin) { cudaFreeHost(ctx->pin); ctx->pin = nullptr; ctx->pin_cap = 0; }
    size_t want = bytes + bytes / 4 + 256;
    cudaError_t e = cudaMallocHost(&ctx->pin, want);
    if (e != cudaSuccess) { cudaGetLastError(); return fail(ctx, SID_ENOMEM, "pinned allocation failed"); }
    ctx->pin_cap = want;
    return SID_OK;
}

long long padded_pitch(int cols) { return ((long long)cols + 15) / 16 * 16 + 16; }

int upload_image(sid_ctx *ctx, DevBuf &buf, long long &pitch, const uint8_t *src, int rows, int cols,
                 long long src_pitch, cudaMemcpyKind kind) {
    if (!src || rows <= 0 || cols <= 0 || src_pitch < cols) return fail(ctx, SID_EINVAL, "bad image arguments");
    pitch = padded_pitch(cols);
    const size_t bytes = (size_t)pitch * rows + IMG_TAIL_SLACK;
    const bool fresh = bytes > buf.cap || !buf.owned;
    int rc = reserve(ctx, buf, bytes);
    if (rc) return rc;
    if (fresh) CU(cudaMemsetAsync(buf.p, 0, buf.cap, ctx->stream));
    CU(cudaMemcpy2DAsync(buf.p, (size_t)pitch, src, (size_t)src_pitch, (size_t)cols, (size_t)rows, kind, ctx->stream));
    return SID_OK;
}

void gaussian_weights(double gw[5]) {
    double w[9], sum = 0.0;
    for (int k = -4; k <= 4; ++k) { w[k + 4] = std::exp(-0.5 * (double)(k * k)); sum += w[k + 4]; }
    for (int k = 0; k < 9; ++k) w[k] /= sum;
    for (int k = 0; k <= 4; ++k) gw[k] = w[4 + k];
}

int check_common(sid_ctx *ctx, int img_size, int n_angles, const double *angle_tab, int rot_order, int mtype) {
    if (!ctx) return SID_EINVAL;
    if (img_size < 2 || img_size > 128) return fail(ctx, SID_EUNSUPPORTED, "img_size must be in [2, 128]");
    if (n_angles <= 0 || !angle_tab) return fail(ctx, SID_EINVAL, "need at least one angle and its table");
    if (rot_order != 0 && rot_order != 1)
        return fail(ctx, SID_EUNSUPPORTED, "rot_order must be 0 or 1 (higher spline orders need a full-image prefilter)");
    if (mtype != SID_TM_CCOEFF_NORMED) return fail(ctx, SID_EUNSUPPORTED, "only TM_CCOEFF_NORMED (5) is implemented");
    return SID_OK;
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *,
                                  const cuuint64_t *, const cuuint32_t *, const cuuint32_t *, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

// uint8 2-D tensor map over image 2 with a box of box_w x box_h bytes (TMA window staging).
bool make_window_tensor_map(sid_ctx *ctx, CUtensorMap *map, int box_w, int box_h, int which = 2) {
    if (!ctx->encode_tried) {
        ctx->encode_tried = true;
        cudaDriverEntryPointQueryResult q;
        void *fn = nullptr;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &q) == cudaSuccess &&
            q == cudaDriverEntryPointSuccess)
            ctx->encode_tiled = fn;
        else
            cudaGetLastError();
    }
    const long long pitch = which == 1 ? ctx->pitch1 : ctx->pitch2;
    if (!ctx->encode_tiled || box_w > 256 || box_h > 256 || (box_w & 15) || (pitch & 15)) return false;
    const cuuint64_t gdim[2] = {(cuuint64_t)(which == 1 ? ctx->cols1 : ctx->cols2), (cuuint64_t)(which == 1 ? ctx->rows1 : ctx->rows2)};
    const cuuint64_t gstride[1] = {(cuuint64_t)pitch};
    const cuuint32_t box[2] = {(cuuint32_t)box_w, (cuuint32_t)box_h};
    const cuuint32_t estr[2] = {1, 1};
    const CUresult r = ((EncodeTiledFn)ctx->encode_tiled)(map, CU_TENSOR_MAP_DATA_TYPE_UINT8, 2, which == 1 ? ctx->img1.p : ctx->img2.p, gdim, gstride, box,
                                                         estr, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE,
                                                         CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    return r == CUDA_SUCCESS;
}

// shared launch path of sid_run / sid_run_device
// Largest shared-memory footprint of pm_tail_kernel for which the tail is split off the correlation kernel
// (SID_PM_TAIL_SMEM_KB overrides; 52 KB = at least 4 tail CTAs per SM)
size_t tail_split_limit() {
    if (const char *e = getenv("SID_PM_TAIL_SMEM_KB")) { const int kb = atoi(e); if (kb > 0) return (size_t)kb * 1024; }
    return (size_t)52 * 1024;
}

int launch_pm(sid_ctx *ctx, long long n, const double *d_c1, const double *d_r1, const double *d_c2fg,
              const double *d_r2fg, const double *d_border, const int *d_order, int max_border, int typ_border,
              int img_size, int n_angles, const double *d_angles, const double *d_tab, int rot_order,
              unsigned flags, double *d_out, int *d_status,
              cudaStream_t st = nullptr, int slot = 0, long long tail_off = 0) {
    if (!st) st = ctx->stream;
    const int s = img_size;
    const int hws = s / 2;
    const int Wmax = 2 * hws + 2 * max_border + 1;
    const int Rmax = Wmax - s + 1;
    if (Rmax < 2) return fail(ctx, SID_EINVAL, "search window not larger than the template");
    PmArgs a;
    memset(&a, 0, sizeof a);
    a.img1 = (const uint8_t *)ctx->img1.p; a.rows1 = ctx->rows1; a.cols1 = ctx->cols1; a.pitch1 = ctx->pitch1;
    a.img2 = (const uint8_t *)ctx->img2.p; a.rows2 = ctx->rows2; a.cols2 = ctx->cols2; a.pitch2 = ctx->pitch2;
    a.n = n;
    a.c1 = d_c1; a.r1 = d_r1; a.c2fg = d_c2fg; a.r2fg = d_r2fg; a.border = d_border; a.order = d_order;
    a.s = s; a.n_angles = n_angles; a.angles = d_angles; a.tab = d_tab;
    a.rot_order = rot_order; a.flags = flags;
    const double area = (double)s * (double)s;
    a.inv_area = 1.0 / area;
    a.sqrt_inv_area = std::sqrt(a.inv_area);
    gaussian_weights(a.gw);
    a.out = d_out; a.status = d_status;
    a.max_rr = Rmax * Rmax;
    a.max_hrw = Wmax * Rmax;
    // correlation path: tcgen05 tensor cores (default), legacy mma.sync (SID_PM_PATH=imma) or the integer pipe
    // (SID_PM_PATH=dp4a); all three give bit-identical results
    const char *path_env = getenv("SID_PM_PATH");
    const bool want_tc = !path_env || strcmp(path_env, "tc") == 0;
    const bool want_ws = path_env ? strcmp(path_env, "ws") == 0 : SID_WS_DEFAULT != 0;
    const bool imma = !(path_env && strcmp(path_env, "dp4a") == 0);
    const bool smth = (flags & SID_HES_SMTH) != 0;
    // Split tail: peak statistics in a second, light kernel when the result map fits its shared memory
    bool split_tail = pm_tail_smem_bytes(a.max_rr, smth) <= tail_split_limit();       // >= 4 tail CTAs per SM; larger maps measured
                                                                             // faster with the fused tail (cfg1: 1.10 vs 1.21 ms)
    if (const char *e = getenv("SID_PM_SPLIT_TAIL")) if (e[0] == '0') split_tail = false;
    alignas(64) CUtensorMap tmap;
    memset(&tmap, 0, sizeof tmap);
    PmTcCfg tg;
    memset(&tg, 0, sizeof tg);
    bool use_tc = false;
    // warp-specialised pipeline kernel (sid_pm_ws_kernel.cuh): small result maps (geometry: radius <= 24; its shared-memory layout fits 227 KB up to radius 20 at s = 35, 22 with two sets of window statistics), needs the split tail
    PmWsCfg wg;
    memset(&wg, 0, sizeof wg);
    alignas(64) CUtensorMap tmap1;                  // image 1: the patch the templates of a point are sampled from
    memset(&tmap1, 0, sizeof tmap1);
    bool use_ws = false;
    if (want_ws && split_tail) {
        // three sets of window statistics where they fit the shared memory, else two (radius 21 ... 22 at img_size 35)
        const int rtyp = 2 * typ_border + (Wmax - 2 * max_border) - s + 1;
        bool fits = false;
        for (int nstat = WS_NSTAT; nstat >= 2 && !fits; --nstat)
            fits = pm_ws_geometry(s, Rmax, Wmax, n_angles, a.max_rr, wg, rtyp, nstat) && (size_t)wg.smem_bytes + 2048 <= (size_t)ctx->max_smem_optin;
        if (fits && make_window_tensor_map(ctx, &tmap, 16, wg.load_rows) && make_window_tensor_map(ctx, &tmap1, wg.pbw, wg.pbh, 1))
            use_ws = true;
    }
    if (use_ws) {
        a.tma = 1; a.ab = wg.nab;
        int rc = reserve(ctx, ctx->counter, 256);
        if (rc) return rc;
        a.counter = (unsigned int *)ctx->counter.p + 16 * slot;
        CU(cudaMemsetAsync(a.counter, 0, sizeof(unsigned int), st));
        a.split_tail = 1;
        const size_t cap_n = (size_t)std::max<long long>(n + tail_off, ctx->tail_hint_n);
        a.tail_stride = (a.max_rr + 3) & ~3;
        // regions of concurrently running launches (bands, border classes) are laid out with ONE stride per host call
        const size_t region_stride = std::max(ctx->tail_region_stride, (size_t)a.tail_stride);
        if ((rc = reserve(ctx, ctx->tail_maps, cap_n * region_stride * sizeof(float)))) return rc;
        if ((rc = reserve(ctx, ctx->tail_recs, cap_n * sizeof(PmTailRec)))) return rc;
        a.tail_maps = (float *)ctx->tail_maps.p + (size_t)tail_off * region_stride;
        a.tail_recs = (PmTailRec *)ctx->tail_recs.p + tail_off;
        const void *kfn = (const void *)pm_ws_kernel;
        if (int arc = allow_max_smem(ctx, kfn)) return arc;
        long long grid = ctx->sm_count;
        if (grid > n) grid = n;
        if (grid < 1) grid = 1;
        if (getenv("SID_DEBUG"))
            fprintf(stderr, "[sid] launch: ws smem=%d nstat=%d nab=%d ks=%d npairs=%d nslots=%d slot_bytes=%d npanels=%d wrows=%d n16max=%d grid=%lld\n",
                    wg.smem_bytes, wg.nstat, wg.nab, wg.ks, wg.npairs, wg.nslots, wg.slot_bytes, wg.npanels, wg.wrows, wg.n16max, grid);
#ifdef SID_WS_PROF
        if ((rc = reserve(ctx, ctx->scratch, (size_t)grid * 6 * 8 * 8))) return rc;
        CU(cudaMemsetAsync(ctx->scratch.p, 0, (size_t)grid * 6 * 8 * 8, st));
        a.scratch = (unsigned char *)ctx->scratch.p;
#endif
        void *params_ws[] = {(void *)&a, (void *)&wg, (void *)&tmap, (void *)&tmap1};
        if (!ctx->k_ev[0]) { CU(cudaEventCreate(&ctx->k_ev[0])); CU(cudaEventCreate(&ctx->k_ev[1])); }
        if (!ctx->k_ev_hold) CU(cudaEventRecord(ctx->k_ev[0], st));
        CU(cudaLaunchKernel(kfn, dim3((unsigned)grid), dim3((unsigned)WS_THREADS), params_ws, (size_t)wg.smem_bytes, st));
        CU(cudaEventRecord(ctx->k_ev[1], st));
        ctx->k_ev_valid = true;
        ctx->k_name = !ctx->k_ev_hold ? "sid::pm_ws_kernel"
                      : strstr(ctx->k_name, "pm_tc_kernel") ? "sid::pm_tc_kernel + sid::pm_ws_kernel"
                      : strstr(ctx->k_name, "pm_points_kernel") ? "sid::pm_points_kernel + sid::pm_ws_kernel" : "sid::pm_ws_kernel";
        ctx->launches += 1;
#ifdef SID_WS_PROF
        {
            std::vector<long long> h((size_t)grid * 48);
            CU(cudaStreamSynchronize(st));
            CU(cudaMemcpy(h.data(), ctx->scratch.p, h.size() * 8, cudaMemcpyDeviceToHost));
            static const char *names[6] = {"control", "mma0", "gather0", "stats0", "epi0", "epi1"};
            const double per_pt = (double)n / (double)grid;
            for (int r = 0; r < 6; ++r) {
                fprintf(stderr, "[ws prof] %-8s clk/point:", names[r]);
                for (int k = 0; k < 8; ++k) {
                    double sum = 0; for (long long c = 0; c < grid; ++c) sum += (double)h[((size_t)c * 6 + r) * 8 + k];
                    fprintf(stderr, " %8.0f", sum / (double)grid / per_pt);
                }
                fprintf(stderr, "\n");
            }
        }
#endif
        const size_t tsm = pm_tail_smem_bytes(a.max_rr, smth);
        if (int arc = allow_max_smem(ctx, (const void *)pm_tail_kernel)) return arc;
        pm_tail_kernel<<<(unsigned)n, pm_tail_threads(a.max_rr), tsm, st>>>(a);
        ctx->launches += 1;
        CU(cudaGetLastError());
        return SID_OK;
    }
    if (want_tc && pm_tc_geometry(s, Rmax, Wmax, n_angles, tg) && make_window_tensor_map(ctx, &tmap, 16, tg.load_rows)) use_tc = true;
    // Measured (profiles/r02_ab_tc_vs_imma.txt): with one CTA per SM (tensor memory > 256 columns: search radius ~100) the
    // latency-bound non-MAC phases make the tcgen05 kernel slower than the mma.sync kernel at three CTAs per SM, so the
    // default takes it only where two CTAs fit; SID_PM_PATH=tc forces it.
    if (use_tc && !path_env && tg.tmem_cols > 256) use_tc = false;
    bool smem_scratch = false;
    size_t smem = 0;
    int variant = 0;
    if (use_tc) {
        // window panels + four byte-shifted copies of the template rows (+ the per-point scratch when two CTAs
        // still fit on an SM)
        a.tma = 1;
        a.ab = tg.nab;
        const size_t base = (size_t)tg.win_bytes + (size_t)tg.tpl_bytes;
        const size_t scratch = pm_scratch_bytes(a.max_rr, a.max_hrw, a.ab, smth, split_tail);
        const size_t wsq_bytes = ((size_t)a.max_rr * 4 + 255) & ~(size_t)255;
        const size_t need = base + scratch + wsq_bytes;
        const size_t cap2 = 108 * 1024;
        if (need <= cap2 && !getenv("SID_PM_GLOBAL_SCRATCH")) { smem_scratch = true; smem = need; }
        else smem = base;
        // window sums on the tensor cores: the two squares windows alias the result maps (dead until the first
        // epilogue), two accumulators of n16hmax columns + two A slots must fit the tensor-memory allocation
        if (smem_scratch) {
            const size_t maps_off = ((size_t)a.max_rr * 12 + 127) & ~(size_t)127;
            const size_t sq_bytes = 2 * (size_t)tg.npanels * (size_t)tg.wrows * 16;
            const bool fits_smem = maps_off + sq_bytes <= scratch;
            const bool fits_tmem = 2 * tg.n16hmax + 8 <= tg.nacc * tg.n16max + (tg.nslot - 2) * tg.slotc && tg.n16hmax <= tg.wrows;
            const char *e = getenv("SID_TC_MMA_SUMS");
            if (fits_smem && fits_tmem && !(e && e[0] == '0')) {
                tg.mma_sums = 1; tg.sq_off = (int)maps_off; tg.wsq_off = (int)scratch;
            }
        }
        if (smem > (size_t)ctx->max_smem_optin - 4096 - (smem_scratch ? 0 : 8192)) use_tc = false;      // window too large: legacy kernels
                                                                          // (the global-scratch variant has 8 KB more static shared memory)
    }
    if (!use_tc) {
    // window staging by TMA when the padded window (+15 bytes: the box must start 16-byte aligned) fits one
    // box of <= 256 x 256 bytes
    memset(&tmap, 0, sizeof tmap);
    a.tma = 0;
    int wpw = pm_window_pitch_words(Wmax, imma);
    {
        const char *e = getenv("SID_PM_TMA");
        const int wpw_tma = pm_window_pitch_words(Wmax + 15, imma);
        if (!(e && e[0] == '0') && make_window_tensor_map(ctx, &tmap, wpw_tma * 4, Wmax)) {
            a.tma = 1; a.tma_wpw = wpw_tma; a.tma_rows = Wmax;
            wpw = wpw_tma;
        }
    }
    a.win_words = (Wmax + (imma ? PM_IMMA_ROW_SLACK : 0)) * wpw + PM_WIN_SLACK;
    const int nw = (s + 3) / 4;
    if (nw == 9) variant = 0; else if (nw == 13) variant = 1; else variant = 2;   // dp4a kernel specialisation by template width in words
    if (imma) {
        a.nc = (s + 7 + 31) / 32;
        a.tpw = (32 * a.nc + 16) / 4;
        a.tpl_off = 8;
        a.ab = std::min(n_angles, PM_IMMA_AB);
    } else {
        a.nc = 0;
        a.tpw = variant == 2 ? (s + 15) / 16 * 4 : (nw + 3) / 4 * 4;
        a.tpl_off = 0;
        a.ab = std::min(n_angles, PM_MAX_AB);
    }
    // Shared memory: window + templates (+ the per-point scratch when it all fits in a third of an SM).
    const size_t smem_cap_fast = 75 * 1024;
    auto base_smem = [&](int ab) { return ((size_t)a.win_words + (size_t)ab * s * a.tpw) * 4; };
    for (int ab = a.ab; ab >= 1 && !smem_scratch; --ab) {
        const size_t need = base_smem(ab) + pm_scratch_bytes(a.max_rr, a.max_hrw, ab, smth, split_tail);
        // fewer resident templates only pays if it keeps all angles in <= the same number of batches
        if (need <= smem_cap_fast && (n_angles + ab - 1) / ab == (n_angles + a.ab - 1) / a.ab) {
            smem_scratch = true; a.ab = ab; smem = need;
        }
    }
    if (getenv("SID_PM_GLOBAL_SCRATCH")) smem_scratch = false;
    if (!smem_scratch) {
        smem = base_smem(a.ab);
        while (smem > (size_t)ctx->max_smem_optin && a.ab > 1) { --a.ab; smem = base_smem(a.ab); }
        if (smem > (size_t)ctx->max_smem_optin)
            return fail(ctx, SID_EUNSUPPORTED, "search window too large for the fused kernel (border too big)");
    }
    }

    // CTA size: fill the CTA with thread tiles (dp4a) / warp tiles (IMMA) of a typical point
    const int Rt = 2 * typ_border + (Wmax - 2 * max_border) - s + 1;
    int threads = PM_THREADS;
    if (use_tc) {
        threads = TC_THREADS;
    } else if (imma) {
        const int warp_tiles = ((Rt + 15) / 16) * ((Rt + 23) / 24);
        double best_eff = -1.0;
        for (int w = 4; w <= PM_IMMA_THREADS / 32; ++w) {
            const double eff = (double)warp_tiles / (double)((warp_tiles + w - 1) / w * w);
            if (eff >= best_eff - 1e-9) { best_eff = eff; threads = 32 * w; }
        }
    } else {
        const int tx = pm_pick_tx(Rt);
        const int ncg = (((Rt + 3) >> 2) + tx - 1) / tx;
        const int nbatch = (n_angles + a.ab - 1) / a.ab;
        const int per = (n_angles + nbatch - 1) / nbatch;
        const long long ntiles = (long long)per * ncg * Rt * 4;
        double best_eff = -1.0;
        for (int bs = 128; bs <= PM_THREADS; bs += 32) {
            const double eff = (double)ntiles / (double)((ntiles + bs - 1) / bs * bs);
            if (eff >= best_eff - 1e-9) { best_eff = eff; threads = bs; }
        }
    }
    if (const char *e = use_tc ? nullptr : getenv("SID_PM_THREADS")) {
        const int v = atoi(e);
        if (v >= 32 && v <= (imma ? PM_IMMA_THREADS : PM_THREADS) && v % 32 == 0) threads = v;
    }

    const void *ktab[8] = {(const void *)pm_points_kernel<9, false, false>, (const void *)pm_points_kernel<13, false, false>,
                           (const void *)pm_points_kernel<0, false, false>, (const void *)pm_points_kernel<9, true, false>,
                           (const void *)pm_points_kernel<13, true, false>, (const void *)pm_points_kernel<0, true, false>,
                           (const void *)pm_points_kernel<0, false, true>, (const void *)pm_points_kernel<0, true, true>};
    const int kidx = use_tc ? (smem_scratch ? 9 : 8) : imma ? (smem_scratch ? 7 : 6) : variant + (smem_scratch ? 3 : 0);
    const void *kfn = use_tc ? (smem_scratch ? (const void *)pm_tc_kernel<true> : (const void *)pm_tc_kernel<false>) : ktab[kidx];
    if (int arc = allow_max_smem(ctx, kfn)) return arc;
    int occ = 0;
    CU(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, kfn, threads, smem));
    if (occ < 1) {
        char msg[256];
        snprintf(msg, sizeof msg, "kernel does not fit on an SM (variant %d, %d threads, %zu B dynamic smem, img_size %d, max border %d, "
                 "%d angles/batch, tma %d)", kidx, threads, smem, s, max_border, a.ab, a.tma);
        return fail(ctx, SID_ECUDA, msg);
    }
    if (getenv("SID_DEBUG"))
        fprintf(stderr, "[sid] launch: tc=%d kidx=%d threads=%d smem=%zu occ(api)=%d tmem_cols=%d nab=%d ks=%d nacc=%d nb8=%d slotc=%d nslot=%d n16max=%d wrows=%d npanels=%d mma_sums=%d\n",
                (int)use_tc, kidx, threads, smem, occ, tg.tmem_cols, tg.nab, tg.ks, tg.nacc, tg.nb8, tg.slotc, tg.nslot, tg.n16max, tg.wrows, tg.npanels, tg.mma_sums);
    if (use_tc) {
        // The occupancy API answers 1 for a kernel that executes tcgen05.alloc with a run-time column count (it has to
        // assume all 512 columns); residency is really bounded by registers, shared memory and the columns we ask for.
        cudaFuncAttributes fa;
        CU(cudaFuncGetAttributes(&fa, kfn));
        int smem_sm = 0, regs_sm = 0;
        CU(cudaDeviceGetAttribute(&smem_sm, cudaDevAttrMaxSharedMemoryPerMultiprocessor, ctx->device));
        CU(cudaDeviceGetAttribute(&regs_sm, cudaDevAttrMaxRegistersPerMultiprocessor, ctx->device));
        const int by_smem = (int)((size_t)smem_sm / (smem + fa.sharedSizeBytes + 1024));
        const int by_regs = regs_sm / std::max(1, fa.numRegs * threads);
        occ = std::max(1, std::min(std::min(by_smem, by_regs), 512 / tg.tmem_cols));
    }
    if (const char *e = getenv("SID_PM_OCC")) { const int v = atoi(e); if (v >= 1 && v <= occ) occ = v; }
    long long grid = (long long)ctx->sm_count * occ;
    if (grid > n) grid = n;
    if (grid < 1) grid = 1;

    a.scratch_per_cta = pm_scratch_bytes(a.max_rr, a.max_hrw, a.ab, smth, split_tail);
    int rc = SID_OK;
    if (!smem_scratch) {
        const size_t per_slot = (size_t)a.scratch_per_cta * (size_t)ctx->sm_count * (size_t)occ;
        rc = reserve(ctx, ctx->scratch, per_slot * 2);
        if (rc) return rc;
        a.scratch = (unsigned char *)ctx->scratch.p + per_slot * (size_t)slot;
    }
    rc = reserve(ctx, ctx->counter, 256);
    if (rc) return rc;
    a.counter = (unsigned int *)ctx->counter.p + 16 * slot;
    CU(cudaMemsetAsync(a.counter, 0, sizeof(unsigned int), st));

    a.split_tail = split_tail ? 1 : 0;
    if (split_tail) {
        const size_t cap_n = (size_t)std::max<long long>(n + tail_off, ctx->tail_hint_n);
        a.tail_stride = (a.max_rr + 3) & ~3;
        const size_t region_stride = std::max(ctx->tail_region_stride, (size_t)a.tail_stride);
        if ((rc = reserve(ctx, ctx->tail_maps, cap_n * region_stride * sizeof(float)))) return rc;
        if ((rc = reserve(ctx, ctx->tail_recs, cap_n * sizeof(PmTailRec)))) return rc;
        a.tail_maps = (float *)ctx->tail_maps.p + (size_t)tail_off * region_stride;
        a.tail_recs = (PmTailRec *)ctx->tail_recs.p + tail_off;
    }
    void *params_legacy[] = {(void *)&a, (void *)&tmap};
    void *params_tc[] = {(void *)&a, (void *)&tg, (void *)&tmap};
    void **params = use_tc ? params_tc : params_legacy;
    if (!ctx->k_ev[0]) { CU(cudaEventCreate(&ctx->k_ev[0])); CU(cudaEventCreate(&ctx->k_ev[1])); }
    if (!ctx->k_ev_hold) CU(cudaEventRecord(ctx->k_ev[0], st));
    CU(cudaLaunchKernel(kfn, dim3((unsigned)grid), dim3((unsigned)threads), params, smem, st));
    CU(cudaEventRecord(ctx->k_ev[1], st));
    ctx->k_ev_valid = true;
    ctx->k_name = use_tc ? "sid::pm_tc_kernel" : imma ? "sid::pm_points_kernel<imma>" : "sid::pm_points_kernel<dp4a>";
    ctx->launches += 1;
    if (split_tail) {
        const size_t tsm = pm_tail_smem_bytes(a.max_rr, smth);
        if (int arc = allow_max_smem(ctx, (const void *)pm_tail_kernel)) return arc;
        pm_tail_kernel<<<(unsigned)n, pm_tail_threads(a.max_rr), tsm, st>>>(a);
        ctx->launches += 1;
        CU(cudaGetLastError());
    }
    return SID_OK;
}

}  // namespace

extern "C" {

const char *sid_version(void) { return "sea_ice_drift_b200 0.3 (sm_100a, exact-integer u8 tcgen05 MCC, warp-specialised pipeline)"; }

int sid_create(sid_ctx **out, int device) {
    if (!out) return SID_EINVAL;
    *out = nullptr;
    sid_ctx *ctx = new (std::nothrow) sid_ctx();
    if (!ctx) return SID_ENOMEM;
    cudaError_t e = cudaSetDevice(device);
    if (e == cudaSuccess) e = cudaStreamCreateWithFlags(&ctx->own_stream, cudaStreamNonBlocking);
    cudaDeviceProp prop;
    if (e == cudaSuccess) e = cudaGetDeviceProperties(&prop, device);
    if (e != cudaSuccess) {
        cudaGetLastError();
        delete ctx;
        return SID_ECUDA;
    }
    ctx->device = device;
    ctx->stream = ctx->own_stream;
    ctx->sm_count = prop.multiProcessorCount;
    ctx->max_smem_optin = (int)prop.sharedMemPerBlockOptin;
    *out = ctx;
    return SID_OK;
}

void sid_destroy(sid_ctx *ctx) {
    if (!ctx) return;
    cudaSetDevice(ctx->device);
    cudaStreamSynchronize(ctx->stream);
    DevBuf *bufs[] = {&ctx->img1, &ctx->img2, &ctx->pts, &ctx->order, &ctx->out, &ctx->status,
                      &ctx->angles, &ctx->scratch, &ctx->counter, &ctx->misc, &ctx->tail_maps, &ctx->tail_recs, &ctx->epi, &ctx->fg};
    for (DevBuf *b : bufs) if (b->p && b->owned) cudaFree(b->p);
    if (ctx->pin) cudaFreeHost(ctx->pin);
    if (ctx->stage) cudaFreeHost(ctx->stage);
    delete ctx->pool;
    for (cudaEvent_t e : ctx->stage_event) if (e) cudaEventDestroy(e);
    for (cudaEvent_t e : ctx->k_ev) if (e) cudaEventDestroy(e);
    for (cudaEvent_t e : ctx->slot_event) if (e) cudaEventDestroy(e);
    for (cudaStream_t st : ctx->band_stream) if (st) cudaStreamDestroy(st);
    if (ctx->copy_stream) {
        cudaStreamDestroy(ctx->copy_stream);
        for (cudaEvent_t e : ctx->band_event) if (e) cudaEventDestroy(e);
    }
    if (ctx->own_stream) cudaStreamDestroy(ctx->own_stream);
    delete ctx;
}

const char *sid_last_error(const sid_ctx *ctx) { return ctx ? ctx->err.c_str() : "null context"; }

int sid_set_stream(sid_ctx *ctx, void *cuda_stream) {
    if (!ctx) return SID_EINVAL;
    ctx->stream = cuda_stream ? (cudaStream_t)cuda_stream : ctx->own_stream;
    return SID_OK;
}

int sid_synchronize(sid_ctx *ctx) {
    if (!ctx) return SID_EINVAL;
    CU(cudaSetDevice(ctx->device));
    CU(cudaStreamSynchronize(ctx->stream));
    return SID_OK;
}

int64_t sid_launch_count(const sid_ctx *ctx) { return ctx ? ctx->launches : 0; }

const char *sid_last_kernel_name(const sid_ctx *ctx) { return ctx ? ctx->k_name : ""; }

double sid_last_kernel_ms(sid_ctx *ctx) {
    if (!ctx || !ctx->k_ev_valid) return -1.0;
    float ms = 0.f;
    if (cudaEventSynchronize(ctx->k_ev[1]) != cudaSuccess || cudaEventElapsedTime(&ms, ctx->k_ev[0], ctx->k_ev[1]) != cudaSuccess) {
        cudaGetLastError();
        return -1.0;
    }
    return (double)ms;
}

static int set_pair_impl(sid_ctx *ctx, const uint8_t *img1, int rows1, int cols1, int64_t pitch1,
                         const uint8_t *img2, int rows2, int cols2, int64_t pitch2, cudaMemcpyKind kind) {
    if (!ctx) return SID_EINVAL;
    CU(cudaSetDevice(ctx->device));
    ctx->have_pair = false;
    int rc = upload_image(ctx, ctx->img1, ctx->pitch1, img1, rows1, cols1, pitch1, kind);
    if (rc) return rc;
    rc = upload_image(ctx, ctx->img2, ctx->pitch2, img2, rows2, cols2, pitch2, kind);
    if (rc) return rc;
    ctx->rows1 = rows1; ctx->cols1 = cols1; ctx->rows2 = rows2; ctx->cols2 = cols2;
    ctx->have_pair = true;
    return SID_OK;
}

int sid_set_pair(sid_ctx *ctx, const uint8_t *img1, int rows1, int cols1, int64_t pitch1,
                 const uint8_t *img2, int rows2, int cols2, int64_t pitch2) {
    return set_pair_impl(ctx, img1, rows1, cols1, pitch1, img2, rows2, cols2, pitch2, cudaMemcpyHostToDevice);
}
int sid_set_pair_device(sid_ctx *ctx, const uint8_t *img1, int rows1, int cols1, int64_t pitch1,
                        const uint8_t *img2, int rows2, int cols2, int64_t pitch2) {
    return set_pair_impl(ctx, img1, rows1, cols1, pitch1, img2, rows2, cols2, pitch2, cudaMemcpyDeviceToDevice);
}

int sid_pair_layout(int rows, int cols, int64_t *pitch, int64_t *bytes) {
    if (rows <= 0 || cols <= 0 || !pitch || !bytes) return SID_EINVAL;
    *pitch = padded_pitch(cols);
    *bytes = *pitch * (int64_t)rows + (int64_t)IMG_TAIL_SLACK;
    return SID_OK;
}

int sid_adopt_pair_device(sid_ctx *ctx, uint8_t *d_img1, int rows1, int cols1, int64_t pitch1, int64_t bytes1,
                          uint8_t *d_img2, int rows2, int cols2, int64_t pitch2, int64_t bytes2) {
    if (!ctx) return SID_EINVAL;
    if (!d_img1 || !d_img2 || rows1 <= 0 || cols1 <= 0 || rows2 <= 0 || cols2 <= 0)
        return fail(ctx, SID_EINVAL, "bad image arguments");
    if (pitch1 < padded_pitch(cols1) || pitch2 < padded_pitch(cols2) || (pitch1 & 15) || (pitch2 & 15) ||
        ((uintptr_t)d_img1 & 255) || ((uintptr_t)d_img2 & 255))
        return fail(ctx, SID_EINVAL, "adopted images need the layout of sid_pair_layout (pitch, 256-byte aligned base)");
    if (bytes1 < pitch1 * (int64_t)rows1 + (int64_t)IMG_TAIL_SLACK || bytes2 < pitch2 * (int64_t)rows2 + (int64_t)IMG_TAIL_SLACK)
        return fail(ctx, SID_EINVAL, "adopted images need the tail slack of sid_pair_layout");
    CU(cudaSetDevice(ctx->device));
    if (ctx->have_pair && !ctx->img1.owned && !ctx->img2.owned && ctx->img1.p == d_img1 && ctx->img2.p == d_img2 &&
        ctx->rows1 == rows1 && ctx->cols1 == cols1 && ctx->pitch1 == pitch1 && ctx->rows2 == rows2 && ctx->cols2 == cols2 &&
        ctx->pitch2 == pitch2)
        return SID_OK;                                          // the same buffers again (refilled in place by the caller)
    CU(cudaStreamSynchronize(ctx->stream));                    // nothing in flight still reads the old pair
    DevBuf *bufs[2] = {&ctx->img1, &ctx->img2};
    for (DevBuf *b : bufs) { if (b->p && b->owned) cudaFree(b->p); b->p = nullptr; b->cap = 0; }
    ctx->img1.p = d_img1; ctx->img1.cap = (size_t)bytes1; ctx->img1.owned = false;
    ctx->img2.p = d_img2; ctx->img2.cap = (size_t)bytes2; ctx->img2.owned = false;
    ctx->rows1 = rows1; ctx->cols1 = cols1; ctx->pitch1 = pitch1;
    ctx->rows2 = rows2; ctx->cols2 = cols2; ctx->pitch2 = pitch2;
    ctx->have_pair = true;
    return SID_OK;
}

int sid_upload_rows(sid_ctx *ctx, uint8_t *d_dst, int64_t dst_pitch, const uint8_t *src, int64_t src_pitch, int cols, int rows) {
    if (!ctx) return SID_EINVAL;
    if (rows == 0) return SID_OK;
    if (!d_dst || !src || cols <= 0 || rows < 0 || dst_pitch < cols || src_pitch < cols) return fail(ctx, SID_EINVAL, "bad arguments");
    CU(cudaSetDevice(ctx->device));
    CU(cudaMemcpy2DAsync(d_dst, (size_t)dst_pitch, src, (size_t)src_pitch, (size_t)cols, (size_t)rows, cudaMemcpyHostToDevice, ctx->stream));
    return SID_OK;
}

static int upload_angles(sid_ctx *ctx, int n_angles, const double *angles, const double *angle_tab,
                         const double **d_angles, const double **d_tab, cudaStream_t st = nullptr) {
    const size_t bytes = (size_t)n_angles * 5 * sizeof(double);
    int rc = reserve(ctx, ctx->angles, bytes);
    if (rc) return rc;
    std::vector<double> h((size_t)n_angles * 5);
    for (int k = 0; k < n_angles; ++k) h[k] = angles ? angles[k] : (double)k;
    memcpy(h.data() + n_angles, angle_tab, (size_t)n_angles * 4 * sizeof(double));
    // small pageable copy: the runtime stages it before returning, so `h` may go out of scope
    CU(cudaMemcpyAsync(ctx->angles.p, h.data(), bytes, cudaMemcpyHostToDevice, st ? st : ctx->stream));
    *d_angles = (const double *)ctx->angles.p;
    *d_tab = *d_angles + n_angles;
    return SID_OK;
}

namespace {

// Largest border (<= max_border) for which launch_pm would pick pm_ws_kernel (its geometry, its shared-memory layout and
// the split tail all fit); -1 if none.  Used to send the points of a call to the kernels by border class.
int ws_border_limit(sid_ctx *ctx, int s, int n_angles, unsigned flags, int max_border) {
    const char *path_env = getenv("SID_PM_PATH");
    const bool want_ws = path_env ? strcmp(path_env, "ws") == 0 : SID_WS_DEFAULT != 0;
    if (!want_ws) return -1;
    if (const char *e = getenv("SID_PM_SPLIT_TAIL")) if (e[0] == '0') return -1;
    const bool smth = (flags & SID_HES_SMTH) != 0;
    const int hws = s / 2;
    for (int b = std::min(max_border, 40); b >= 1; --b) {
        const int Wmax = 2 * hws + 2 * b + 1, Rmax = Wmax - s + 1;
        if (Rmax < 2) break;
        if (pm_tail_smem_bytes(Rmax * Rmax, smth) > tail_split_limit()) continue;
        PmWsCfg g;
        memset(&g, 0, sizeof g);
        // the launch pads the angle planes for ITS typical map width (<= 31 words each): try the worst case, so that a
        // border this function accepts is never refused by the launch
        // (the smallest layout: two sets of window statistics)
        bool ok = pm_ws_geometry(s, Rmax, Wmax, n_angles, Rmax * Rmax, g, 0, 2) && (size_t)g.smem_bytes + 2048 <= (size_t)ctx->max_smem_optin;
        if (ok && (size_t)g.smem_bytes + 2048 + 31 * 4 * 2 * 3 > (size_t)ctx->max_smem_optin) {
            for (int rt = 2; rt <= Rmax && ok; ++rt)
                ok = pm_ws_geometry(s, Rmax, Wmax, n_angles, Rmax * Rmax, g, rt, 2) && (size_t)g.smem_bytes + 2048 <= (size_t)ctx->max_smem_optin;
        }
        if (ok) return b;
    }
    return -1;
}

// Device-resident callers (sid_run_device): split the point indices by border class on the device
__global__ void partition_by_border_kernel(const double *__restrict__ border, int n, int limit, int *__restrict__ order_small,
                                           int *__restrict__ order_big, unsigned *__restrict__ counts) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    bool small = false, valid = i < n;
    if (valid) {
        const double b = border[i];
        int v = 0;
        if (isfinite(b) && b >= 0.0 && b < 4096.0) v = (int)ceil(b);
        small = v <= limit;
    }
    const unsigned m_small = __ballot_sync(0xffffffffu, valid && small), m_big = __ballot_sync(0xffffffffu, valid && !small);
    const int lane = threadIdx.x & 31;
    unsigned base_s = 0, base_b = 0;
    if (lane == 0) {
        if (m_small) base_s = atomicAdd(&counts[0], (unsigned)__popc(m_small));
        if (m_big) base_b = atomicAdd(&counts[1], (unsigned)__popc(m_big));
    }
    base_s = __shfl_sync(0xffffffffu, base_s, 0); base_b = __shfl_sync(0xffffffffu, base_b, 0);
    const unsigned below = (1u << lane) - 1u;
    if (valid && small) order_small[base_s + __popc(m_small & below)] = i;
    if (valid && !small) order_big[base_b + __popc(m_big & below)] = i;
}

// true when `p` is ordinary pageable host memory (cudaMemcpyAsync from it is a slow, synchronous staged copy)
bool is_pageable(const void *p) {
    cudaPointerAttributes at;
    if (cudaPointerGetAttributes(&at, p) != cudaSuccess) { cudaGetLastError(); return true; }
    return at.type == cudaMemoryTypeUnregistered;
}

// rows x width bytes from (src, spitch) to (dst, dpitch), split over `nthreads` host threads
// One row: streaming (non-temporal) 16-byte stores when the destination allows -- the staging buffer is written once and
// read by the DMA engine, so allocating its lines in the cache only costs a read-for-ownership per line
inline void copy_row_stream(uint8_t *dst, const uint8_t *src, size_t width) {
    if ((reinterpret_cast<uintptr_t>(dst) & 15u) != 0 || width < 256) { memcpy(dst, src, width); return; }
    const size_t n16 = width / 16;
    const __m128i *s16 = reinterpret_cast<const __m128i *>(src);
    __m128i *d16 = reinterpret_cast<__m128i *>(dst);
    size_t i = 0;
    for (; i + 4 <= n16; i += 4) {
        const __m128i v0 = _mm_loadu_si128(s16 + i), v1 = _mm_loadu_si128(s16 + i + 1);
        const __m128i v2 = _mm_loadu_si128(s16 + i + 2), v3 = _mm_loadu_si128(s16 + i + 3);
        _mm_stream_si128(d16 + i, v0); _mm_stream_si128(d16 + i + 1, v1);
        _mm_stream_si128(d16 + i + 2, v2); _mm_stream_si128(d16 + i + 3, v3);
    }
    for (; i < n16; ++i) _mm_stream_si128(d16 + i, _mm_loadu_si128(s16 + i));
    if (width & 15u) memcpy(dst + n16 * 16, src + n16 * 16, width & 15u);
}

// Rows of up to two images into the pinned staging slot, split over the pool's threads by row count
struct RowCopy { uint8_t *dst; size_t dpitch; const uint8_t *src; size_t spitch, width; int rows; };
void copy_rows_pool(HostPool *pool, const RowCopy *jobs, int njobs) {
    long long total = 0;
    for (int j = 0; j < njobs; ++j) total += std::max(0, jobs[j].rows);
    if (total <= 0) return;
    const int parts = pool ? pool->size() : 1;
    const std::function<void(int)> work = [&](int part) {
        long long lo = total * part / parts, hi = total * (part + 1) / parts, base = 0;
        for (int j = 0; j < njobs; ++j) {
            const RowCopy &c = jobs[j];
            const long long r0 = std::max(lo, base) - base, r1 = std::min(hi, base + c.rows) - base;
            for (long long r = r0; r < r1; ++r) copy_row_stream(c.dst + (size_t)r * c.dpitch, c.src + (size_t)r * c.spitch, c.width);
            base += c.rows;
        }
        _mm_sfence();                    // the streaming stores are globally visible before the DMA is enqueued
    };
    if (pool && parts > 1) pool->run(work); else work(0);
}

struct HostPair {
    const uint8_t *img1, *img2;
    int rows1, cols1, rows2, cols2;
    long long pitch1, pitch2;
};

constexpr int MAX_BANDS = 16;

// Body of sid_run and sid_run_pair.  With `pair` the images are uploaded in row bands on a second
// stream while the fused kernel already works on the points whose windows and templates lie
// entirely inside the bands that have arrived (PCIe copy overlapped with compute).
int run_host(sid_ctx *ctx, const HostPair *pair, int64_t n, const double *c1, const double *r1, const double *c2fg,
             const double *r2fg, const double *border, int img_size, int n_angles, const double *angles,
             const double *angle_tab, int rot_order, unsigned flags, int mtype, double *out, int *status) {
    int rc = check_common(ctx, img_size, n_angles, angle_tab, rot_order, mtype);
    if (rc) return rc;
    if (!pair && !ctx->have_pair) return fail(ctx, SID_ENOPAIR, "sid_set_pair has not been called");
    if (n < 0 || (n > 0 && (!c1 || !r1 || !c2fg || !r2fg || !border)) || !angles)
        return fail(ctx, SID_EINVAL, "null point array");
    if (n > 0x7fffffffLL) return fail(ctx, SID_EINVAL, "too many points");
    CU(cudaSetDevice(ctx->device));
    if (pair) {
        if (!pair->img1 || !pair->img2 || pair->rows1 <= 0 || pair->cols1 <= 0 || pair->rows2 <= 0 || pair->cols2 <= 0 ||
            pair->pitch1 < pair->cols1 || pair->pitch2 < pair->cols2)
            return fail(ctx, SID_EINVAL, "bad image arguments");
        if (n == 0) return sid_set_pair(ctx, pair->img1, pair->rows1, pair->cols1, pair->pitch1, pair->img2, pair->rows2,
                                        pair->cols2, pair->pitch2);
    }
    if (n == 0) return SID_OK;

    // ---- bands (upload granularity); a single band means "everything is already resident".  The bands TAPER: while the
    //      upload runs the kernels of band k hide behind the copy of band k + 1, but the kernels of the LAST band run after
    //      the last byte has arrived -- so the last bands are short (weights below; SID_BAND_PLAN="w0,w1,..." overrides,
    //      SID_BANDS=n gives n equal bands).  Boundaries are multiples of 64 rows.
    const int rows1 = pair ? pair->rows1 : ctx->rows1, rows2 = pair ? pair->rows2 : ctx->rows2;
    const int max_rows = std::max(rows1, rows2);
    int nbands = 1;
    std::vector<double> weights;
    if (pair) {
        if (max_rows >= 8192) weights = {6, 6, 6, 6, 6, 5, 4, 3, 2, 1};
        else weights.assign((size_t)std::max(1, std::min(8, max_rows / 512)), 1.0);
        if (const char *e = getenv("SID_BANDS")) weights.assign((size_t)std::max(1, std::min(MAX_BANDS, atoi(e))), 1.0);
        if (const char *e = getenv("SID_BAND_PLAN")) {
            std::vector<double> w;
            for (const char *q = e; *q && (int)w.size() < MAX_BANDS;) {
                char *end = nullptr;
                const double v = strtod(q, &end);
                if (end == q) break;
                if (v > 0.0) w.push_back(v);
                q = (*end == ',') ? end + 1 : end;
            }
            if (!w.empty()) weights = w;
        }
        nbands = (int)weights.size();
    }
    std::vector<int> band_lo((size_t)nbands + 1, 0);
    {
        double total = 0.0, acc = 0.0;
        for (double w : weights) total += w;
        for (int k = 1; k < nbands; ++k) {
            acc += weights[(size_t)k - 1];
            int y = (int)((double)max_rows * acc / total / 64.0 + 0.5) * 64;
            band_lo[(size_t)k] = std::max(band_lo[(size_t)k - 1], std::min(max_rows, y));
        }
        band_lo[(size_t)nbands] = max_rows;
    }
    int band_rows = 0;                                          // tallest band (sizes the staging slots)
    for (int k = 0; k < nbands; ++k) band_rows = std::max(band_rows, band_lo[(size_t)k + 1] - band_lo[(size_t)k]);
    std::vector<unsigned char> band_of64((size_t)max_rows / 64 + 2, (unsigned char)(nbands - 1));
    for (int k = 0, j = 0; k < nbands; ++k)
        for (; j * 64 < band_lo[(size_t)k + 1] && (size_t)j < band_of64.size(); ++j) band_of64[(size_t)j] = (unsigned char)k;

    // ---- image upload in bands on the copy stream
    if (pair) {
        ctx->have_pair = false;
        if (!ctx->copy_stream) {
            CU(cudaStreamCreateWithFlags(&ctx->copy_stream, cudaStreamNonBlocking));
            for (int k = 0; k <= MAX_BANDS; ++k) CU(cudaEventCreateWithFlags(&ctx->band_event[k], cudaEventDisableTiming));
        }
        ctx->pitch1 = padded_pitch(pair->cols1);
        ctx->pitch2 = padded_pitch(pair->cols2);
        const size_t b1 = (size_t)ctx->pitch1 * pair->rows1 + IMG_TAIL_SLACK, b2 = (size_t)ctx->pitch2 * pair->rows2 + IMG_TAIL_SLACK;
        const bool fresh1 = b1 > ctx->img1.cap || !ctx->img1.owned, fresh2 = b2 > ctx->img2.cap || !ctx->img2.owned;
        if ((rc = reserve(ctx, ctx->img1, b1))) return rc;
        if ((rc = reserve(ctx, ctx->img2, b2))) return rc;
        // the copy stream must not overwrite images an earlier launch on the compute stream still reads
        CU(cudaEventRecord(ctx->band_event[MAX_BANDS], ctx->stream));
        CU(cudaStreamWaitEvent(ctx->copy_stream, ctx->band_event[MAX_BANDS], 0));
        if (fresh1) CU(cudaMemsetAsync(ctx->img1.p, 0, ctx->img1.cap, ctx->copy_stream));
        if (fresh2) CU(cudaMemsetAsync(ctx->img2.p, 0, ctx->img2.cap, ctx->copy_stream));
        ctx->rows1 = pair->rows1; ctx->cols1 = pair->cols1; ctx->rows2 = pair->rows2; ctx->cols2 = pair->cols2;
    }

    // Pageable host images (plain NumPy arrays): each band goes through a pinned double buffer, filled by a few host
    // threads while the previous band's DMA and kernels run; band k's kernels are launched right behind its copy.
    bool staged = false;
    int copy_threads = 1;
    if (pair) {
        staged = is_pageable(pair->img1) || is_pageable(pair->img2);
        if (const char *e = getenv("SID_STAGED_UPLOAD")) staged = e[0] != '0';
        copy_threads = (int)std::max(1u, std::min(8u, std::thread::hardware_concurrency() / 2));
        if (const char *e = getenv("SID_UPLOAD_THREADS")) copy_threads = std::max(1, std::min(64, atoi(e)));
    }
    const size_t slot_bytes = pair ? ((size_t)band_rows * ((size_t)pair->cols1 + (size_t)pair->cols2) + 255) / 256 * 256 : 0;
    if (staged) {
        if (2 * slot_bytes > ctx->stage_cap) {
            if (ctx->stage) { cudaFreeHost(ctx->stage); ctx->stage = nullptr; ctx->stage_cap = 0; }
            if (cudaMallocHost(&ctx->stage, 2 * slot_bytes) != cudaSuccess) { cudaGetLastError(); return fail(ctx, SID_ENOMEM, "pinned staging allocation failed"); }
            ctx->stage_cap = 2 * slot_bytes;
        }
        for (int k = 0; k < 2; ++k) if (!ctx->stage_event[k]) CU(cudaEventCreateWithFlags(&ctx->stage_event[k], cudaEventDisableTiming));
    }
    auto enqueue_band_copy = [&](int k) -> int {
        const int ya = band_lo[(size_t)k], nrows = band_lo[(size_t)k + 1] - ya;
        const int n1 = std::min(nrows, pair->rows1 - ya), n2 = std::min(nrows, pair->rows2 - ya);
        const uint8_t *s1 = pair->img1 + (size_t)ya * pair->pitch1, *s2 = pair->img2 + (size_t)ya * pair->pitch2;
        size_t sp1 = (size_t)pair->pitch1, sp2 = (size_t)pair->pitch2;
        if (staged) {
            const int slot = k & 1;
            if (k >= 2) CU(cudaEventSynchronize(ctx->stage_event[slot]));      // the DMA of band k-2 has drained this slot
            uint8_t *d1 = (uint8_t *)ctx->stage + (size_t)slot * slot_bytes, *d2 = d1 + (size_t)band_rows * pair->cols1;
            if (copy_threads > 1 && (!ctx->pool || ctx->pool->size() != copy_threads)) {
                delete ctx->pool;
                ctx->pool = new HostPool(copy_threads - 1);
            }
            const RowCopy jobs[2] = {{d1, (size_t)pair->cols1, s1, sp1, (size_t)pair->cols1, std::max(n1, 0)},
                                     {d2, (size_t)pair->cols2, s2, sp2, (size_t)pair->cols2, std::max(n2, 0)}};
            copy_rows_pool(copy_threads > 1 ? ctx->pool : nullptr, jobs, 2);
            s1 = d1; s2 = d2; sp1 = (size_t)pair->cols1; sp2 = (size_t)pair->cols2;
        }
        if (n1 > 0)
            CU(cudaMemcpy2DAsync((char *)ctx->img1.p + (size_t)ya * ctx->pitch1, (size_t)ctx->pitch1, s1, sp1, (size_t)pair->cols1,
                                 (size_t)n1, cudaMemcpyHostToDevice, ctx->copy_stream));
        if (n2 > 0)
            CU(cudaMemcpy2DAsync((char *)ctx->img2.p + (size_t)ya * ctx->pitch2, (size_t)ctx->pitch2, s2, sp2, (size_t)pair->cols2,
                                 (size_t)n2, cudaMemcpyHostToDevice, ctx->copy_stream));
        if (staged) CU(cudaEventRecord(ctx->stage_event[k & 1], ctx->copy_stream));
        CU(cudaEventRecord(ctx->band_event[k], ctx->copy_stream));
        return SID_OK;
    };
    // band 0 goes out before the host-side preparation below (point order, small uploads), which then overlaps its DMA;
    // the small uploads queue behind it on the copy engine (the first kernel needs band 0 anyway), the other bands behind them
    if (pair && !staged && (rc = enqueue_band_copy(0))) return rc;

    // ---- borders: bound, typical value, and the processing order: band of the last image row a point
    //      touches first, then largest window first (counting sort on the combined key)
    int max_border = 0;
    std::vector<int> ib((size_t)n), band((size_t)n, 0);
    const double ext = 0.5 * std::sqrt(2.0) * (img_size + 2) + 3.0;      // template reach around (c1, r1) incl. rounding
    const int hws = img_size / 2;
    for (int64_t i = 0; i < n; ++i) {
        const double b = border[i];
        int v = 0;
        if (std::isfinite(b) && b >= 0.0 && b < 4096.0) v = (int)std::ceil(b);   // fractional borders: the window can be one pixel wider
        ib[(size_t)i] = v;
        max_border = std::max(max_border, v);
        if (nbands > 1) {
            double need = 0.0;
            if (std::isfinite(r1[i]) && std::isfinite(r2fg[i])) need = std::max(r1[i] + ext, r2fg[i] + hws + v + 2.0);
            band[(size_t)i] = need <= 0.0 ? 0 : (int)band_of64[(size_t)std::min((double)(band_of64.size() - 1), need / 64.0)];
        }
    }
    const size_t nkeys = (size_t)nbands * (size_t)(max_border + 1);
    std::vector<int> hist(nkeys + 1, 0);
    auto key_of = [&](int64_t i) { return (size_t)band[(size_t)i] * (size_t)(max_border + 1) + (size_t)(max_border - ib[(size_t)i]); };
    for (int64_t i = 0; i < n; ++i) ++hist[key_of(i) + 1];
    for (size_t k = 1; k < hist.size(); ++k) hist[k] += hist[k - 1];
    std::vector<int> band_start((size_t)nbands + 1, 0);
    for (int k = 0; k <= nbands; ++k) band_start[(size_t)k] = hist[std::min(nkeys, (size_t)k * (size_t)(max_border + 1))];
    std::vector<long long> per_b((size_t)max_border + 1, 0);
    for (int64_t i = 0; i < n; ++i) ++per_b[(size_t)ib[(size_t)i]];
    auto median_border = [&](int b_lo, int b_hi) {          // median of the borders in [b_lo, b_hi] = the typical point of a class
        long long total = 0, acc = 0;
        for (int b = b_lo; b <= b_hi; ++b) total += per_b[(size_t)b];
        for (int b = b_hi; b >= b_lo; --b) { acc += per_b[(size_t)b]; if (2 * acc > total) return b; }
        return b_hi;
    };
    int typ_border = median_border(0, max_border);
    // ---- border classes.  A launch is sized by its largest border, and the fastest kernel (pm_ws_kernel) only takes small
    //      result maps: with the reference's default borders (20 ... 50 by the distance to the nearest keypoint,
    //      pmlib.py:315-322) most points sit at the minimum -- 86 % of BASELINE configs[0]'s grid with a real ORB first
    //      guess -- and a few far from any keypoint would drag the whole call onto the kernels for large maps.  So the points
    //      whose border fits pm_ws_kernel get their own launch per band, behind the larger ones (SID_PM_CLASSES=0: one launch).
    int small_max = -1, typ_small = 0, typ_big = typ_border;
    bool two_class = false;
    {
        const char *e = getenv("SID_PM_CLASSES");
        const int limit = (e && e[0] == '0') ? -1 : ws_border_limit(ctx, img_size, n_angles, flags, max_border);
        if (limit >= 0 && limit < max_border) {
            long long n_small = 0;
            for (int b = 0; b <= limit; ++b) if (per_b[(size_t)b]) { n_small += per_b[(size_t)b]; small_max = b; }
            if (n_small >= 256 && small_max >= 1) {
                two_class = true;
                typ_small = median_border(0, small_max);
                typ_big = median_border(limit + 1, max_border);
            }
        }
    }
    std::vector<int> band_mid((size_t)nbands, 0);           // first point of the small class inside each band's range
    for (int k = 0; k < nbands; ++k)
        band_mid[(size_t)k] = two_class ? hist[(size_t)k * (size_t)(max_border + 1) + (size_t)(max_border - small_max)]
                                        : band_start[(size_t)k + 1];

    const size_t pts_bytes = (size_t)n * 5 * sizeof(double);
    const size_t ord_bytes = (size_t)n * sizeof(int);
    const size_t out_bytes = (size_t)n * 5 * sizeof(double);
    const size_t st_bytes = (size_t)n * sizeof(int);
    rc = reserve_pinned(ctx, pts_bytes + ord_bytes + out_bytes + st_bytes + 64);
    if (rc) return rc;
    if ((rc = reserve(ctx, ctx->pts, pts_bytes))) return rc;
    if ((rc = reserve(ctx, ctx->order, ord_bytes))) return rc;
    if ((rc = reserve(ctx, ctx->out, out_bytes))) return rc;
    if ((rc = reserve(ctx, ctx->status, st_bytes))) return rc;

    double *hp = (double *)ctx->pin;
    memcpy(hp, c1, (size_t)n * 8); memcpy(hp + n, r1, (size_t)n * 8); memcpy(hp + 2 * n, c2fg, (size_t)n * 8);
    memcpy(hp + 3 * n, r2fg, (size_t)n * 8); memcpy(hp + 4 * n, border, (size_t)n * 8);
    int *hord = (int *)((char *)ctx->pin + pts_bytes);
    for (int64_t i = 0; i < n; ++i) hord[hist[key_of(i)]++] = (int)i;

    // ---- small uploads.  With pinned images they go on the COPY stream, between band 0 and band 1: the copy engine serves
    //      one stream's queue in order, but between streams it does not keep the issue order (measured: point arrays enqueued
    //      on the compute stream while band 0 was in flight were served after ALL bands -- 8.3 instead of 5.0 ms per pair).
    //      The staged path enqueues its bands one by one below, so there the small uploads simply go first.
    cudaStream_t up = (pair && !staged) ? ctx->copy_stream : ctx->stream;
    CU(cudaMemcpyAsync(ctx->pts.p, hp, pts_bytes, cudaMemcpyHostToDevice, up));
    CU(cudaMemcpyAsync(ctx->order.p, hord, ord_bytes, cudaMemcpyHostToDevice, up));
    const double *d_angles, *d_tab;
    if ((rc = upload_angles(ctx, n_angles, angles, angle_tab, &d_angles, &d_tab, up))) return rc;
    if (!ctx->slot_event[0])
        for (int k = 0; k < 3; ++k) CU(cudaEventCreateWithFlags(&ctx->slot_event[k], cudaEventDisableTiming));
    CU(cudaEventRecord(ctx->slot_event[0], up));                       // point / angle uploads are enqueued

    if (pair && !staged)
        for (int k = 1; k < nbands; ++k) if ((rc = enqueue_band_copy(k))) return rc;

    const double *dp = (const double *)ctx->pts.p;
    ctx->tail_hint_n = n;
    ctx->tail_region_stride = 0;
    if (two_class) {
        // one stride for the tail-map regions of both classes (launches of different bands run concurrently)
        auto rr_of = [&](int b) { const int R = 2 * (img_size / 2) + 2 * b + 1 - img_size + 1; return R * R; };
        const bool smth = (flags & SID_HES_SMTH) != 0;
        const int rr_big = rr_of(max_border), rr_small = rr_of(small_max);
        ctx->tail_region_stride = (size_t)(((pm_tail_smem_bytes(rr_big, smth) <= tail_split_limit() ? rr_big : rr_small) + 3) & ~3);
    }
    const bool two_streams = pair && nbands > 1;
    if (two_streams) {
        if (!ctx->band_stream[0])
            for (int k = 0; k < 2; ++k) CU(cudaStreamCreateWithFlags(&ctx->band_stream[k], cudaStreamNonBlocking));
        for (int k = 0; k < 2; ++k) CU(cudaStreamWaitEvent(ctx->band_stream[k], ctx->slot_event[0], 0));
    } else if (up != ctx->stream) {
        CU(cudaStreamWaitEvent(ctx->stream, ctx->slot_event[0], 0));
    }
    for (int k = 0; k < nbands; ++k) {
        if (staged && (rc = enqueue_band_copy(k))) return rc;
        // consecutive bands go to alternating streams: the next band's CTAs fill the SMs while this one drains
        cudaStream_t st = two_streams ? ctx->band_stream[k & 1] : ctx->stream;
        if (pair) CU(cudaStreamWaitEvent(st, ctx->band_event[k], 0));
        const int lo = band_start[(size_t)k], hi = band_start[(size_t)k + 1], mid = band_mid[(size_t)k];
        if (hi <= lo) continue;
        if (mid > lo) {         // the larger borders (all points when the call is not split into classes)
            rc = launch_pm(ctx, mid - lo, dp, dp + n, dp + 2 * n, dp + 3 * n, dp + 4 * n, (const int *)ctx->order.p + lo,
                           max_border, two_class ? typ_big : typ_border, img_size, n_angles, d_angles, d_tab, rot_order, flags,
                           (double *)ctx->out.p, (int *)ctx->status.p, st, two_streams ? (k & 1) : 0, lo);
            if (rc) { ctx->tail_region_stride = 0; return rc; }
        }
        if (hi > mid) {         // the class of pm_ws_kernel
            ctx->k_ev_hold = mid > lo;
            rc = launch_pm(ctx, hi - mid, dp, dp + n, dp + 2 * n, dp + 3 * n, dp + 4 * n, (const int *)ctx->order.p + mid,
                           small_max, typ_small, img_size, n_angles, d_angles, d_tab, rot_order, flags,
                           (double *)ctx->out.p, (int *)ctx->status.p, st, two_streams ? (k & 1) : 0, mid);
            ctx->k_ev_hold = false;
            if (rc) { ctx->tail_region_stride = 0; return rc; }
        }
    }
    ctx->tail_region_stride = 0;
    if (two_streams) {
        for (int k = 0; k < 2; ++k) {
            CU(cudaEventRecord(ctx->slot_event[1 + k], ctx->band_stream[k]));
            CU(cudaStreamWaitEvent(ctx->stream, ctx->slot_event[1 + k], 0));
        }
    }
    if (pair) ctx->have_pair = true;
    double *hout = (double *)((char *)ctx->pin + pts_bytes + ord_bytes);
    int *hst = (int *)((char *)hout + out_bytes);
    ctx->table_n = n;                                          // the (n, 5) table stays in ctx->out for sid_pm_epilogue_affine
    if (out) CU(cudaMemcpyAsync(hout, ctx->out.p, out_bytes, cudaMemcpyDeviceToHost, ctx->stream));
    if (status) CU(cudaMemcpyAsync(hst, ctx->status.p, st_bytes, cudaMemcpyDeviceToHost, ctx->stream));
    CU(cudaStreamSynchronize(ctx->stream));
    CU(cudaGetLastError());
    if (out) memcpy(out, hout, out_bytes);
    if (status) memcpy(status, hst, st_bytes);
    return SID_OK;
}

}  // namespace

int sid_run(sid_ctx *ctx, int64_t n, const double *c1, const double *r1, const double *c2fg,
            const double *r2fg, const double *border, int img_size, int n_angles, const double *angles,
            const double *angle_tab, int rot_order, unsigned flags, int mtype, double *out, int *status) {
    return run_host(ctx, nullptr, n, c1, r1, c2fg, r2fg, border, img_size, n_angles, angles, angle_tab, rot_order,
                    flags, mtype, out, status);
}

int sid_run_pair(sid_ctx *ctx, const uint8_t *img1, int rows1, int cols1, int64_t pitch1,
                 const uint8_t *img2, int rows2, int cols2, int64_t pitch2,
                 int64_t n, const double *c1, const double *r1, const double *c2fg,
                 const double *r2fg, const double *border, int img_size, int n_angles, const double *angles,
                 const double *angle_tab, int rot_order, unsigned flags, int mtype, double *out, int *status) {
    if (!ctx) return SID_EINVAL;
    HostPair hp{img1, img2, rows1, cols1, rows2, cols2, (long long)pitch1, (long long)pitch2};
    return run_host(ctx, &hp, n, c1, r1, c2fg, r2fg, border, img_size, n_angles, angles, angle_tab, rot_order,
                    flags, mtype, out, status);
}

int sid_run_device(sid_ctx *ctx, int64_t n, const double *d_c1, const double *d_r1, const double *d_c2fg,
                   const double *d_r2fg, const double *d_border, int max_border, int img_size, int n_angles,
                   const double *angles, const double *angle_tab, int rot_order, unsigned flags, int mtype,
                   double *d_out, int *d_status) {
    int rc = check_common(ctx, img_size, n_angles, angle_tab, rot_order, mtype);
    if (rc) return rc;
    if (!ctx->have_pair) return fail(ctx, SID_ENOPAIR, "sid_set_pair has not been called");
    if (n < 0 || (n > 0 && (!d_c1 || !d_r1 || !d_c2fg || !d_r2fg || !d_border || !d_out)) || !angles)
        return fail(ctx, SID_EINVAL, "null point array");
    if (n == 0) return SID_OK;
    CU(cudaSetDevice(ctx->device));
    if (max_border <= 0) {
        std::vector<double> hb((size_t)n);
        CU(cudaMemcpyAsync(hb.data(), d_border, (size_t)n * 8, cudaMemcpyDeviceToHost, ctx->stream));
        CU(cudaStreamSynchronize(ctx->stream));
        max_border = 1;
        for (double b : hb) if (std::isfinite(b) && b >= 0.0 && b < 4096.0) max_border = std::max(max_border, (int)std::ceil(b));
    }
    const double *d_angles, *d_tab;
    if ((rc = upload_angles(ctx, n_angles, angles, angle_tab, &d_angles, &d_tab))) return rc;
    ctx->tail_region_stride = 0;
    // border classes as in sid_run (see run_host): the indices are split on the device, one 8-byte read-back sizes the launches
    const char *ce = getenv("SID_PM_CLASSES");
    const int limit = (ce && ce[0] == '0') || n < 512 ? -1 : ws_border_limit(ctx, img_size, n_angles, flags, max_border);
    if (limit >= 0 && limit < max_border) {
        if ((rc = reserve(ctx, ctx->order, (size_t)n * 2 * sizeof(int) + 64))) return rc;
        int *d_small = (int *)ctx->order.p, *d_big = d_small + n;
        unsigned *d_counts = (unsigned *)(d_big + n);
        CU(cudaMemsetAsync(d_counts, 0, 8, ctx->stream));
        partition_by_border_kernel<<<(unsigned)((n + 255) / 256), 256, 0, ctx->stream>>>(d_border, (int)n, limit, d_small, d_big, d_counts);
        ctx->launches += 1;
        unsigned counts[2] = {0, 0};
        CU(cudaMemcpyAsync(counts, d_counts, 8, cudaMemcpyDeviceToHost, ctx->stream));
        CU(cudaStreamSynchronize(ctx->stream));
        CU(cudaGetLastError());
        const long long n_small = counts[0], n_big = counts[1];
        if (n_small >= 256 && n_big > 0) {
            auto rr_of = [&](int b) { const int R = 2 * (img_size / 2) + 2 * b + 1 - img_size + 1; return R * R; };
            const bool smth = (flags & SID_HES_SMTH) != 0;
            ctx->tail_region_stride = (size_t)(((pm_tail_smem_bytes(rr_of(max_border), smth) <= tail_split_limit() ? rr_of(max_border) : rr_of(limit)) + 3) & ~3);
            ctx->tail_hint_n = n;
            rc = launch_pm(ctx, n_big, d_c1, d_r1, d_c2fg, d_r2fg, d_border, d_big, max_border, max_border, img_size,
                           n_angles, d_angles, d_tab, rot_order, flags, d_out, d_status, nullptr, 0, 0);
            if (!rc) {
                ctx->k_ev_hold = true;
                rc = launch_pm(ctx, n_small, d_c1, d_r1, d_c2fg, d_r2fg, d_border, d_small, limit, limit, img_size,
                               n_angles, d_angles, d_tab, rot_order, flags, d_out, d_status, nullptr, 0, n_big);
                ctx->k_ev_hold = false;
            }
            ctx->tail_region_stride = 0;
            ctx->tail_hint_n = 0;
            return rc;
        }
        if (n_big == 0) max_border = limit;              // every point fits the pipeline kernel
    }
    return launch_pm(ctx, n, d_c1, d_r1, d_c2fg, d_r2fg, d_border, nullptr, max_border, max_border, img_size,
                     n_angles, d_angles, d_tab, rot_order, flags, d_out, d_status);
}

// ------------------------------------------------------------------ post-processing epilogue
namespace {
__device__ __forceinline__ double affine_eval(const double *m, double c, double r) {      // (m0*c + m1*r) + m2, NumPy's order, no FMA
    return __dadd_rn(__dadd_rn(__dmul_rn(m[0], c), __dmul_rn(m[1], r)), m[2]);
}
__global__ void pm_epilogue_fill_kernel(double *out, long long n) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) out[i] = nan("");
}
struct EpiArgs { double xy[6]; double ll[6]; };
// reference pmlib.py:462-497: sub-pixel remainder, pixel -> x/y and lon/lat, u = x2 - x1, v = y2 - y1, _fill_gpi scatter
__global__ void pm_epilogue_kernel(long long n_valid, const int *gidx, const double *res, const double *c2pm1, const double *r2pm1,
                                   const EpiArgs e, long long n_grid, double *out) {
    const long long k = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= n_valid) return;
    const long long g = gidx[k];
    const double c = c2pm1[g], r = r2pm1[g];
    const double c2 = __dadd_rn(res[5 * k], __dsub_rn(c, rint(c)));          // results[:, 0] + (c2pm1 - round(c2pm1))[gpi]
    const double r2 = __dadd_rn(res[5 * k + 1], __dsub_rn(r, rint(r)));
    out[g] = __dsub_rn(affine_eval(e.xy, c2, r2), affine_eval(e.xy, c, r));
    out[n_grid + g] = __dsub_rn(affine_eval(e.xy + 3, c2, r2), affine_eval(e.xy + 3, c, r));
    out[2 * n_grid + g] = res[5 * k + 2];
    out[3 * n_grid + g] = res[5 * k + 3];
    out[4 * n_grid + g] = res[5 * k + 4];
    out[5 * n_grid + g] = affine_eval(e.ll, c2, r2);
    out[6 * n_grid + g] = affine_eval(e.ll + 3, c2, r2);
}
}  // namespace

int sid_pm_epilogue_affine(sid_ctx *ctx, int64_t n_valid, const int32_t *grid_index, int64_t n_grid, const double *c2pm1,
                           const double *r2pm1, const double *results, const double *xy, const double *ll, double *out) {
    if (!ctx) return SID_EINVAL;
    if (n_valid < 0 || n_grid <= 0 || n_valid > n_grid || (n_valid > 0 && !grid_index) || !c2pm1 || !r2pm1 || !xy || !ll || !out)
        return fail(ctx, SID_EINVAL, "bad epilogue arguments");
    if (!results && ctx->table_n != n_valid)
        return fail(ctx, SID_EINVAL, "no device result table of that length (run sid_run / sid_run_pair first, or pass results)");
    CU(cudaSetDevice(ctx->device));
    const size_t b_idx = ((size_t)n_valid * 4 + 255) & ~(size_t)255, b_grid = ((size_t)n_grid * 8 + 255) & ~(size_t)255;
    const size_t b_res = results ? (((size_t)n_valid * 40 + 255) & ~(size_t)255) : 0, b_out = (size_t)n_grid * 56;
    int rc = reserve(ctx, ctx->epi, b_idx + 2 * b_grid + b_res + b_out);
    if (rc) return rc;
    char *base = (char *)ctx->epi.p;
    int *d_idx = (int *)base;
    double *d_c = (double *)(base + b_idx), *d_r = (double *)(base + b_idx + b_grid);
    double *d_res = results ? (double *)(base + b_idx + 2 * b_grid) : (double *)ctx->out.p;
    double *d_out = (double *)(base + b_idx + 2 * b_grid + b_res);
    if (n_valid) CU(cudaMemcpyAsync(d_idx, grid_index, (size_t)n_valid * 4, cudaMemcpyHostToDevice, ctx->stream));
    CU(cudaMemcpyAsync(d_c, c2pm1, (size_t)n_grid * 8, cudaMemcpyHostToDevice, ctx->stream));
    CU(cudaMemcpyAsync(d_r, r2pm1, (size_t)n_grid * 8, cudaMemcpyHostToDevice, ctx->stream));
    if (results && n_valid) CU(cudaMemcpyAsync(d_res, results, (size_t)n_valid * 40, cudaMemcpyHostToDevice, ctx->stream));
    EpiArgs e;
    memcpy(e.xy, xy, sizeof e.xy);
    memcpy(e.ll, ll, sizeof e.ll);
    pm_epilogue_fill_kernel<<<(unsigned)((7 * n_grid + 255) / 256), 256, 0, ctx->stream>>>(d_out, 7 * n_grid);
    if (n_valid) pm_epilogue_kernel<<<(unsigned)((n_valid + 255) / 256), 256, 0, ctx->stream>>>(n_valid, d_idx, d_res, d_c, d_r, e, n_grid, d_out);
    ctx->launches += n_valid ? 2 : 1;
    CU(cudaGetLastError());
    CU(cudaMemcpyAsync(out, d_out, b_out, cudaMemcpyDeviceToHost, ctx->stream));
    CU(cudaStreamSynchronize(ctx->stream));
    return SID_OK;
}

// ------------------------------------------------------------------ first guess
namespace {
struct HostGrid {
    std::vector<double> x, y;
    std::vector<int> start, index;
    double x0 = 0, y0 = 0, cell = 1;
    int nx = 1, ny = 1;
};
// uniform hash grid, ~2 points per cell, points sorted by cell (counting sort)
void build_grid(const double *px, const double *py, int n, HostGrid &g) {
    double xmin = 1e300, xmax = -1e300, ymin = 1e300, ymax = -1e300;
    for (int i = 0; i < n; ++i) { xmin = std::min(xmin, px[i]); xmax = std::max(xmax, px[i]); ymin = std::min(ymin, py[i]); ymax = std::max(ymax, py[i]); }
    if (n == 0) { xmin = ymin = 0; xmax = ymax = 1; }
    const double w = std::max(xmax - xmin, 1.0), h = std::max(ymax - ymin, 1.0);
    g.cell = std::max(std::sqrt(w * h * 2.0 / std::max(n, 1)), 1.0);
    g.nx = std::min(2048, std::max(1, (int)std::ceil(w / g.cell) + 1));
    g.ny = std::min(2048, std::max(1, (int)std::ceil(h / g.cell) + 1));
    g.cell = std::max(g.cell, std::max(w / (g.nx - 0.5), h / (g.ny - 0.5)));
    g.x0 = xmin; g.y0 = ymin;
    const double inv = 1.0 / g.cell;
    auto cell_of = [&](double v, double v0, int m) { int c = (int)std::floor((v - v0) * inv); return c < 0 ? 0 : (c >= m ? m - 1 : c); };
    std::vector<int> cid((size_t)n);
    g.start.assign((size_t)g.nx * g.ny + 1, 0);
    for (int i = 0; i < n; ++i) { cid[(size_t)i] = cell_of(py[i], g.y0, g.ny) * g.nx + cell_of(px[i], g.x0, g.nx); ++g.start[(size_t)cid[(size_t)i] + 1]; }
    for (size_t c = 0; c + 1 < g.start.size(); ++c) g.start[c + 1] += g.start[c];
    std::vector<int> fill(g.start.begin(), g.start.end() - 1);
    g.x.resize((size_t)n); g.y.resize((size_t)n); g.index.resize((size_t)n);
    for (int i = 0; i < n; ++i) { const int k = fill[(size_t)cid[(size_t)i]]++; g.x[(size_t)k] = px[i]; g.y[(size_t)k] = py[i]; g.index[(size_t)k] = i; }
}
// convex hull (Andrew's monotone chain), counter-clockwise, collinear points dropped
void convex_hull(const double *px, const double *py, int n, std::vector<int> &hull) {
    std::vector<int> idx((size_t)n);
    for (int i = 0; i < n; ++i) idx[(size_t)i] = i;
    std::sort(idx.begin(), idx.end(), [&](int a, int b) { return px[a] < px[b] || (px[a] == px[b] && py[a] < py[b]); });
    auto cross = [&](int o, int a, int b) { return (px[a] - px[o]) * (py[b] - py[o]) - (py[a] - py[o]) * (px[b] - px[o]); };
    hull.assign((size_t)2 * n + 2, 0);
    int k = 0;
    for (int i = 0; i < n; ++i) { while (k >= 2 && cross(hull[(size_t)k - 2], hull[(size_t)k - 1], idx[(size_t)i]) <= 0) --k; hull[(size_t)k++] = idx[(size_t)i]; }
    for (int i = n - 2, t = k + 1; i >= 0; --i) { while (k >= t && cross(hull[(size_t)k - 2], hull[(size_t)k - 1], idx[(size_t)i]) <= 0) --k; hull[(size_t)k++] = idx[(size_t)i]; }
    hull.resize((size_t)std::max(0, k - 1));
}
}  // namespace

int sid_first_guess(sid_ctx *ctx, int n_src, const double *sx, const double *sy, const double *vx, const double *vy,
                    int n_seed, const double *kx, const double *ky, int64_t n_q, const double *qx, const double *qy,
                    double *out_vx, double *out_vy, double *out_dist, int32_t *out_flag) {
    if (!ctx) return SID_EINVAL;
    if (n_src < 0 || n_seed < 0 || n_q < 0 || (n_src > 0 && (!sx || !sy || !vx || !vy)) || (n_seed > 0 && (!kx || !ky)) ||
        (n_q > 0 && (!qx || !qy || !out_vx || !out_vy || !out_dist || !out_flag)))
        return fail(ctx, SID_EINVAL, "bad first-guess arguments");
    if (n_q == 0) return SID_OK;
    CU(cudaSetDevice(ctx->device));
    HostGrid gs, gk;
    build_grid(sx, sy, n_src, gs);
    build_grid(kx, ky, n_seed, gk);
    std::vector<int> hull;
    if (n_src >= 3) convex_hull(sx, sy, n_src, hull);
    const int nh = (int)hull.size();
    std::vector<double> hx((size_t)nh), hy((size_t)nh);
    for (int i = 0; i < nh; ++i) { hx[(size_t)i] = sx[hull[(size_t)i]]; hy[(size_t)i] = sy[hull[(size_t)i]]; }
    // one staging block: everything the kernel reads, 256-byte aligned pieces
    struct Piece { const void *src; size_t bytes; size_t off; };
    std::vector<Piece> pieces;
    size_t total = 0;
    auto add = [&](const void *p, size_t bytes) { pieces.push_back({p, bytes, total}); total += (bytes + 255) & ~(size_t)255; return pieces.size() - 1; };
    const size_t i_sx = add(gs.x.data(), (size_t)n_src * 8), i_sy = add(gs.y.data(), (size_t)n_src * 8);
    const size_t i_ss = add(gs.start.data(), gs.start.size() * 4), i_si = add(gs.index.data(), (size_t)n_src * 4);
    const size_t i_vx = add(vx, (size_t)n_src * 8), i_vy = add(vy, (size_t)n_src * 8);
    const size_t i_hx = add(hx.data(), (size_t)nh * 8), i_hy = add(hy.data(), (size_t)nh * 8), i_hi = add(hull.data(), (size_t)nh * 4);
    const size_t i_kx = add(gk.x.data(), (size_t)n_seed * 8), i_ky = add(gk.y.data(), (size_t)n_seed * 8);
    const size_t i_ks = add(gk.start.data(), gk.start.size() * 4), i_ki = add(gk.index.data(), (size_t)n_seed * 4);
    const size_t i_qx = add(qx, (size_t)n_q * 8), i_qy = add(qy, (size_t)n_q * 8);
    const size_t o_vx = total; total += ((size_t)n_q * 8 + 255) & ~(size_t)255;
    const size_t o_vy = total; total += ((size_t)n_q * 8 + 255) & ~(size_t)255;
    const size_t o_d = total; total += ((size_t)n_q * 8 + 255) & ~(size_t)255;
    const size_t o_f = total; total += ((size_t)n_q * 4 + 255) & ~(size_t)255;
    int rc = reserve(ctx, ctx->fg, total);
    if (rc) return rc;
    char *base = (char *)ctx->fg.p;
    for (const Piece &pc : pieces)
        if (pc.bytes) CU(cudaMemcpyAsync(base + pc.off, pc.src, pc.bytes, cudaMemcpyHostToDevice, ctx->stream));
    FgArgs a;
    memset(&a, 0, sizeof a);
    auto at = [&](size_t i) { return base + pieces[i].off; };
    a.src.x = (const double *)at(i_sx); a.src.y = (const double *)at(i_sy); a.src.start = (const int *)at(i_ss); a.src.index = (const int *)at(i_si);
    a.src.x0 = gs.x0; a.src.y0 = gs.y0; a.src.cell = gs.cell; a.src.inv_cell = 1.0 / gs.cell; a.src.nx = gs.nx; a.src.ny = gs.ny; a.src.n = n_src;
    a.vx = (const double *)at(i_vx); a.vy = (const double *)at(i_vy);
    a.hx = (const double *)at(i_hx); a.hy = (const double *)at(i_hy); a.hidx = (const int *)at(i_hi); a.nh = nh;
    a.seed.x = (const double *)at(i_kx); a.seed.y = (const double *)at(i_ky); a.seed.start = (const int *)at(i_ks); a.seed.index = (const int *)at(i_ki);
    a.seed.x0 = gk.x0; a.seed.y0 = gk.y0; a.seed.cell = gk.cell; a.seed.inv_cell = 1.0 / gk.cell; a.seed.nx = gk.nx; a.seed.ny = gk.ny; a.seed.n = n_seed;
    a.nq = n_q; a.qx = (const double *)at(i_qx); a.qy = (const double *)at(i_qy);
    a.out_vx = (double *)(base + o_vx); a.out_vy = (double *)(base + o_vy); a.out_dist = (double *)(base + o_d); a.out_flag = (int *)(base + o_f);
    first_guess_kernel<<<(unsigned)((n_q + 127) / 128), 128, 0, ctx->stream>>>(a);
    ctx->launches += 1;
    CU(cudaGetLastError());
    CU(cudaMemcpyAsync(out_vx, base + o_vx, (size_t)n_q * 8, cudaMemcpyDeviceToHost, ctx->stream));
    CU(cudaMemcpyAsync(out_vy, base + o_vy, (size_t)n_q * 8, cudaMemcpyDeviceToHost, ctx->stream));
    CU(cudaMemcpyAsync(out_dist, base + o_d, (size_t)n_q * 8, cudaMemcpyDeviceToHost, ctx->stream));
    CU(cudaMemcpyAsync(out_flag, base + o_f, (size_t)n_q * 4, cudaMemcpyDeviceToHost, ctx->stream));
    CU(cudaStreamSynchronize(ctx->stream));
    return SID_OK;
}

// ------------------------------------------------------------------ single-call entry points

int sid_get_template(sid_ctx *ctx, const uint8_t *img, int rows, int cols, int64_t pitch, double c, double r,
                     const double *angle_tab, int s, int rot_order, uint8_t *out) {
    if (!ctx) return SID_EINVAL;
    if (!img || !out || !angle_tab || s <= 0 || s > 4096) return fail(ctx, SID_EINVAL, "bad get_template arguments");
    if (rot_order != 0 && rot_order != 1) return fail(ctx, SID_EUNSUPPORTED, "rot_order must be 0 or 1");
    CU(cudaSetDevice(ctx->device));
    long long dp;
    int rc = upload_image(ctx, ctx->img1, dp, img, rows, cols, pitch, cudaMemcpyHostToDevice);
    ctx->have_pair = false;
    if (rc) return rc;
    const size_t tb = (size_t)s * s;
    if ((rc = reserve(ctx, ctx->misc, tb + 64 + 4 * 8 + 16))) return rc;
    uint8_t *d_t = (uint8_t *)ctx->misc.p;
    const size_t off_tab = (tb + 63) / 64 * 64;
    double *d_tab = (double *)(d_t + off_tab);
    uint32_t *d_stats = (uint32_t *)(d_tab + 4);
    CU(cudaMemcpyAsync(d_tab, angle_tab, 32, cudaMemcpyHostToDevice, ctx->stream));
    CU(cudaMemsetAsync(d_stats, 0, 12, ctx->stream));
    templates_kernel<<<1, 256, 0, ctx->stream>>>((const uint8_t *)ctx->img1.p, rows, cols, dp, c, r, d_tab, s, rot_order, d_t, d_stats);
    ctx->launches += 1;
    CU(cudaGetLastError());
    CU(cudaMemcpyAsync(out, d_t, tb, cudaMemcpyDeviceToHost, ctx->stream));
    CU(cudaStreamSynchronize(ctx->stream));
    return SID_OK;
}

// NCC map of a device image/template pair into d_out; isum/isq/tstats are device scratch.
static int match_on_device(sid_ctx *ctx, const uint8_t *d_img, int H, int W, long long pitch, const uint8_t *d_tpl,
                           int th, int tw, long long tp, uint32_t *d_isum, uint32_t *d_isq, bool integrals_ready,
                           uint32_t *d_tstats, float *d_out) {
    const int RH = H - th + 1, RW = W - tw + 1;
    if (!integrals_ready) {
        integral_rows_kernel<<<(H + 1 + 127) / 128, 128, 0, ctx->stream>>>(d_img, H, W, pitch, d_isum, d_isq);
        integral_cols_kernel<<<(W + 1 + 127) / 128, 128, 0, ctx->stream>>>(H, W, d_isum, d_isq);
        ctx->launches += 2;
    }
    CU(cudaMemsetAsync(d_tstats, 0, 8, ctx->stream));
    template_sums_kernel<<<std::min(64, (th * tw + 255) / 256), 256, 0, ctx->stream>>>(d_tpl, th, tw, tp, d_tstats);
    MtArgs a;
    a.img = d_img; a.H = H; a.W = W; a.pitch = pitch;
    a.tpl = d_tpl; a.th = th; a.tw = tw; a.tp = tp;
    a.isum = d_isum; a.isq = d_isq; a.tstats = d_tstats;
    const double area = (double)th * (double)tw;
    a.inv_area = 1.0 / area; a.sqrt_inv_area = std::sqrt(a.inv_area);
    a.out = d_out;
    a.tpw = (tw + 15) / 16 * 4;
    int n16 = (MT_COLS + tw + 3 + 15) / 16 + 1;
    if ((n16 & 1) == 0) ++n16;
    a.wpw = n16 * 4;
    const size_t smem = ((size_t)(MT_ROWS + th - 1) * a.wpw + PM_WIN_SLACK + (size_t)th * a.tpw) * 4;
    if (smem > (size_t)ctx->max_smem_optin) return fail(ctx, SID_EUNSUPPORTED, "template too large for match_template");
    if (int arc = allow_max_smem(ctx, (const void *)match_template_kernel)) return arc;
    dim3 grid((RW + MT_COLS - 1) / MT_COLS, (RH + MT_ROWS - 1) / MT_ROWS);
    match_template_kernel<<<grid, 256, smem, ctx->stream>>>(a);
    ctx->launches += 2;
    CU(cudaGetLastError());
    return SID_OK;
}

int sid_match_template(sid_ctx *ctx, const uint8_t *img, int H, int W, int64_t pitch, const uint8_t *tpl,
                       int th, int tw, int64_t tpitch, int method, float *out) {
    if (!ctx) return SID_EINVAL;
    if (method != SID_TM_CCOEFF_NORMED) return fail(ctx, SID_EUNSUPPORTED, "only TM_CCOEFF_NORMED (5) is implemented");
    if (!img || !tpl || !out || th <= 0 || tw <= 0 || H < th || W < tw || pitch < W || tpitch < tw)
        return fail(ctx, SID_EINVAL, "bad match_template arguments (image must not be smaller than the template)");
    if ((long long)th * tw > 32768) return fail(ctx, SID_EUNSUPPORTED, "template area above 32768 pixels");
    CU(cudaSetDevice(ctx->device));
    long long dp;
    int rc = upload_image(ctx, ctx->img2, dp, img, H, W, pitch, cudaMemcpyHostToDevice);
    ctx->have_pair = false;
    if (rc) return rc;
    const int RH = H - th + 1, RW = W - tw + 1;
    const size_t int_bytes = (size_t)(H + 1) * (W + 1) * 4;
    const size_t tpl_bytes = ((size_t)th * tw + 255) / 256 * 256;
    const size_t out_bytes = (size_t)RH * RW * 4;
    if ((rc = reserve(ctx, ctx->misc, 2 * int_bytes + tpl_bytes + out_bytes + 1024))) return rc;
    unsigned char *base = (unsigned char *)ctx->misc.p;
    uint32_t *d_isum = (uint32_t *)base, *d_isq = (uint32_t *)(base + int_bytes);
    uint8_t *d_tpl = base + 2 * int_bytes;
    float *d_out = (float *)(base + 2 * int_bytes + tpl_bytes);
    uint32_t *d_ts = (uint32_t *)(base + 2 * int_bytes + tpl_bytes + out_bytes);
    d_ts = (uint32_t *)(((uintptr_t)d_ts + 15) & ~(uintptr_t)15);
    CU(cudaMemcpy2DAsync(d_tpl, (size_t)tw, tpl, (size_t)tpitch, (size_t)tw, (size_t)th, cudaMemcpyHostToDevice, ctx->stream));
    rc = match_on_device(ctx, (const uint8_t *)ctx->img2.p, H, W, dp, d_tpl, th, tw, tw, d_isum, d_isq, false, d_ts, d_out);
    if (rc) return rc;
    CU(cudaMemcpyAsync(out, d_out, out_bytes, cudaMemcpyDeviceToHost, ctx->stream));
    CU(cudaStreamSynchronize(ctx->stream));
    CU(cudaGetLastError());
    return SID_OK;
}

int sid_knn_hamming2(sid_ctx *ctx, const uint8_t *d1, int n1, const uint8_t *d2, int n2, int desc_bytes,
                     int32_t *idx, int32_t *dist) {
    if (!ctx) return SID_EINVAL;
    if (desc_bytes != 32) return fail(ctx, SID_EUNSUPPORTED, "only 32-byte (ORB) descriptors are implemented");
    if (n1 < 0 || n2 < 0 || (n1 > 0 && (!d1 || !idx || !dist)) || (n2 > 0 && !d2)) return fail(ctx, SID_EINVAL, "bad knn arguments");
    if (n1 == 0) return SID_OK;
    if (n2 == 0) { for (int k = 0; k < 2 * n1; ++k) { idx[k] = -1; dist[k] = -1; } return SID_OK; }
    CU(cudaSetDevice(ctx->device));
    const int qblocks = (n1 + KNN_THREADS - 1) / KNN_THREADS;
    int nseg = (ctx->sm_count * 8 + qblocks - 1) / qblocks;               // >= 8 CTAs per SM in flight
    nseg = std::max(1, std::min(nseg, std::min(64, (n2 + KNN_TILE - 1) / KNN_TILE)));
    int seg_len = ((n2 + nseg - 1) / nseg + KNN_TILE - 1) / KNN_TILE * KNN_TILE;
    nseg = (n2 + seg_len - 1) / seg_len;
    const size_t b1 = (size_t)n1 * 32, b2 = (size_t)n2 * 32, bp = (size_t)nseg * n1 * sizeof(Top2), bo = (size_t)n1 * 2 * 4;
    const size_t o2 = (b1 + 255) / 256 * 256, op = o2 + (b2 + 255) / 256 * 256, oi = op + (bp + 255) / 256 * 256, od = oi + (bo + 255) / 256 * 256;
    int rc = reserve(ctx, ctx->misc, od + bo + 256);
    if (rc) return rc;
    unsigned char *base = (unsigned char *)ctx->misc.p;
    CU(cudaMemcpyAsync(base, d1, b1, cudaMemcpyHostToDevice, ctx->stream));
    CU(cudaMemcpyAsync(base + o2, d2, b2, cudaMemcpyHostToDevice, ctx->stream));
    knn_hamming2_kernel<<<dim3((unsigned)qblocks, (unsigned)nseg), KNN_THREADS, 0, ctx->stream>>>(
        (const uint4 *)base, n1, (const uint4 *)(base + o2), n2, seg_len, (Top2 *)(base + op));
    knn_merge_kernel<<<(n1 + 255) / 256, 256, 0, ctx->stream>>>((const Top2 *)(base + op), n1, nseg, (int *)(base + oi), (int *)(base + od));
    ctx->launches += 2;
    CU(cudaGetLastError());
    CU(cudaMemcpyAsync(idx, base + oi, bo, cudaMemcpyDeviceToHost, ctx->stream));
    CU(cudaMemcpyAsync(dist, base + od, bo, cudaMemcpyDeviceToHost, ctx->stream));
    CU(cudaStreamSynchronize(ctx->stream));
    CU(cudaGetLastError());
    return SID_OK;
}

int sid_deformation(sid_ctx *ctx, int n, const double *x, const double *y, const double *u, const double *v,
                    int m, const int32_t *tri, const double *area_in,
                    double *e1, double *e2, double *e3, double *area, double *perim) {
    if (!ctx) return SID_EINVAL;
    if (n < 0 || m < 0 || (n > 0 && (!x || !y || !u || !v)) || (m > 0 && (!tri || !e1 || !e2 || !e3 || !area || !perim)))
        return fail(ctx, SID_EINVAL, "bad deformation arguments");
    if (m == 0) return SID_OK;
    CU(cudaSetDevice(ctx->device));
    const size_t bn = ((size_t)n * 8 + 255) / 256 * 256, bm = ((size_t)m * 8 + 255) / 256 * 256;
    const size_t bt = ((size_t)m * 12 + 255) / 256 * 256;
    int rc = reserve(ctx, ctx->misc, 4 * bn + bt + 6 * bm + 256);
    if (rc) return rc;
    unsigned char *base = (unsigned char *)ctx->misc.p;
    DeforArgs a;
    const double *src[4] = {x, y, u, v};
    double *dn[4];
    for (int k = 0; k < 4; ++k) {
        dn[k] = (double *)(base + k * bn);
        if (n > 0) CU(cudaMemcpyAsync(dn[k], src[k], (size_t)n * 8, cudaMemcpyHostToDevice, ctx->stream));
    }
    a.x = dn[0]; a.y = dn[1]; a.u = dn[2]; a.v = dn[3];
    int32_t *dt = (int32_t *)(base + 4 * bn);
    CU(cudaMemcpyAsync(dt, tri, (size_t)m * 12, cudaMemcpyHostToDevice, ctx->stream));
    a.tri = dt;
    double *dm = (double *)(base + 4 * bn + bt);
    a.e1 = dm; a.e2 = (double *)((char *)dm + bm); a.e3 = (double *)((char *)dm + 2 * bm);
    a.area = (double *)((char *)dm + 3 * bm); a.perim = (double *)((char *)dm + 4 * bm);
    double *dain = (double *)((char *)dm + 5 * bm);
    a.area_in = nullptr;
    if (area_in) {
        CU(cudaMemcpyAsync(dain, area_in, (size_t)m * 8, cudaMemcpyHostToDevice, ctx->stream));
        a.area_in = dain;
    }
    a.n = n; a.m = m;
    deformation_kernel<<<(m + 255) / 256, 256, 0, ctx->stream>>>(a);
    ctx->launches += 1;
    CU(cudaGetLastError());
    double *dst[5] = {e1, e2, e3, area, perim};
    for (int k = 0; k < 5; ++k)
        CU(cudaMemcpyAsync(dst[k], (char *)dm + k * bm, (size_t)m * 8, cudaMemcpyDeviceToHost, ctx->stream));
    CU(cudaStreamSynchronize(ctx->stream));
    CU(cudaGetLastError());
    return SID_OK;
}

int sid_get_hessian(sid_ctx *ctx, const float *ccm, int rows, int cols, unsigned flags, float *out) {
    if (!ctx) return SID_EINVAL;
    if (!ccm || !out || rows < 2 || cols < 2) return fail(ctx, SID_EINVAL, "get_hessian needs a map of at least 2 x 2");
    CU(cudaSetDevice(ctx->device));
    const size_t n = (size_t)rows * cols, b = n * 4;
    int rc = reserve(ctx, ctx->misc, 4 * b + 64);
    if (rc) return rc;
    float *d = (float *)ctx->misc.p;
    CU(cudaMemcpyAsync(d, ccm, b, cudaMemcpyHostToDevice, ctx->stream));
    HesArgs a;
    a.ccm = d; a.rows = rows; a.cols = cols; a.flags = flags; gaussian_weights(a.gw);
    a.tmp_a = d + n; a.tmp_b = d + 2 * n; a.out = d + 3 * n;
    hessian_map_kernel<<<1, 1024, 0, ctx->stream>>>(a);
    ctx->launches += 1;
    CU(cudaGetLastError());
    CU(cudaMemcpyAsync(out, a.out, b, cudaMemcpyDeviceToHost, ctx->stream));
    CU(cudaStreamSynchronize(ctx->stream));
    CU(cudaGetLastError());
    return SID_OK;
}

int sid_rotate_and_match(sid_ctx *ctx, const uint8_t *img1, int rows1, int cols1, int64_t pitch1, double c1,
                         double r1, int img_size, const uint8_t *image2, int H, int W, int64_t pitch2,
                         int n_angles, const double *angle_tab, int rot_order, unsigned flags, int mtype,
                         int *valid, double *dc, double *dr, int *best_angle_idx, float *best_r, float *best_h,
                         float *best_result, uint8_t *best_template) {
    int rc = check_common(ctx, img_size, n_angles, angle_tab, rot_order, mtype);
    if (rc) return rc;
    if (!valid || !dc || !dr || !best_angle_idx || !best_r || !best_h) return fail(ctx, SID_EINVAL, "null output pointer");
    const int s = img_size;
    if (!img1 || !image2 || H < s + 1 || W < s + 1)
        return fail(ctx, SID_EINVAL, "search window must exceed the template by at least one pixel in each axis");
    CU(cudaSetDevice(ctx->device));
    long long dp1, dp2;
    ctx->have_pair = false;
    if ((rc = upload_image(ctx, ctx->img1, dp1, img1, rows1, cols1, pitch1, cudaMemcpyHostToDevice))) return rc;
    if ((rc = upload_image(ctx, ctx->img2, dp2, image2, H, W, pitch2, cudaMemcpyHostToDevice))) return rc;
    const int RH = H - s + 1, RW = W - s + 1;
    const size_t rr = (size_t)RH * RW;
    const size_t int_bytes = ((size_t)(H + 1) * (W + 1) * 4 + 255) / 256 * 256;
    const size_t tpl_bytes = ((size_t)n_angles * s * s + 255) / 256 * 256;
    const size_t map_bytes = (rr * 4 + 255) / 256 * 256;
    const size_t small = 4096 + (size_t)n_angles * 64;
    if ((rc = reserve(ctx, ctx->misc, 2 * int_bytes + tpl_bytes + ((size_t)n_angles + 3) * map_bytes + small))) return rc;
    unsigned char *base = (unsigned char *)ctx->misc.p;
    uint32_t *d_isum = (uint32_t *)base, *d_isq = (uint32_t *)(base + int_bytes);
    uint8_t *d_tpl = base + 2 * int_bytes;
    float *d_maps = (float *)(base + 2 * int_bytes + tpl_bytes);     // one map per angle, then tmp_a, tmp_b, hes
    unsigned char *sm = base + 2 * int_bytes + tpl_bytes + ((size_t)n_angles + 3) * map_bytes;
    double *d_tab = (double *)sm;                                     // n_angles * 4 doubles
    uint32_t *d_stats = (uint32_t *)(sm + (size_t)n_angles * 32);     // n_angles * 3
    uint32_t *d_ts = d_stats + (size_t)n_angles * 3 + 4;              // 2
    unsigned long long *d_key = (unsigned long long *)(((uintptr_t)(d_ts + 4) + 15) & ~(uintptr_t)15);   // one per angle
    float *d_peak = (float *)(d_key + n_angles + 1);
    const size_t mstride = map_bytes / 4;

    CU(cudaMemcpyAsync(d_tab, angle_tab, (size_t)n_angles * 32, cudaMemcpyHostToDevice, ctx->stream));
    CU(cudaMemsetAsync(d_stats, 0, (size_t)n_angles * 12, ctx->stream));
    templates_kernel<<<n_angles, 256, 0, ctx->stream>>>((const uint8_t *)ctx->img1.p, rows1, cols1, dp1, c1, r1, d_tab, s,
                                                       rot_order, d_tpl, d_stats);
    ctx->launches += 1;
    std::vector<uint32_t> hstats((size_t)n_angles * 3);
    CU(cudaMemcpyAsync(hstats.data(), d_stats, (size_t)n_angles * 12, cudaMemcpyDeviceToHost, ctx->stream));
    CU(cudaStreamSynchronize(ctx->stream));
    CU(cudaGetLastError());
    *valid = 1;
    for (int a = 0; a < n_angles; ++a) if (hstats[(size_t)a * 3 + 2]) *valid = 0;
    if (!*valid) {
        *dc = NAN; *dr = NAN; *best_angle_idx = -1; *best_r = NAN; *best_h = NAN;
        return SID_OK;
    }
    // every angle's map and arg-max are enqueued back to back; ONE read-back of the keys decides the winner (strict '>' in
    // angle order, like the reference loop pmlib.py:155-166)
    for (int a = 0; a < n_angles; ++a) {
        float *d_map = d_maps + (size_t)a * mstride;
        rc = match_on_device(ctx, (const uint8_t *)ctx->img2.p, H, W, dp2, d_tpl + (size_t)a * s * s, s, s, s,
                             d_isum, d_isq, a > 0, d_ts, d_map);
        if (rc) return rc;
        argmax_kernel<<<1, 1024, 0, ctx->stream>>>(d_map, (int)rr, d_key + a);
        ctx->launches += 1;
    }
    std::vector<unsigned long long> keys((size_t)n_angles, 0ull);
    CU(cudaMemcpyAsync(keys.data(), d_key, (size_t)n_angles * 8, cudaMemcpyDeviceToHost, ctx->stream));
    CU(cudaStreamSynchronize(ctx->stream));
    CU(cudaGetLastError());
    float br = -INFINITY; int ba = -1; uint32_t bidx = 0;
    for (int a = 0; a < n_angles; ++a) {
        const unsigned long long key = keys[(size_t)a];
        const uint32_t hi = (uint32_t)(key >> 32);
        const uint32_t ubits = (hi & 0x80000000u) ? (hi & 0x7fffffffu) : ~hi;
        float v; memcpy(&v, &ubits, 4);
        if (v > br) { br = v; ba = a; bidx = 0xffffffffu - (uint32_t)(key & 0xffffffffull); }
    }
    if (ba < 0) {
        *valid = 0; *dc = NAN; *dr = NAN; *best_angle_idx = -1; *best_r = NAN; *best_h = NAN;
        return SID_OK;
    }
    PeakArgs pa;
    pa.best = d_maps + (size_t)ba * mstride; pa.rows = RH; pa.cols = RW; pa.idx = (int)bidx; pa.r = br; pa.flags = flags;
    gaussian_weights(pa.gw);
    pa.tmp_a = d_maps + (size_t)n_angles * mstride; pa.tmp_b = pa.tmp_a + mstride; pa.hes = pa.tmp_b + mstride; pa.out = d_peak;
    peak_stats_kernel<<<1, 1024, 0, ctx->stream>>>(pa);
    ctx->launches += 1;
    float hp[2];
    CU(cudaMemcpyAsync(hp, d_peak, 8, cudaMemcpyDeviceToHost, ctx->stream));
    if (best_result) CU(cudaMemcpyAsync(best_result, pa.best, rr * 4, cudaMemcpyDeviceToHost, ctx->stream));
    if (best_template) CU(cudaMemcpyAsync(best_template, d_tpl + (size_t)ba * s * s, (size_t)s * s, cudaMemcpyDeviceToHost, ctx->stream));
    CU(cudaStreamSynchronize(ctx->stream));
    CU(cudaGetLastError());
    const int bi = (int)(bidx / (uint32_t)RW), bj = (int)(bidx % (uint32_t)RW);
    *dr = (double)bi - (double)(H - s) / 2.0;
    *dc = (double)bj - (double)(W - s) / 2.0;
    *best_angle_idx = ba;
    *best_r = hp[0];
    *best_h = hp[1];
    return SID_OK;
}

}  // extern "C"
