/*
 * sid_b200.h -- C ABI of the B200-native pattern-matching (MCC) hot path.
 *
 * Drop-in boundary for nansencenter/sea_ice_drift v0.7.1 (paths below are under
 * the reference tree, sea_ice_drift/):
 *
 *   sid_run / sid_run_pair / sid_run_device
 *                              replace the per-point loop of pattern_matching,
 *                              pmlib.py:430-448 (_init_pool + Pool.map(use_mcc_mp)),
 *                              i.e. use_mcc (pmlib.py:176-212) for every grid point.
 *                              Argument order mirrors _init_pool's tuple (pmlib.py:438).
 *   sid_rotate_and_match       replaces rotate_and_match, pmlib.py:117-174.
 *   sid_get_template           replaces get_template, pmlib.py:89-115.
 *   sid_match_template         replaces the template_matcher plug-in call
 *                              cv2.matchTemplate(image, templ, TM_CCOEFF_NORMED),
 *                              pmlib.py:120 (default) / pmlib.py:156 (call site).
 *   sid_get_hessian            replaces get_hessian, pmlib.py:36-59.
 *
 * Conventions
 *   - plain pointers and sizes only; every function returns 0 or a negative
 *     SID_E* code and never throws/aborts; sid_last_error() gives the message.
 *   - images are uint8, row-major, `pitch` bytes between rows; pixel value 0 means
 *     "invalid" exactly as in the reference (lib.py:52-57).
 *   - `angle_tab` holds 4 doubles per angle: cos a, sin a, tcdot0, tcdot1 with
 *     a = radians(angle - alpha0), tc = int(img_size/2.)+1 and
 *     tcdot = [tc,tc] . [[cos a,-sin a],[sin a,cos a]]  (pmlib.py:105-110); the
 *     caller evaluates it with NumPy so results do not depend on a device libm.
 *   - a context is bound to one device and one stream; it is not re-entrant.
 *   - invalid POINTS are not errors: they yield NaN rows, exactly where the
 *     reference returns NaN (a 0 pixel in any angle's template, pmlib.py:152-154).
 *     Points whose search window starts outside image 2 (the reference would wrap
 *     or fail inside cv2) also yield NaN rows and status -1.
 */
#ifndef SID_B200_H
#define SID_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct sid_ctx sid_ctx;

#define SID_OK 0
#define SID_EINVAL (-1)      /* bad argument                                  */
#define SID_ECUDA (-2)       /* CUDA runtime error (see sid_last_error)       */
#define SID_ENOMEM (-3)      /* host or device allocation failed              */
#define SID_EUNSUPPORTED (-4)/* e.g. mtype != TM_CCOEFF_NORMED, rot_order > 1 */
#define SID_ENOPAIR (-5)     /* sid_set_pair has not been called              */

#define SID_TM_CCOEFF_NORMED 5 /* == cv2.TM_CCOEFF_NORMED */

/* flags for sid_run* / sid_rotate_and_match */
#define SID_HES_NORM 1u      /* get_hessian(hes_norm=True)  -- reference default */
#define SID_HES_SMTH 2u      /* get_hessian(hes_smth=True)                       */
#define SID_MCC_NORM 4u      /* rotate_and_match(mcc_norm=True)                  */

/* point status written by sid_run* (optional output) */
#define SID_PT_VALID 1
#define SID_PT_NAN 0         /* zero pixel in a template -> NaN row              */
#define SID_PT_REJECTED (-1) /* window not inside image 2 / too small            */

const char *sid_version(void);

/* Create / destroy a context on CUDA device `device`. */
int sid_create(sid_ctx **out, int device);
void sid_destroy(sid_ctx *ctx);
const char *sid_last_error(const sid_ctx *ctx);

/* Run all work of this context on `cuda_stream` (a cudaStream_t passed as void*;
 * NULL = the context's own stream, (void*)1 = cudaStreamLegacy, the default stream). */
int sid_set_stream(sid_ctx *ctx, void *cuda_stream);
int sid_synchronize(sid_ctx *ctx);

/* Upload the image pair (host pointers); it stays resident on the device until
 * the next call (time-series reuse).  Replaces the images in _init_pool's tuple. */
int sid_set_pair(sid_ctx *ctx,
                 const uint8_t *img1, int rows1, int cols1, int64_t pitch1,
                 const uint8_t *img2, int rows2, int cols2, int64_t pitch2);
/* Same, but the sources are DEVICE pointers (device-to-device copy into the
 * context's padded layout), e.g. a peer-broadcast image. */
int sid_set_pair_device(sid_ctx *ctx,
                        const uint8_t *d_img1, int rows1, int cols1, int64_t pitch1,
                        const uint8_t *d_img2, int rows2, int cols2, int64_t pitch2);

/* Layout the library uses for a resident image: row pitch (a legal TMA stride: multiple of 16, >= cols + 16) and the
 * total allocation (pitch * rows + tail slack).  For callers that fill the pair themselves (sid_adopt_pair_device). */
int sid_pair_layout(int rows, int cols, int64_t *pitch, int64_t *bytes);
/* Use CALLER-OWNED device buffers as the resident pair, without a copy: the multi-GPU path, where every rank
 * uploads one row slab of each image and an in-place NCCL all-gather over NVLink completes the buffers (replaces
 * the fork/copy-on-write replication of the reference's Pool, pmlib.py:430-448).  The buffers must follow
 * sid_pair_layout (256-byte aligned base, pitch, bytes) and outlive their use; the library never frees them. */
int sid_adopt_pair_device(sid_ctx *ctx,
                          uint8_t *d_img1, int rows1, int cols1, int64_t pitch1, int64_t bytes1,
                          uint8_t *d_img2, int rows2, int cols2, int64_t pitch2, int64_t bytes2);

/* First guess of pattern_matching on the device (reference pmlib.py:249-324 prepare_first_guess, pmlib.py:61-77
 * get_distance_to_nearest_keypoint; lib.py:179-201 interpolation_near): for every query point q
 *   out_vx/out_vy  the Delaunay-linear interpolant of (vx, vy) given at the sources (sx, sy) -- what
 *                  scipy.interpolate.griddata(method='linear') returns -- NaN outside the sources' convex hull;
 *   out_dist       the distance to the nearest (kx, ky) point (the reference samples a full-image distance transform);
 *   out_flag       0 ok, 1 outside the hull, 2 not resolved numerically (caller falls back to its host path).
 * No triangulation is built: each point finds its own Delaunay triangle by pivoting (csrc/sid_fg_kernel.cuh). */
int sid_first_guess(sid_ctx *ctx, int n_src, const double *sx, const double *sy, const double *vx, const double *vy,
                    int n_seed, const double *kx, const double *ky,
                    int64_t n_q, const double *qx, const double *qy,
                    double *out_vx, double *out_vy, double *out_dist, int32_t *out_flag);

/* Post-processing of pattern_matching for AFFINE geolocation (reference pmlib.py:462-497 and lib.py:408-412), on the
 * device: sub-pixel remainder c2 += c2pm1 - round(c2pm1), pixel -> destination x/y and lon/lat, u = x2 - x1,
 * v = y2 - y1, and the _fill_gpi scatter into NaN-filled grids.
 *   n_valid, grid_index  the valid grid points (gpi) in table order and their flat grid positions
 *   c2pm1, r2pm1         n_grid image-2 pixel coordinates of the grid (host)
 *   results              host (n_valid, 5) table, or NULL: use the table the previous sid_run / sid_run_pair left on
 *                        the device (those accept out == NULL, so the table never visits the host)
 *   xy, ll               2 x 3 row-major affine maps pixel -> destination x/y and pixel -> lon/lat: m0*c + m1*r + m2
 *   out                  host, 7 x n_grid doubles: u, v, a, r, h, lon2, lat2                                         */
int sid_pm_epilogue_affine(sid_ctx *ctx, int64_t n_valid, const int32_t *grid_index, int64_t n_grid,
                           const double *c2pm1, const double *r2pm1, const double *results,
                           const double *xy, const double *ll, double *out);

/* Asynchronous 2-D host -> device copy of `rows` image rows on the context's stream (a rank's row slab of the
 * pair, before the all-gather that completes the adopted buffers). */
int sid_upload_rows(sid_ctx *ctx, uint8_t *d_dst, int64_t dst_pitch, const uint8_t *src, int64_t src_pitch, int cols, int rows);

/* use_mcc for n grid points (host arrays in, host array out, synchronous).
 *   c1, r1        float pixel coordinates on image 1
 *   c2fg, r2fg    integer-valued first guess on image 2
 *   border        integer-valued search radius per point
 *   out           n x 5 row-major doubles: c2, r2, angle, r, h  (NaN rows = invalid)
 *   status        optional (may be NULL), n ints, SID_PT_*                          */
int sid_run(sid_ctx *ctx, int64_t n,
            const double *c1, const double *r1,
            const double *c2fg, const double *r2fg, const double *border,
            int img_size, int n_angles, const double *angles, const double *angle_tab,
            int rot_order, unsigned flags, int mtype,
            double *out, int *status);

/* sid_set_pair + sid_run in ONE call, with the host-to-device copy of the image pair overlapped
 * with the computation: the pair is uploaded in row bands on a second stream and the points are
 * processed band by band as soon as every image row they touch has arrived.  Results are
 * identical to sid_set_pair followed by sid_run; the pair stays resident afterwards.  Page-locked host
 * images are copied by DMA directly; pageable ones (plain malloc / NumPy memory) go band by band through
 * a pinned double buffer that a few host threads fill while the previous band is in flight (EW-sized
 * pair, B200: 6.1 ms per call pinned, 9.7 ms pageable, against 26 ms for a plain pageable cudaMemcpy). */
int sid_run_pair(sid_ctx *ctx,
                 const uint8_t *img1, int rows1, int cols1, int64_t pitch1,
                 const uint8_t *img2, int rows2, int cols2, int64_t pitch2,
                 int64_t n,
                 const double *c1, const double *r1,
                 const double *c2fg, const double *r2fg, const double *border,
                 int img_size, int n_angles, const double *angles, const double *angle_tab,
                 int rot_order, unsigned flags, int mtype,
                 double *out, int *status);

/* Same with DEVICE pointers for the point arrays and outputs; asynchronous on
 * the context's stream (no host synchronisation).  `max_border` must bound
 * border[] (it sizes shared memory / scratch); pass <= 0 to have the library
 * read the borders back (synchronises). */
int sid_run_device(sid_ctx *ctx, int64_t n,
                   const double *d_c1, const double *d_r1,
                   const double *d_c2fg, const double *d_r2fg, const double *d_border,
                   int max_border,
                   int img_size, int n_angles, const double *angles, const double *angle_tab,
                   int rot_order, unsigned flags, int mtype,
                   double *d_out, int *d_status);

/* Number of kernel launches issued by this context so far (bench accounting). */
int64_t sid_launch_count(const sid_ctx *ctx);
/* Device time (ms, CUDA events on the launching stream) of the most recent launch of the fused
 * point kernel; waits for it to finish.  -1 if nothing was launched yet. */
double sid_last_kernel_ms(sid_ctx *ctx);
/* Name of the point kernel the most recent launch used ("sid::pm_ws_kernel", "sid::pm_tc_kernel",
 * "sid::pm_points_kernel<imma>", "sid::pm_points_kernel<dp4a>"); "" before the first launch.  The paths
 * are interchangeable (bit-identical tables); the dispatch picks by search radius. */
const char *sid_last_kernel_name(const sid_ctx *ctx);

/* rotate_and_match for one point against an explicit search window `image2`
 * (host pointers).  Outputs: *valid (0 -> the reference returns 7 x NaN),
 * dc, dr, best_angle_idx (index into the angle list), best_r, best_h, and
 * optionally the best NCC map ((H-s+1) x (W-s+1) floats) and template (s x s). */
int sid_rotate_and_match(sid_ctx *ctx,
                         const uint8_t *img1, int rows1, int cols1, int64_t pitch1,
                         double c1, double r1, int img_size,
                         const uint8_t *image2, int H, int W, int64_t pitch2,
                         int n_angles, const double *angle_tab,
                         int rot_order, unsigned flags, int mtype,
                         int *valid, double *dc, double *dr, int *best_angle_idx,
                         float *best_r, float *best_h,
                         float *best_result, uint8_t *best_template);

/* get_template: s x s rotated/shifted uint8 patch (angle_tab: 4 doubles). */
int sid_get_template(sid_ctx *ctx,
                     const uint8_t *img, int rows, int cols, int64_t pitch,
                     double c, double r, const double *angle_tab, int s, int rot_order,
                     uint8_t *out);

/* template_matcher plug-in: zero-mean normalised cross-correlation map,
 * (H-th+1) x (W-tw+1) float32, method must be SID_TM_CCOEFF_NORMED. */
int sid_match_template(sid_ctx *ctx,
                       const uint8_t *img, int H, int W, int64_t pitch,
                       const uint8_t *tpl, int th, int tw, int64_t tpitch,
                       int method, float *out);

/* Feature-tracking matcher (the caller side of the hot path): the two nearest train descriptors
 * in Hamming distance for every query descriptor, i.e. what
 * cv2.BFMatcher(cv2.NORM_HAMMING).knnMatch(d1, d2, k=2) returns at the reference's call site
 * ftlib.py:92-99 (the `matcher` plug-in kwarg of get_match_coords).  Host pointers, descriptors are rows of
 * desc_bytes bytes (32 = ORB).  idx / dist: n1 x 2 ints, best first; equal distances go to the lower train
 * index (OpenCV's order); -1 where fewer than two train descriptors exist. */
int sid_knn_hamming2(sid_ctx *ctx, const uint8_t *d1, int n1, const uint8_t *d2, int n2, int desc_bytes,
                     int32_t *idx, int32_t *dist);

/* Deformation of triangular elements (the consumer side of the hot path): what
 * get_deformation_on_triangulation (libdefor.py:50-99) and get_deformation_elems (libdefor.py:4-48) return.
 * Host pointers.  x, y, u, v: n node values; tri: m x 3 node indices, row-major; area_in: optional m element
 * areas (NULL = Heron's formula from the side lengths, as get_deformation_on_triangulation does).
 * Outputs, m doubles each: e1 divergence, e2 shear, e3 vorticity, area, perimeter; NaN for an element with a
 * node index outside [0, n). */
int sid_deformation(sid_ctx *ctx, int n, const double *x, const double *y, const double *u, const double *v,
                    int m, const int32_t *tri, const double *area_in,
                    double *e1, double *e2, double *e3, double *area, double *perim);

/* get_hessian of a float32 map (rows x cols) -> float32 map of the same shape. */
int sid_get_hessian(sid_ctx *ctx, const float *ccm, int rows, int cols,
                    unsigned flags, float *out);

#ifdef __cplusplus
}
#endif
#endif /* SID_B200_H */
