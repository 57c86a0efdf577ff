"""Drop-in replacement for the pattern-matching part of ``sea_ice_drift.pmlib``.

Same function names, argument meaning, defaults, return shapes and NaN behaviour
as the reference (``/root/reference/sea_ice_drift/pmlib.py``); the per-point MCC
loop (reference pmlib.py:430-448 and everything it calls) runs as ONE batched
launch of hand-written sm_100a kernels through the C ABI in
``include/sid_b200.h``.  There is no CPU implementation behind these functions:
without the CUDA library / a CUDA device they raise.

Kept on the host, as in the reference: the lon/lat <-> pixel transforms done by
the Nansat objects, the first-guess interpolation and the final re-gridding.
"""
from __future__ import absolute_import, print_function

import time

import numpy as np
from scipy import ndimage
from scipy.spatial import cKDTree

from . import _lib
from ._lib import SID_TM_CCOEFF_NORMED, flags_from_kwargs
from .lib import interpolation_poly, interpolation_near, _fill_gpi

TM_CCOEFF_NORMED = SID_TM_CCOEFF_NORMED          # == cv2.TM_CCOEFF_NORMED
DEFAULT_SRS = '+proj=latlong +datum=WGS84 +ellps=WGS84 +no_defs'

# module-level state with the reference's names (reference pmlib.py:33-34), so code that
# drives use_mcc_mp by hand keeps working
shared_args = None
shared_kwargs = None

def _ctx(kwargs=None):
    device = None if not kwargs else kwargs.get('device')
    return _lib.default_context(device)


def _is_builtin_matcher(fn):
    if fn is None or fn is match_template:
        return True
    return getattr(fn, '__name__', '') == 'matchTemplate' and 'cv2' in (getattr(fn, '__module__', '') or 'cv2')


def match_template(image, templ, method=TM_CCOEFF_NORMED, **kwargs):
    """GPU ``template_matcher``: same call shape as ``cv2.matchTemplate(image, templ,
    cv2.TM_CCOEFF_NORMED)`` (reference pmlib.py:120, 156); float32 map of shape
    (H-th+1, W-tw+1)."""
    return _ctx(kwargs).match_template(image, templ, method)


def get_hessian(ccm, hes_norm=True, hes_smth=False, **kwargs):
    """Peak-sharpness map of a cross-correlation matrix (reference pmlib.py:36-59)."""
    return _ctx(kwargs).get_hessian(ccm, flags_from_kwargs(hes_norm, hes_smth, False))


def get_template(img, c, r, a, s, rot_order=0, **kwargs):
    """Rotated/shifted s x s uint8 template around (c, r) (reference pmlib.py:89-115)."""
    return _ctx(kwargs).get_template(img, c, r, a, s, rot_order)


try:                                        # the reference's default (pmlib.py:120); recognised and run on the GPU
    import cv2 as _cv2
    _DEFAULT_MATCHER = _cv2.matchTemplate
except ImportError:                         # no OpenCV installed: same behaviour, our own callable as the default
    _DEFAULT_MATCHER = match_template


def rotate_and_match(img1, c1, r1, img_size, image2, alpha0,
                     angles=[-3, 0, 3],
                     mtype=TM_CCOEFF_NORMED,
                     template_matcher=_DEFAULT_MATCHER,
                     mcc_norm=False,
                     **kwargs):
    """Best match of the rotated templates of one point inside ``image2``
    (reference pmlib.py:117-174).  Returns ``(dc, dr, best_a, best_r, best_h,
    best_result, best_template)`` or seven NaNs when a template holds a 0 pixel.

    ``template_matcher`` left at None, set to :func:`match_template` or to
    ``cv2.matchTemplate`` selects the GPU matcher; any other callable is honoured
    as the reference's plug-in and called once per angle."""
    rot_order = kwargs.get('rot_order', 0)
    flags = flags_from_kwargs(kwargs.get('hes_norm', True), kwargs.get('hes_smth', False), mcc_norm)
    ctx = _ctx(kwargs)
    if _is_builtin_matcher(template_matcher):
        res = ctx.rotate_and_match(img1, c1, r1, img_size, image2, list(angles), alpha0, rot_order, flags, mtype)
        if res is None:
            return (np.nan,) * 7
        dc, dr, ia, r, h, ccm, tpl = res
        return dc, dr, angles[ia], r, h, ccm, tpl
    # user-supplied matcher: templates and Hessian still come from the GPU
    best = None
    for angle in angles:
        tpl = ctx.get_template(img1, c1, r1, angle - alpha0, img_size, rot_order)
        if tpl.min() == 0:
            return (np.nan,) * 7
        ccm = template_matcher(image2, tpl, mtype)
        if best is None or ccm.max() > best[0]:
            best = (ccm.max(), angle, ccm, tpl, np.unravel_index(np.argmax(ccm), ccm.shape))
    peak, angle, ccm, tpl, ij = best
    h = ctx.get_hessian(ccm, flags & 3)[ij]
    dr = ij[0] - (image2.shape[0] - tpl.shape[0]) / 2.
    dc = ij[1] - (image2.shape[1] - tpl.shape[1]) / 2.
    if mcc_norm:
        peak = (peak - np.median(ccm)) / np.std(ccm)
    return dc, dr, angle, peak, h, ccm, tpl


def use_mcc_batch(c1, r1, c2fg, r2fg, border, img1, img2, img_size, alpha0, **kwargs):
    """``use_mcc`` for every point at once: the body of the reference's Pool map
    (pmlib.py:436-448) as one fused kernel launch.  Returns an (N, 5) float64 table
    ``c2, r2, angle, r, h`` with NaN rows where the reference returns NaN.

    ``resident=True`` (extension): reuse the image pair uploaded by this context's previous call instead of
    uploading ``img1`` / ``img2`` again -- an explicit contract, for callers that run several grids or kwargs
    sets over one pair; raises nothing and uploads anyway when no pair is resident."""
    matcher = kwargs.get('template_matcher')
    if not _is_builtin_matcher(matcher):
        rows = [use_mcc(a, b, c, d, e, img1, img2, img_size, alpha0, **kwargs)
                for a, b, c, d, e in zip(c1, r1, c2fg, r2fg, border)]
        return np.array(rows, dtype=np.float64).reshape(-1, 5)
    ctx = _ctx(kwargs)
    flags = flags_from_kwargs(kwargs.get('hes_norm', True), kwargs.get('hes_smth', False),
                              kwargs.get('mcc_norm', False))
    args = (c1, r1, c2fg, r2fg, border, img_size, list(kwargs.get('angles', [-3, 0, 3])), alpha0,
            kwargs.get('rot_order', 0), flags, kwargs.get('mtype', TM_CCOEFF_NORMED))
    # The pair is uploaded on EVERY call (overlapped with the matching, so it costs ~1.5 ms per EW pair) unless
    # the caller states explicitly that the pair of its previous call is still the one to match: the library
    # never infers identity from pointers or sampled checksums (an in-place edit such as masking land, or a new
    # array at a recycled address, would silently match against the old pair).
    if kwargs.get('resident') and ctx._pair_key is not None:
        return ctx.run(*args)
    out = ctx.run_pair(img1, img2, *args)
    ctx._pair_key = (np.shape(img1), np.shape(img2))
    return out


def use_mcc(c1, r1, c2fg, r2fg, border, img1, img2, img_size, alpha0, **kwargs):
    """MCC for one point (reference pmlib.py:176-212): ``(c2, r2, a, r, h)``."""
    matcher = kwargs.get('template_matcher')
    if not _is_builtin_matcher(matcher):
        hws = int(img_size / 2.)
        image = img2[int(r2fg - hws - border):int(r2fg + hws + border + 1),
                     int(c2fg - hws - border):int(c2fg + hws + border + 1)]
        dc, dr, a, r, h = rotate_and_match(img1, c1, r1, img_size, image, alpha0, **kwargs)[:5]
        return c2fg + dc, r2fg + dr, a, r, h
    row = use_mcc_batch([c1], [r1], [c2fg], [r2fg], [border], img1, img2, img_size, alpha0, **kwargs)[0]
    if np.isnan(row[0]):
        return np.nan, np.nan, np.nan, np.nan, np.nan
    angles = list(kwargs.get('angles', [-3, 0, 3]))
    angle = angles[int(np.argmin(np.abs(np.asarray(angles, dtype=np.float64) - row[2])))]
    return row[0], row[1], angle, np.float32(row[3]), np.float32(row[4])


def use_mcc_mp(i):
    """``use_mcc`` on point ``i`` of the module-level ``shared_args`` /
    ``shared_kwargs`` (reference pmlib.py:214-247)."""
    sa = shared_args
    return use_mcc(sa[0][i], sa[1][i], sa[2][i], sa[3][i], sa[4][i], sa[5], sa[6], sa[7], sa[8], **shared_kwargs)


def get_distance_to_nearest_keypoint(x1, y1, shape):
    """Full-resolution image of the distance (px) to the nearest keypoint
    (reference pmlib.py:61-77)."""
    seed = np.zeros(shape, dtype=bool)
    seed[np.uint16(y1), np.uint16(x1)] = True
    return ndimage.distance_transform_edt(~seed, return_distances=True, return_indices=False)


def _distance_at(x1, y1, cols, rows):
    """Same values as ``get_distance_to_nearest_keypoint(...)[rows, cols]`` without the
    full-image transform: exact nearest-seed distance between integer pixels."""
    seeds = np.column_stack([np.uint16(y1).astype(np.float64), np.uint16(x1).astype(np.float64)])
    dist, _ = cKDTree(seeds).query(np.column_stack([rows.astype(np.float64), cols.astype(np.float64)]))
    return dist


def get_initial_rotation(n1, n2):
    """Angle (degrees) between the two scenes' pixel grids (reference pmlib.py:79-87)."""
    lons, lats = n2.get_corners()
    x0, y0 = n1.transform_points([lons[0]], [lats[0]], 1)
    x1, y1 = n1.transform_points([lons[1]], [lats[1]], 1)
    return np.degrees(np.arctan2(x1 - x0, y1 - y0)[0])


def _device_present():
    """True when a CUDA device (and the library) can be used for the first guess."""
    try:
        _lib.default_context()
        return True
    except Exception:
        return False


def _near_device(ctx, sx, sy, vx, vy, qx, qy, seeds=None):
    """interpolation_near(method='linear') (and optionally the nearest-seed distance) through sid_first_guess.
    Returns (vx_q, vy_q, dist, resolved, unique): ``resolved`` False when some point could not be decided numerically,
    ``unique`` False when some grid point lies in a cell of cocircular keypoints (triangulation not unique)."""
    kx, ky = seeds if seeds is not None else (np.zeros(0), np.zeros(0))
    ovx, ovy, dist, flag = ctx.first_guess(sx, sy, vx, vy, kx, ky, qx, qy)
    shape = np.shape(qx)
    return ovx.reshape(shape), ovy.reshape(shape), dist.reshape(shape), not np.any(flag == 2), not np.any(flag == 3)


def prepare_first_guess(c2pm1, r2pm1, n1, c1, r1, n2, c2, r2, img_size,
                        min_fg_pts=5, min_border=20, max_border=50, old_border=True, **kwargs):
    """First-guess position on image 2 and search radius for every grid point
    (reference pmlib.py:249-324): Delaunay-linear interpolation of the feature
    tracking vectors with a polynomial fallback outside their hull; the radius is
    the distance to the nearest keypoint clamped to [min_border, max_border].

    ``first_guess='auto'`` (the default where a CUDA device is present) evaluates the Delaunay-linear interpolant and
    the nearest-keypoint distances on the GPU without building a triangulation (``sid_first_guess``: every grid point
    finds its own Delaunay triangle by pivoting) and returns that when the result is unique -- then it equals the SciPy /
    reference values.  Where four keypoints are cocircular (integer pixel coordinates, e.g. ORB keypoints under an
    identity geolocation) the Delaunay triangulation is not unique and Qhull's arbitrary choice may differ: ``'auto'``
    then takes the SciPy path so that the result is always the reference's; ``first_guess='device'`` keeps the (equally
    valid) device result.  ``first_guess='host'`` is the SciPy path (one Qhull triangulation shared by both value sets,
    KD-tree distances), also taken for ``method != 'linear'``."""
    n2_shape = n2.shape()
    lon1, lat1 = n1.transform_points(c1, r1)
    c1n2, r1n2 = n2.transform_points(lon1, lat1, 1)
    c2p2, r2p2 = np.round(interpolation_poly(c1n2, r1n2, c2, r2, c2pm1, r2pm1, **kwargs))
    mode = kwargs.get('first_guess')
    if mode is None:
        mode = 'auto' if (kwargs.get('method', 'linear') == 'linear' and len(np.atleast_1d(c1)) >= 3
                          and _device_present()) else 'host'
    done = False
    if mode in ('device', 'auto'):
        ctx = _ctx(kwargs)
        c2pm1 = np.asarray(c2pm1, dtype=np.float64)
        r2pm1 = np.asarray(r2pm1, dtype=np.float64)
        inside = ((c2pm1 >= 0) * (c2pm1 < n2_shape[1]) * (r2pm1 >= 0) * (r2pm1 < n2_shape[0]))
        seeds = (np.uint16(c2).astype(np.float64), np.uint16(r2).astype(np.float64))
        integer_grid = np.array_equal(np.round(c2pm1), c2pm1) and np.array_equal(np.round(r2pm1), r2pm1)
        if old_border:
            vx, vy, dist, ok, unique = _near_device(ctx, c1n2, r1n2, c2, r2, c2pm1, r2pm1, seeds if integer_grid else None)
            ok = ok and (unique or mode == 'device')
            if ok and not integer_grid:      # the reference samples the distance image at the ROUNDED grid positions
                dist = _near_device(ctx, np.zeros(0), np.zeros(0), np.zeros(0), np.zeros(0),
                                    np.round(c2pm1).astype(np.int16).astype(np.float64),
                                    np.round(r2pm1).astype(np.int16).astype(np.float64), seeds)[2]
            if ok:
                c2fg, r2fg = np.round(vx), np.round(vy)
                border = np.zeros(c2pm1.size) + max_border
                border[inside.ravel()] = dist.ravel()[inside.ravel()]
                done = True
        else:
            c2tst, r2tst = interpolation_poly(c1n2, r1n2, c2, r2, c1n2, r1n2, **kwargs)
            vx, vy, _, ok, unique = _near_device(ctx, c1n2, r1n2, c2, r2, c2pm1, r2pm1)
            dx, dy, _, ok2, _ = _near_device(ctx, c1n2, r1n2, c2 - c2tst, r2 - r2tst, c2pm1, r2pm1)
            if ok and ok2 and (unique or mode == 'device'):
                c2fg, r2fg = np.round(vx), np.round(vy)
                border = np.hypot(dx, dy)
                done = True
    if not done:
        c2fg, r2fg = np.round(interpolation_near(c1n2, r1n2, c2, r2, c2pm1, r2pm1, **kwargs))
        if old_border:
            border = np.zeros(c2pm1.size) + max_border
            inside = ((c2pm1 >= 0) * (c2pm1 < n2_shape[1]) * (r2pm1 >= 0) * (r2pm1 < n2_shape[0]))
            border[inside] = _distance_at(c2, r2,
                                          np.round(c2pm1[inside]).astype(np.int16),
                                          np.round(r2pm1[inside]).astype(np.int16))
        else:
            c2tst, r2tst = interpolation_poly(c1n2, r1n2, c2, r2, c1n2, r1n2, **kwargs)
            c2dif, r2dif = interpolation_near(c1n2, r1n2, c2 - c2tst, r2 - r2tst, c2pm1, r2pm1, **kwargs)
            border = np.hypot(c2dif, r2dif)
    border[border < min_border] = min_border
    border[border > max_border] = max_border
    outside_hull = np.isnan(c2fg)
    border[outside_hull] = max_border
    border = np.floor(border)
    c2fg[outside_hull] = c2p2[outside_hull]
    r2fg[np.isnan(r2fg)] = r2p2[np.isnan(r2fg)]
    return c2fg, r2fg, border


def _nsr(srs):
    try:
        from nansat import NSR
        return NSR(srs)
    except ImportError:
        return srs


def pattern_matching(lon_pm1, lat_pm1, n1, c1, r1, n2, c2, r2,
                     margin=0, img_size=35, threads=5, srs=DEFAULT_SRS, **kwargs):
    """Pattern matching between two scenes (reference pmlib.py:326-497).

    Returns ``(u, v, a, r, h, lon2_dst, lat2_dst)``, each shaped like ``lon_pm1``
    with NaN where no vector could be found.  ``threads`` is accepted for
    compatibility and ignored: all grid points go through one GPU launch (or one
    launch per rank when ``torch.distributed`` is initialised, see sharding.py)."""
    t0 = time.time()
    img1, img2 = n1[1], n2[1]
    dst_shape = lon_pm1.shape

    c2pm1, r2pm1 = n2.transform_points(lon_pm1.flatten(), lat_pm1.flatten(), 1)
    c2pm1i, r2pm1i = np.round([c2pm1, r2pm1])
    lon1i, lat1i = n2.transform_points(c2pm1i, r2pm1i)
    c1pm1i, r1pm1i = n1.transform_points(lon1i, lat1i, 1)

    c2fg, r2fg, brd2 = prepare_first_guess(c2pm1i, r2pm1i, n1, c1, r1, n2, c2, r2, img_size, **kwargs)

    hws = round(img_size / 2) + 1
    hws_hypot = np.hypot(hws, hws)
    shape1, shape2 = n1.shape(), n2.shape()
    gpi = ((c2fg - brd2 - hws - margin > 0) *
           (r2fg - brd2 - hws - margin > 0) *
           (c2fg + brd2 + hws + margin < shape2[1]) *
           (r2fg + brd2 + hws + margin < shape2[0]) *
           (c1pm1i - hws_hypot - margin > 0) *
           (r1pm1i - hws_hypot - margin > 0) *
           (c1pm1i + hws_hypot + margin < shape1[1]) *
           (r1pm1i + hws_hypot + margin < shape1[0]))
    alpha0 = get_initial_rotation(n1, n2)

    from .sharding import use_mcc_batch_sharded, _dist
    # Affine geolocation (an object exposing ``affine_maps(srs) -> (pixel->x/y, pixel->lon/lat)`` 2 x 3 matrices, e.g.
    # synthetic.ArrayDomain): the post-processing below (reference pmlib.py:462-497) runs on the device too, on the
    # table the kernels left there -- the (N, 5) table never visits the host.  Real Nansat objects (GDAL, possibly
    # thin-plate-spline geolocation) take the host path.
    affine = getattr(n2, 'affine_maps', None)
    if (affine is not None and _dist() is None and kwargs.get('device_epilogue', True) and gpi.any()
            and _is_builtin_matcher(kwargs.get('template_matcher'))):
        xy, ll = affine(srs)
        ctx = _ctx(kwargs)
        flags = flags_from_kwargs(kwargs.get('hes_norm', True), kwargs.get('hes_smth', False), kwargs.get('mcc_norm', False))
        ctx.run_pair(img1, img2, c1pm1i[gpi], r1pm1i[gpi], c2fg[gpi], r2fg[gpi], brd2[gpi], img_size,
                     list(kwargs.get('angles', [-3, 0, 3])), alpha0, kwargs.get('rot_order', 0), flags,
                     kwargs.get('mtype', TM_CCOEFF_NORMED), keep_on_device=True)
        grids = ctx.pm_epilogue_affine(gpi, c2pm1, r2pm1, xy, ll)
        print('\n', 'Pattern matching - OK! (%3.0f sec)' % (time.time() - t0))
        return tuple(grid.reshape(dst_shape) for grid in grids)
    results = use_mcc_batch_sharded(c1pm1i[gpi], r1pm1i[gpi], c2fg[gpi], r2fg[gpi], brd2[gpi],
                                    img1, img2, img_size, alpha0, **kwargs)

    print('\n', 'Pattern matching - OK! (%3.0f sec)' % (time.time() - t0))
    if len(results) == 0:
        nan_grid = np.zeros(dst_shape) + np.nan
        return tuple(nan_grid.copy() for _ in range(7))

    c2pm2 = results[:, 0] + (c2pm1 - c2pm1i)[gpi]
    r2pm2 = results[:, 1] + (r2pm1 - r2pm1i)[gpi]
    nsr = _nsr(srs)
    xpm1, ypm1 = n2.transform_points(c2pm1, r2pm1, 0, nsr)
    xpm2, ypm2 = n2.transform_points(c2pm2, r2pm2, 0, nsr)
    lon_pm2, lat_pm2 = n2.transform_points(c2pm2, r2pm2, 0)
    u = _fill_gpi(dst_shape, gpi, xpm2) - xpm1.reshape(dst_shape)
    v = _fill_gpi(dst_shape, gpi, ypm2) - ypm1.reshape(dst_shape)
    a = _fill_gpi(dst_shape, gpi, results[:, 2])
    r = _fill_gpi(dst_shape, gpi, results[:, 3])
    h = _fill_gpi(dst_shape, gpi, results[:, 4])
    return (u, v, a, r, h,
            _fill_gpi(dst_shape, gpi, lon_pm2), _fill_gpi(dst_shape, gpi, lat_pm2))
