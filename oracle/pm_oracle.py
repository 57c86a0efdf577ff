"""TEST INFRASTRUCTURE ONLY -- NumPy / SciPy / OpenCV port of the reference's
per-point MCC loop, used (a) as a second checker next to the plain-C oracle and
(b) as the timed CPU baseline of ``bench.py`` (``cpu_baseline.kind = "port"`` and
``--impl reference``), because the Python reference itself cannot travel to the
GPU box.  It calls the same third-party routines the reference calls
(``scipy.ndimage.affine_transform``, ``cv2.matchTemplate``, ``np.gradient`` ...)
so its speed is representative of the reference on the same cores.

Follows (paths under /root/reference/sea_ice_drift/):
  peak_sharpness   pmlib.py:36-59    (get_hessian)
  rotated_patch    pmlib.py:89-115   (get_template)
  sweep_angles     pmlib.py:117-174  (rotate_and_match)
  match_point      pmlib.py:176-212  (use_mcc)
  run_points       pmlib.py:214-247, 430-448 (_init_pool / use_mcc_mp / Pool.map)
Parity status: pinned against the live reference by tests/test_oracle_vs_reference.py
(in the build container) and by the committed fixtures in tests/golden/.
"""
import multiprocessing as mp
import warnings

import numpy as np
import cv2
from scipy import ndimage

NAN7 = (np.nan,) * 7
_STATE = {}


def peak_sharpness(ccm, hes_norm=True, hes_smth=False, **_):
    field = ndimage.gaussian_filter(ccm, 1) if hes_smth else ccm
    g_rows, g_cols = np.gradient(field)
    curv = np.hypot(np.gradient(g_cols)[1], np.gradient(g_rows)[0])
    if not hes_norm:
        return curv
    return (curv - np.median(curv)) / np.std(curv)


def rotated_patch(img, c, r, a, s, rot_order=0, **_):
    centre = int(s / 2.) + 1
    rad = np.radians(a)
    rot = np.array([[np.cos(rad), -np.sin(rad)], [np.sin(rad), np.cos(rad)]])
    shift = np.array([r, c]) - np.array([centre, centre]).dot(rot)
    return ndimage.affine_transform(img, rot.T, offset=shift, output_shape=(s, s),
                                    order=rot_order, cval=0.0, output=np.uint8)


def sweep_angles(img1, c1, r1, img_size, image2, alpha0, angles=(-3, 0, 3),
                 mtype=cv2.TM_CCOEFF_NORMED, template_matcher=cv2.matchTemplate,
                 mcc_norm=False, **kw):
    top = None
    for angle in angles:
        patch = rotated_patch(img1, c1, r1, angle - alpha0, img_size, **kw)
        if patch.min() == 0 or patch.shape[0] < img_size or patch.shape[1] < img_size:
            return NAN7
        ccm = template_matcher(image2, patch, mtype)
        peak = ccm.max()
        if top is None or peak > top[0]:
            top = (peak, angle, ccm, patch, np.unravel_index(np.argmax(ccm), ccm.shape))
    peak, angle, ccm, patch, where = top
    sharp = peak_sharpness(ccm, **kw)[where]
    d_row = where[0] - (image2.shape[0] - patch.shape[0]) / 2.
    d_col = where[1] - (image2.shape[1] - patch.shape[1]) / 2.
    if mcc_norm:
        peak = (peak - np.median(ccm)) / np.std(ccm)
    return d_col, d_row, angle, peak, sharp, ccm, patch


def match_point(c1, r1, c2fg, r2fg, border, img1, img2, img_size, alpha0, **kw):
    half = int(img_size / 2.)
    rows = slice(int(r2fg - half - border), int(r2fg + half + border + 1))
    cols = slice(int(c2fg - half - border), int(c2fg + half + border + 1))
    d_col, d_row, angle, peak, sharp = sweep_angles(
        img1, c1, r1, img_size, img2[rows, cols], alpha0, **kw)[:5]
    return c2fg + d_col, r2fg + d_row, angle, peak, sharp


def _install(state):
    _STATE.clear()
    _STATE.update(state)


def _one(i):
    s = _STATE
    return match_point(s["c1"][i], s["r1"][i], s["c2fg"][i], s["r2fg"][i], s["brd"][i],
                       s["img1"], s["img2"], s["img_size"], s["alpha0"], **s["kw"])


def run_points(c1, r1, c2fg, r2fg, brd, img1, img2, img_size, alpha0, threads=1, **kw):
    """Every grid point through match_point; serial when threads <= 1, otherwise a
    fork Pool with the images inherited copy-on-write.  Returns (n, 5) float64."""
    state = dict(c1=c1, r1=r1, c2fg=c2fg, r2fg=r2fg, brd=brd, img1=img1, img2=img2,
                 img_size=img_size, alpha0=alpha0, kw=kw)
    n = len(c1)
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        if threads <= 1:
            _install(state)
            rows = [_one(i) for i in range(n)]
        else:
            # OpenCV's own worker threads do not survive fork(): park them in the
            # parent first (one OpenCV thread per worker, as one worker owns one core).
            before = cv2.getNumThreads()
            cv2.setNumThreads(1)
            try:
                ctx = mp.get_context("fork")
                with ctx.Pool(threads, initializer=_install, initargs=(state,)) as pool:
                    rows = pool.map(_one, range(n), chunksize=max(1, n // (threads * 8)))
            finally:
                cv2.setNumThreads(before)
    if n == 0:
        return np.zeros((0, 5))
    return np.array(rows, dtype=np.float64)
