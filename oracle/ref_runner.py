"""TEST INFRASTRUCTURE ONLY -- runs the UNMODIFIED reference's per-point loop (the Pool section of
``pattern_matching``, /root/reference/sea_ice_drift/pmlib.py:430-448) on arrays.

``run_reference_points`` does exactly what the reference does between ``_init_pool`` and ``Pool.map``: it stores the
nine positional arguments and the kwargs dict in the reference module's own globals ``shared_args`` /
``shared_kwargs`` (pmlib.py:33-34, 430-434) and maps the reference's own ``use_mcc_mp`` (pmlib.py:214-247) over the
point indices -- in this process when ``threads <= 1`` (pmlib.py:436-439), else in a fork ``Pool`` (pmlib.py:442-448).
The reference module comes from ``oracle/ref_import.py`` (``/root/reference`` here, the ``oracle/_ref`` copy on the
GPU box).  Used by the ``-m gpu`` full-size parity tests and by ``bench.py``'s CPU legs (``kind: "reference"``).
"""
import contextlib
import io
import multiprocessing as mp
import warnings

import numpy as np

from . import ref_import

_REF = {}


def reference_module():
    if "pm" not in _REF:
        _REF["pm"] = ref_import.load_reference()
    return _REF["pm"]


def available():
    return ref_import.reference_available()


def _init_pool(*args):
    pm = reference_module()
    pm.shared_args = args[:9]
    pm.shared_kwargs = args[9]


def _one(i):
    # the reference prints a progress line every 100 points (pmlib.py:243-246); keep the workers quiet
    with contextlib.redirect_stdout(io.StringIO()), warnings.catch_warnings():
        warnings.simplefilter("ignore")
        return reference_module().use_mcc_mp(i)


def run_reference_points(c1, r1, c2fg, r2fg, border, img1, img2, img_size, alpha0, threads=1, **kwargs):
    """(n, 5) float64 table [c2, r2, angle, r, h] from the reference's own ``use_mcc_mp``."""
    import cv2
    n = len(c1)
    if n == 0:
        return np.zeros((0, 5))
    args = (np.asarray(c1), np.asarray(r1), np.asarray(c2fg), np.asarray(r2fg), np.asarray(border),
            img1, img2, img_size, alpha0, kwargs)
    reference_module()
    if threads <= 1:
        _init_pool(*args)
        rows = [_one(i) for i in range(n)]
    else:
        before = cv2.getNumThreads()
        cv2.setNumThreads(1)            # OpenCV's worker threads do not survive fork(); one per worker anyway
        try:
            with mp.get_context("fork").Pool(threads, initializer=_init_pool, initargs=args) as pool:
                rows = pool.map(_one, range(n), chunksize=max(1, n // (threads * 8)))
        finally:
            cv2.setNumThreads(before)
    return np.array(rows, dtype=np.float64)


def orb_first_guess(img1, img2, n_features=20000, ratio_test=0.6, seed_cv=0):
    """Keypoint matches of BASELINE configs[0] ("ORB first guess"): the reference's own ``find_key_points`` and
    ``get_match_coords`` (ftlib.py:26-116: cv2.ORB + BFMatcher Hamming kNN + Lowe ratio test) followed by its
    ``lstsq_filter`` (ftlib.py:203-234).  Returns x1, y1, x2, y2 (pixels)."""
    import importlib
    import cv2
    reference_module()
    ft = importlib.import_module("sea_ice_drift.ftlib")
    cv2.setRNGSeed(seed_cv)
    with contextlib.redirect_stdout(io.StringIO()):
        kp1, d1 = ft.find_key_points(img1, nFeatures=n_features)
        kp2, d2 = ft.find_key_points(img2, nFeatures=n_features)
        x1, y1, x2, y2 = ft.get_match_coords(kp1, d1, kp2, d2, ratio_test=ratio_test)
        x1, y1, x2, y2 = ft.lstsq_filter(x1, y1, x2, y2)
    return x1, y1, x2, y2
