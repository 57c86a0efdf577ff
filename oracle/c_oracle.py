"""TEST INFRASTRUCTURE ONLY -- ctypes front-end of the plain-C oracle
(oracle/mcc_oracle.c).  Function names mirror the reference (pmlib.py)."""
import ctypes as C
import os

import numpy as np

from .angle_table import angle_table
from . import build as _build

_lib = None
_u8p = C.POINTER(C.c_uint8)
_f32p = C.POINTER(C.c_float)
_f64p = C.POINTER(C.c_double)
_i32p = C.POINTER(C.c_int)


def lib():
    global _lib
    if _lib is None:
        path = _build.build()
        _lib = C.CDLL(path)
        _lib.sido_version.restype = C.c_char_p
    return _lib


def _u8(a):
    a = np.asarray(a)
    if a.dtype != np.uint8 or a.ndim != 2 or a.strides[1] != 1:
        a = np.ascontiguousarray(a, dtype=np.uint8)
    return a


def _ptr(a, t):
    return a.ctypes.data_as(t)


def get_template(img, c, r, a, s, rot_order=0, **kwargs):
    img = _u8(img)
    tab = angle_table([a], 0.0, s)
    out = np.zeros((s, s), np.uint8)
    rc = lib().sido_get_template(_ptr(img, _u8p), C.c_int(img.shape[0]), C.c_int(img.shape[1]),
                                 C.c_int64(img.strides[0]), C.c_double(c), C.c_double(r),
                                 _ptr(tab, _f64p), C.c_int(s), C.c_int(rot_order), _ptr(out, _u8p))
    if rc:
        raise ValueError("sido_get_template rc=%d" % rc)
    return out


def match_template(image, templ, method=5):
    if method != 5:
        raise ValueError("only TM_CCOEFF_NORMED (5)")
    image = _u8(image)
    templ = _u8(templ)
    H, W = image.shape
    th, tw = templ.shape
    out = np.zeros((H - th + 1, W - tw + 1), np.float32)
    rc = lib().sido_match_template(_ptr(image, _u8p), C.c_int(H), C.c_int(W), C.c_int64(image.strides[0]),
                                   _ptr(templ, _u8p), C.c_int(th), C.c_int(tw), C.c_int64(templ.strides[0]),
                                   _ptr(out, _f32p))
    if rc:
        raise ValueError("sido_match_template rc=%d" % rc)
    return out


def get_hessian(ccm, hes_norm=True, hes_smth=False, **kwargs):
    ccm = np.ascontiguousarray(ccm, dtype=np.float32)
    out = np.zeros_like(ccm)
    rc = lib().sido_hessian(_ptr(ccm, _f32p), C.c_int(ccm.shape[0]), C.c_int(ccm.shape[1]),
                            C.c_int(bool(hes_norm)), C.c_int(bool(hes_smth)), _ptr(out, _f32p))
    if rc:
        raise ValueError("sido_hessian rc=%d" % rc)
    return out


def rotate_and_match(img1, c1, r1, img_size, image2, alpha0, angles=[-3, 0, 3], mtype=5,
                     template_matcher=None, mcc_norm=False, rot_order=0, hes_norm=True,
                     hes_smth=False, **kwargs):
    img1 = _u8(img1)
    image2 = _u8(image2)
    tab = angle_table(angles, alpha0, img_size)
    H, W = image2.shape
    res = np.zeros((H - img_size + 1, W - img_size + 1), np.float32)
    tpl = np.zeros((img_size, img_size), np.uint8)
    valid, ba = C.c_int(0), C.c_int(-1)
    dc, dr = C.c_double(0), C.c_double(0)
    br, bh = C.c_float(0), C.c_float(0)
    rc = lib().sido_rotate_and_match(
        _ptr(img1, _u8p), C.c_int(img1.shape[0]), C.c_int(img1.shape[1]), C.c_int64(img1.strides[0]),
        C.c_double(c1), C.c_double(r1), C.c_int(img_size),
        _ptr(image2, _u8p), C.c_int(H), C.c_int(W), C.c_int64(image2.strides[0]),
        C.c_int(len(angles)), _ptr(tab, _f64p), C.c_int(rot_order),
        C.c_int(bool(hes_norm)), C.c_int(bool(hes_smth)), C.c_int(bool(mcc_norm)),
        C.byref(valid), C.byref(dc), C.byref(dr), C.byref(ba), C.byref(br), C.byref(bh),
        _ptr(res, _f32p), _ptr(tpl, _u8p))
    if rc:
        raise ValueError("sido_rotate_and_match rc=%d" % rc)
    if not valid.value:
        return (np.nan,) * 7
    return (dc.value, dr.value, angles[ba.value], np.float32(br.value), np.float32(bh.value), res, tpl)


def use_mcc_batch(c1, r1, c2fg, r2fg, border, img1, img2, img_size, alpha0, angles=[-3, 0, 3],
                  mcc_norm=False, rot_order=0, hes_norm=True, hes_smth=False, threads=None, **kwargs):
    """All points of the reference's Pool map (pmlib.py:436-448) -> (n,5) float64, status (n,)."""
    img1 = _u8(img1)
    img2 = _u8(img2)
    arrs = [np.ascontiguousarray(np.atleast_1d(x), dtype=np.float64) for x in (c1, r1, c2fg, r2fg, border)]
    n = arrs[0].size
    tab = angle_table(angles, alpha0, img_size)
    ang = np.ascontiguousarray(np.asarray(angles, dtype=np.float64))
    out = np.full((n, 5), np.nan, np.float64)
    status = np.zeros(n, np.int32)
    fn = lib().sido_use_mcc_batch

    def chunk(lo, hi):
        return fn(
            C.c_int64(hi - lo), *[_ptr(a[lo:hi], _f64p) for a in arrs],
            _ptr(img1, _u8p), C.c_int(img1.shape[0]), C.c_int(img1.shape[1]), C.c_int64(img1.strides[0]),
            _ptr(img2, _u8p), C.c_int(img2.shape[0]), C.c_int(img2.shape[1]), C.c_int64(img2.strides[0]),
            C.c_int(img_size), C.c_int(len(angles)), _ptr(ang, _f64p), _ptr(tab, _f64p),
            C.c_int(rot_order), C.c_int(bool(hes_norm)), C.c_int(bool(hes_smth)), C.c_int(bool(mcc_norm)),
            _ptr(out[lo:hi], _f64p), _ptr(status[lo:hi], _i32p))

    # plain threads over chunks (ctypes drops the GIL); no OpenMP, so forking stays safe
    workers = int(threads) if threads else (os.cpu_count() or 1)
    step = max(1, min(256, -(-n // max(1, workers))))
    spans = [(lo, min(n, lo + step)) for lo in range(0, n, step)]
    if workers <= 1 or len(spans) <= 1:
        rcs = [chunk(lo, hi) for lo, hi in spans]
    else:
        from concurrent.futures import ThreadPoolExecutor
        with ThreadPoolExecutor(workers) as ex:
            rcs = list(ex.map(lambda sp: chunk(*sp), spans))
    for rc in rcs:
        if rc:
            raise ValueError("sido_use_mcc_batch rc=%d" % rc)
    return out, status


def knn_hamming2(d1, d2):
    """Two nearest train descriptors per query, OpenCV ordering (reference ftlib.py:95-96)."""
    d1 = np.ascontiguousarray(d1, dtype=np.uint8)
    d2 = np.ascontiguousarray(d2, dtype=np.uint8)
    idx = np.full((d1.shape[0], 2), -1, np.int32)
    dist = np.full((d1.shape[0], 2), -1, np.int32)
    rc = lib().sido_knn_hamming2(_ptr(d1, _u8p), C.c_int(d1.shape[0]), _ptr(d2, _u8p), C.c_int(d2.shape[0]),
                                 C.c_int(d1.shape[1]), _ptr(idx, _i32p), _ptr(dist, _i32p))
    if rc:
        raise ValueError("sido_knn_hamming2 rc=%d" % rc)
    return idx, dist
