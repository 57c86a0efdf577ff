"""Host-side logic around the GPU call: angle table, first guess, the whole
pattern_matching orchestration (with the GPU batch replaced by the CPU oracle so it
runs here), compared with the reference run live when it is available.  CPU only."""
import contextlib
import io
import warnings

import numpy as np
import pytest
from scipy import ndimage

import sea_ice_drift_b200 as sid
from sea_ice_drift_b200 import _lib, pmlib, sharding, synthetic as syn
from oracle import c_oracle as co
from oracle.angle_table import angle_table as oracle_angle_table
from oracle.ref_import import reference_available, load_reference


def oracle_compute(c1, r1, c2fg, r2fg, border, img1, img2, img_size, alpha0, **kw):
    kw = {k: v for k, v in kw.items() if k in ("angles", "rot_order", "hes_norm", "hes_smth", "mcc_norm")}
    return co.use_mcc_batch(c1, r1, c2fg, r2fg, border, img1, img2, img_size, alpha0, **kw)[0]


def test_angle_table_matches_oracle_definition():
    for s in (35, 50, 51):
        a = _lib.angle_table([-3, 0, 3, 17.5], -3.85, s)
        b = oracle_angle_table([-3, 0, 3, 17.5], -3.85, s)
        assert np.array_equal(a, b)
    t = _lib.angle_table([0], 0.0, 35)[0]
    assert t.tolist() == [1.0, 0.0, 18.0, 18.0]


def test_distance_at_equals_full_edt():
    rng = np.random.default_rng(0)
    shape = (300, 400)
    x1, y1 = rng.uniform(0, 399, 50), rng.uniform(0, 299, 50)
    full = sid.get_distance_to_nearest_keypoint(x1, y1, shape)
    cols, rows = rng.integers(0, 400, 500), rng.integers(0, 300, 500)
    assert np.array_equal(pmlib._distance_at(x1, y1, cols, rows), full[rows, cols])


def test_fill_gpi_and_interpolators():
    from sea_ice_drift_b200.lib import _fill_gpi, interpolation_poly, interpolation_near
    gpi = np.array([True, False, True, True, False, False])
    out = _fill_gpi((2, 3), gpi, np.array([1., 2., 3.]))
    assert out.shape == (2, 3) and np.isnan(out[0, 1]) and out[1, 0] == 3.0
    x1, y1 = np.meshgrid(np.arange(5.), np.arange(5.)); x1, y1 = x1.ravel(), y1.ravel()
    x2, y2 = 2 * x1 + 1, 0.5 * y1 - x1
    gx, gy = interpolation_poly(x1, y1, x2, y2, np.array([1.5, 10.]), np.array([2.5, -3.]))
    assert np.allclose(gx, [4., 21.]) and np.allclose(gy, [-0.25, -11.5])
    nx, ny = interpolation_near(x1, y1, x2, y2, np.array([1.5, 10.]), np.array([2.5, -3.]))
    assert np.isclose(nx[0], 4.) and np.isnan(nx[1])


def _scene():
    side = 700
    img1 = syn.speckle_image((side, side), seed=21)
    m = syn.rotation_matrix((side, side), 1.2)
    m[0, 2] += 5
    img2 = syn.warp_pair(img1, m, seed=21)
    n1 = syn.ArrayDomain(img1, lon0=10.0, lat0=80.0)
    n2 = syn.ArrayDomain(img2, lon0=10.0, lat0=80.0)
    rng = np.random.default_rng(4)
    kx, ky = rng.uniform(30, side - 30, 400), rng.uniform(30, side - 30, 400)
    k2x, k2y = syn.apply_affine(m, kx, ky)
    k2x, k2y = k2x + rng.normal(0, 0.7, kx.size), k2y + rng.normal(0, 0.7, kx.size)
    gx, gy = np.meshgrid(np.linspace(40, side - 40, 14), np.linspace(40, side - 40, 13))
    lon, lat = n1.transform_points(gx, gy)
    return n1, n2, kx, ky, k2x, k2y, lon, lat, m


def test_pattern_matching_orchestration_runs_and_finds_drift(monkeypatch):
    n1, n2, kx, ky, k2x, k2y, lon, lat, m = _scene()
    monkeypatch.setattr(sharding, "use_mcc_batch_sharded",
                        lambda *a, **k: oracle_compute(*a, **{kk: vv for kk, vv in k.items() if kk != "compute"}))
    with contextlib.redirect_stdout(io.StringIO()):
        u, v, a, r, h, lon2, lat2 = sid.pattern_matching(lon, lat, n1, kx, ky, n2, k2x, k2y, threads=3, angles=[-3, 0, 3],
                                                         device_epilogue=False)      # host post-processing under test
    assert u.shape == lon.shape == h.shape
    ok = np.isfinite(u)
    assert ok.sum() > 0.6 * ok.size
    c2, r2 = n2.transform_points(lon2[ok], lat2[ok], 1)
    c1, r1 = n1.transform_points(lon[ok], lat[ok], 1)
    tx, ty = syn.apply_affine(m, c1, r1)
    assert np.median(np.abs(c2 - tx + 1.0)) < 1.0 and np.median(np.abs(r2 - ty + 1.0)) < 1.0
    assert np.nanmedian(r) > 0.5


@pytest.mark.skipif(not reference_available(), reason="reference tree not present")
def test_pattern_matching_equals_reference_pattern_matching(monkeypatch):
    """Whole drop-in function against the reference's pattern_matching (live, threads=1)
    with the same Nansat stand-in; the GPU batch is replaced by the exact CPU oracle."""
    warnings.simplefilter("ignore")
    ref_pm = load_reference()
    import nansat
    nansat.NSR = lambda srs=None: srs
    ref_pm.NSR = nansat.NSR
    n1, n2, kx, ky, k2x, k2y, lon, lat, m = _scene()
    monkeypatch.setattr(sharding, "use_mcc_batch_sharded",
                        lambda *a, **k: oracle_compute(*a, **{kk: vv for kk, vv in k.items() if kk != "compute"}))
    kw = dict(angles=[-3, 0, 3], min_border=20, max_border=40)
    with contextlib.redirect_stdout(io.StringIO()):
        mine = sid.pattern_matching(lon, lat, n1, kx, ky, n2, k2x, k2y, threads=1, device_epilogue=False, **kw)
        ref = ref_pm.pattern_matching(lon, lat, n1, kx, ky, n2, k2x, k2y, threads=1, **kw)
    for name, a, b in zip("u v a r h lon2 lat2".split(), mine, ref):
        assert a.shape == b.shape
        assert np.array_equal(np.isnan(a), np.isnan(b)), name
        tol = dict(u=1e-9, v=1e-9, a=0, r=1e-4, h=1e-4, lon2=1e-9, lat2=1e-9)[name]
        assert np.nanmax(np.abs(a - b)) <= tol, (name, np.nanmax(np.abs(a - b)))


@pytest.mark.skipif(not reference_available(), reason="reference tree not present")
def test_prepare_first_guess_equals_reference():
    warnings.simplefilter("ignore")
    ref_pm = load_reference()
    n1, n2, kx, ky, k2x, k2y, lon, lat, m = _scene()
    c2pm1, r2pm1 = np.round(n2.transform_points(lon.flatten(), lat.flatten(), 1))
    for old in (True, False):
        a = sid.prepare_first_guess(c2pm1, r2pm1, n1, kx, ky, n2, k2x, k2y, 35, old_border=old)
        b = ref_pm.prepare_first_guess(c2pm1, r2pm1, n1, kx, ky, n2, k2x, k2y, 35, old_border=old)
        for x, y in zip(a, b):
            assert np.array_equal(x, y, equal_nan=True)
    assert sid.get_initial_rotation(n1, n2) == ref_pm.get_initial_rotation(n1, n2)


def test_user_template_matcher_plugin_is_detected():
    import cv2
    assert pmlib._is_builtin_matcher(None) and pmlib._is_builtin_matcher(cv2.matchTemplate)
    assert pmlib._is_builtin_matcher(sid.match_template)
    assert not pmlib._is_builtin_matcher(lambda img, tpl, m: None)


def test_integration_snippet_angle_table_equals_binding():
    """The angle table a reference maintainer builds in INTEGRATION.md section 2 is the one the binding passes."""
    for img_size, alpha0, angles in ((35, 0.0, [-3, 0, 3]), (50, -3.85, [0.5, 7.25]), (51, 12.0, list(range(-10, 11)))):
        tab = np.empty((len(angles), 4))
        tc = np.array([int(img_size / 2.) + 1] * 2)
        for k, ang in enumerate(np.asarray(angles, dtype=np.float64)):
            a = np.radians(ang - alpha0)
            T = np.array([[np.cos(a), -np.sin(a)], [np.sin(a), np.cos(a)]])
            tab[k] = (T[0, 0], T[1, 0]) + tuple(tc.dot(T))
        assert np.array_equal(tab, _lib.angle_table(angles, alpha0, img_size))
