"""The two CPU oracles against the golden vectors produced by the live reference
(oracle/make_golden.py).  CPU only."""
import numpy as np
import pytest

from oracle import c_oracle as co
from oracle import pm_oracle as po
from tests.helpers import variant_inputs, classify, make_exact_lookup, assert_parity

VARIANTS = ["default", "one_angle", "rot_order1", "hes_smth", "raw_hes_mcc_norm", "even50_7angles",
            "s51_b60", "s21_b9"]


@pytest.mark.parametrize("name", VARIANTS)
def test_c_oracle_matches_reference_table(golden_points, name):
    g = golden_points
    pts, s, alpha0, angles, opts, ref = variant_inputs(g, name)
    got, status = co.use_mcc_batch(*pts, g["img1"], g["img2"], s, alpha0, angles=angles, **opts)
    lookup = make_exact_lookup(co, pts, g["img1"], g["img2"], s, alpha0, angles, opts)
    stats = classify(got, ref, lookup)
    assert_parity(stats)
    assert np.array_equal(status == 1, ~np.isnan(ref[:, 0]))


@pytest.mark.parametrize("name", ["default", "rot_order1", "even50_7angles"])
def test_numpy_port_reproduces_reference_bitwise(golden_points, name):
    g = golden_points
    pts, s, alpha0, angles, opts, ref = variant_inputs(g, name)
    got = po.run_points(*pts, g["img1"], g["img2"], s, alpha0, threads=1, angles=angles, **opts)
    assert np.array_equal(got, ref, equal_nan=True)


def test_numpy_port_pool_equals_serial(golden_points):
    g = golden_points
    pts, s, alpha0, angles, opts, ref = variant_inputs(g, "default")
    got = po.run_points(*pts, g["img1"], g["img2"], s, alpha0, threads=2, angles=angles, **opts)
    assert np.array_equal(got, ref, equal_nan=True)


def test_templates_bit_exact(golden_stages):
    g = golden_stages
    for k, (c, r, ang, s, order) in enumerate(g["tpl_cases"]):
        got = co.get_template(g["img"], c, r, ang, int(s), rot_order=int(order))
        assert np.array_equal(got, g["tpl_%d" % k]), "template case %d" % k


def test_match_template_within_cv2_noise(golden_stages):
    g = golden_stages
    for k in range(4):
        got = co.match_template(g["mt_win_%d" % k], g["mt_tpl_%d" % k])
        ref = g["mt_out_%d" % k]
        assert got.shape == ref.shape and got.dtype == np.float32
        assert np.abs(got - ref).max() < 1e-5


def test_match_template_degenerate_inputs():
    rng = np.random.default_rng(0)
    win = rng.integers(1, 256, (60, 70), dtype=np.uint8)
    assert np.all(co.match_template(win, np.full((20, 20), 9, np.uint8)) == 1.0)       # flat template -> 1
    assert np.all(co.match_template(np.full((60, 70), 9, np.uint8), win[:20, :20]) == 0.0)  # flat window -> 0
    exact = co.match_template(win, win[11:31, 17:37].copy())
    assert exact[11, 17] == 1.0 and np.argmax(exact) == 11 * exact.shape[1] + 17


def test_hessian_matches_reference(golden_stages):
    g = golden_stages
    for k in range(4):
        ccm = g["mt_out_%d" % k]
        for hn, hs in ((1, 0), (0, 0), (1, 1), (0, 1)):
            got = co.get_hessian(ccm, hes_norm=bool(hn), hes_smth=bool(hs))
            ref = g["hes_%d_%d%d" % (k, hn, hs)]
            assert np.abs(got - ref).max() <= 2e-6 * (1 + np.abs(ref).max())


def test_rotate_and_match_explicit_window(golden_points, golden_stages):
    g, st = golden_points, golden_stages
    angles = [-3, -2, -1, 0, 1, 2, 3]
    dc, dr, a, r, h, ccm, tpl = co.rotate_and_match(g["img1"], 210.3, 120.6, 50, g["img2"][40:190, 130:300], -3.85,
                                                   angles=angles)
    ref = st["ram_scalars"]
    assert (dc, dr, a) == (ref[0], ref[1], ref[2])
    assert abs(r - ref[3]) < 1e-4 and abs(h - ref[4]) < 1e-4
    assert np.array_equal(tpl, st["ram_template"])
    assert np.abs(ccm - st["ram_result"]).max() < 1e-5


def test_even_template_gives_half_integer_displacements(golden_points):
    # use_mcc cuts a window of s + 2b + 1 pixels for even s, so (H - s) / 2 is a half-integer (pmlib.py:168-169, 200-202)
    g = golden_points
    pts, s, alpha0, angles, opts, ref = variant_inputs(g, "even50_7angles")
    got, _ = co.use_mcc_batch(*pts, g["img1"], g["img2"], s, alpha0, angles=angles, **opts)
    ok = ~np.isnan(got[:, 0])
    assert np.all(np.abs((got[ok, 0] - pts[2][ok]) % 1.0 - 0.5) < 1e-12)
    assert np.all(np.abs((got[ok, 1] - pts[3][ok]) % 1.0 - 0.5) < 1e-12)


def test_nan_on_zero_pixel_and_rejected_window(golden_points):
    g = golden_points
    img1, img2 = g["img1"], g["img2"]
    out, status = co.use_mcc_batch([245.0, 100.0, 100.0], [160.0, 100.0, 100.0], [245.0, 10.0, 100.0],
                                   [160.0, 100.0, 100.0], [20.0, 20.0, 20.0], img1, img2, 35, 0.0)
    assert status.tolist() == [0, -1, 1]             # zero patch, window off the left edge, fine
    assert np.isnan(out[0]).all() and np.isnan(out[1]).all() and not np.isnan(out[2]).any()


def test_empty_batch():
    out, status = co.use_mcc_batch([], [], [], [], [], np.ones((50, 50), np.uint8), np.ones((50, 50), np.uint8), 35, 0.0)
    assert out.shape == (0, 5) and status.shape == (0,)
