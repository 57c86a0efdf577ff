"""BASELINE.json's configurations AT THEIR STATED SIZE: the CUDA path (through the C ABI) against the UNMODIFIED
reference's own ``pmlib.use_mcc_mp`` / ``pattern_matching``, run on the same box from the ``oracle/_ref`` copy that
``oracle/build_ref.py`` places (git-ignored, travels with the snapshot).  Needs a B200 (-m gpu).

What is asserted (north_star): displacement and best angle identical except classified argmax ties, identical NaN
pattern, |dr| <= 1e-4; for h the strict 1e-4 absolute figure is reported and bounded explicitly (tests/helpers.py)."""
import contextlib
import io
import os

import numpy as np
import pytest

import sea_ice_drift_b200 as sid
from sea_ice_drift_b200 import _lib, synthetic as syn
from oracle import c_oracle as co, ref_runner
from tests.helpers import classify, make_exact_lookup, assert_parity

pytestmark = [pytest.mark.gpu,
              pytest.mark.skipif(not ref_runner.available(), reason="no reference copy (run oracle/build_ref.py in the build container)")]

THREADS = max(1, min(32, len(os.sched_getaffinity(0))))
OPTS = dict(rot_order=0, hes_norm=True, hes_smth=False, mcc_norm=False)


def _compare_sample(gpu_ctx, cfg_name, n_sample, seed, parity_kw=None, **mk):
    img1, img2, c1, r1, c2, r2, b, cfg = syn.make_config(cfg_name, seed=0, **mk)
    s, angles = cfg["img_size"], cfg["angles"]
    gpu_ctx.set_pair(img1, img2)
    out = gpu_ctx.run(c1, r1, c2, r2, b, s, angles, 0.0)
    sel = np.sort(np.random.default_rng(seed).choice(len(c1), min(n_sample, len(c1)), replace=False))
    pts = [x[sel] for x in (c1, r1, c2, r2, b)]
    ref = ref_runner.run_reference_points(*pts, img1, img2, s, 0.0, threads=THREADS, angles=angles)
    stats = classify(out[sel], ref, make_exact_lookup(co, pts, img1, img2, s, 0.0, angles, OPTS))
    print(cfg_name, {k: v for k, v in stats.items() if k != "unexplained"}, "unexplained", len(stats["unexplained"]))
    assert_parity(stats, **(parity_kw or {}))
    return img1.shape, len(c1), stats


def test_config2_full_size_vs_reference(gpu_ctx):
    shape, n, stats = _compare_sample(gpu_ctx, "cfg2", 2000, 11)
    assert shape == (10400, 10400) and n > 39000 and stats["n_compared"] > 1900


def test_config3_full_size_21_angles_vs_reference(gpu_ctx):
    shape, n, stats = _compare_sample(gpu_ctx, "cfg3", 1000, 12)
    assert shape == (10400, 10400) and n > 39000 and stats["n_compared"] > 950


def test_config4_full_size_margin100_vs_reference(gpu_ctx):
    # MEASURED DEVIATION, stated not hidden: with the 201 x 201 maps of this configuration |h| reaches 30-50 and the
    # reference's float32 DFT noise shows up as |dh| up to 2.7e-4 (31 of 320 sampled points above 1e-4 absolute, all
    # positions / angles identical, |dr| <= 2.4e-6): the bounds below are those measured figures with ~2x headroom
    shape, n, stats = _compare_sample(gpu_ctx, "cfg4", 320, 13, parity_kw=dict(h_tol=2.5e-4, h_abs_ceil=6e-4, h_abs_frac=0.25))
    assert shape == (10400, 10400) and n > 150000 and stats["n_compared"] > 300


def test_config5_two_full_size_pairs_vs_reference(gpu_ctx):
    """BASELINE configs[4]: the time series, two of its 16 full-size pairs (300 x 300 grids) through the batched
    series API, each against the reference on a sample."""
    from sea_ice_drift_b200 import sharding
    pairs, samples = [], []
    for k in (0, 1):
        img1, img2, c1, r1, c2, r2, b, cfg = syn.make_config("cfg5", seed=k)
        pairs.append((img1, img2, c1, r1, c2, r2, b))
        samples.append(np.sort(np.random.default_rng(20 + k).choice(len(c1), 400, replace=False)))
    tables = sharding.use_mcc_series(pairs, 35, alpha0=0.0, angles=[-3, 0, 3])
    for (img1, img2, c1, r1, c2, r2, b), sel, table in zip(pairs, samples, tables):
        assert img1.shape == (10400, 10400) and len(c1) > 88000 and table.shape == (len(c1), 5)
        pts = [x[sel] for x in (c1, r1, c2, r2, b)]
        ref = ref_runner.run_reference_points(*pts, img1, img2, 35, 0.0, threads=THREADS, angles=[-3, 0, 3])
        stats = classify(table[sel], ref, make_exact_lookup(co, pts, img1, img2, 35, 0.0, [-3, 0, 3], OPTS))
        assert_parity(stats)


def test_config1_orb_first_guess_whole_pattern_matching_vs_reference(gpu_ctx):
    """BASELINE configs[0]: 2000 x 2000 pair, uniform 12-px shift, ORB first guess (the reference's own
    find_key_points / get_match_coords), 50 x 50 PM grid, img_size 35, angles [0], default borders 20..50 --
    the WHOLE drop-in ``pattern_matching`` against the reference's ``pattern_matching`` on the same inputs."""
    img1, img2, _, _, _, _, _, cfg = syn.make_config("cfg1", seed=0)
    assert img1.shape == (2000, 2000)
    x1, y1, x2, y2 = ref_runner.orb_first_guess(img1, img2)
    assert len(x1) > 500, "ORB found too few matches for a first guess"
    n1, n2 = syn.ArrayDomain(img1), syn.ArrayDomain(img2)
    gx, gy = np.meshgrid(np.linspace(100, 1900, 50), np.linspace(100, 1900, 50))
    lon, lat = n2.transform_points(gx, gy)
    ref_pm = ref_runner.reference_module()
    import nansat
    nansat.NSR = lambda srs=None: srs
    ref_pm.NSR = nansat.NSR
    kw = dict(angles=[0], img_size=35)
    with contextlib.redirect_stdout(io.StringIO()):
        mine = sid.pattern_matching(lon, lat, n1, x1, y1, n2, x2, y2, **kw)
        ref = ref_pm.pattern_matching(lon, lat, n1, x1, y1, n2, x2, y2, threads=THREADS, **kw)
    names = "u v a r h lon2 lat2".split()
    m, r = dict(zip(names, mine)), dict(zip(names, ref))
    for name in names:
        assert m[name].shape == r[name].shape == lon.shape
        assert np.array_equal(np.isnan(m[name]), np.isnan(r[name])), name
    ok = np.isfinite(r["u"])
    assert ok.sum() > 2000                                    # nearly the whole 50 x 50 grid is valid
    same = ok & (m["u"] == r["u"]) & (m["v"] == r["v"]) & (m["a"] == r["a"])
    # a differing vector must be an argmax tie of the reference's float32 correlation: equal r to ~1e-6
    differ = ok & ~same
    assert differ.sum() <= max(2, ok.sum() // 200), int(differ.sum())
    assert np.all(np.abs(m["r"][differ] - r["r"][differ]) < 4e-6)
    assert np.abs(m["r"][same] - r["r"][same]).max() <= 1e-4
    dh = np.abs(m["h"][same] - r["h"][same])
    assert dh.max() <= 5e-4 and (dh > 1e-4).sum() <= max(3, same.sum() // 50), (dh.max(), int((dh > 1e-4).sum()))
    # the known answer: a uniform (+12, +12) px shift, minus the reference's -1 px template-centre bias
    c2, r2 = n2.transform_points(m["lon2"][ok], m["lat2"][ok], 1)
    assert np.median(np.abs(c2 - (gx[ok] + 12 - 1))) < 0.75 and np.median(np.abs(r2 - (gy[ok] + 12 - 1))) < 0.75
