"""Deformation step: oracle pinned by the reference's own outputs (tests/golden/defor.npz),
GPU path (sid_deformation through libdefor) against fixture and oracle."""
import os

import numpy as np
import pytest

from oracle import defor_oracle

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "defor.npz")
# FP64 tolerance of the GPU path: every operation but hypot() is the same IEEE operation in the same order as
# NumPy's; the side lengths may differ by 1 ulp, which the near-cancelling sums of e1/e3 amplify -- measured
# against the scale of the velocity gradients (max |e2|).
RTOL = 1e-12


def cases():
    g = np.load(GOLD)
    n = len({k.split("_")[0] for k in g.files})
    return [{k.split("_", 1)[1]: g[k] for k in g.files if k.startswith("c%d_" % i)} for i in range(n)]


def test_oracle_equals_reference_outputs_bit_for_bit():
    for c in cases():
        e1, e2, e3, a, p = defor_oracle.deformation(c["x"], c["y"], c["u"], c["v"], c["tri"])
        for got, key in ((e1, "e1"), (e2, "e2"), (e3, "e3"), (a, "area"), (p, "perim")):
            assert np.array_equal(got, c[key]), key
        f1, f2, f3, _, _ = defor_oracle.deformation(c["x"], c["y"], c["u"], c["v"], c["tri"], area=c["a_user"])
        assert np.array_equal(f1, c["f1"]) and np.array_equal(f2, c["f2"]) and np.array_equal(f3, c["f3"])


def test_oracle_known_answers():
    # rigid rotation u = -w y, v = w x: no divergence, no shear, vorticity 2w; pure dilation u = k x, v = k y
    x = np.array([0.0, 1000.0, 0.0, 1000.0]); y = np.array([0.0, 0.0, 1000.0, 1000.0])
    tri = np.array([[0, 1, 2], [1, 3, 2]])
    w, k = 1e-6, 3e-7
    e1, e2, e3, a, p = defor_oracle.deformation(x, y, -w * y, w * x, tri)
    assert np.allclose(e1, 0, atol=1e-18) and np.allclose(e2, 0, atol=1e-18) and np.allclose(e3, 2 * w)
    assert np.allclose(a, 5e5) and np.allclose(p, 2000 + 1000 * np.sqrt(2))
    e1, e2, e3, _, _ = defor_oracle.deformation(x, y, k * x, k * y, tri)
    assert np.allclose(e1, 2 * k) and np.allclose(e2, 0, atol=1e-18) and np.allclose(e3, 0, atol=1e-18)


def test_quality_mask_and_triangulation_host_side():
    from sea_ice_drift_b200 import libdefor
    r = np.array([0.5, 0.9, np.nan, 0.8]); h = np.array([5.0, 6.0, 7.0, np.nan])
    assert libdefor.quality_mask(r, h).tolist() == [False, True, False, False]
    c = cases()[1]
    t = libdefor.triangulate(c["x"], c["y"])
    as_set = lambda tt: {tuple(sorted(row)) for row in tt.tolist()}
    assert as_set(t) == as_set(c["tri"])
    x, y = c["x"], c["y"]
    cross = (x[t[:, 1]] - x[t[:, 0]]) * (y[t[:, 2]] - y[t[:, 0]]) - (x[t[:, 2]] - x[t[:, 0]]) * (y[t[:, 1]] - y[t[:, 0]])
    assert (cross > 0).all()


def _close(got, want, scale):
    return np.max(np.abs(got - want)) <= RTOL * scale


@pytest.mark.gpu
def test_gpu_deformation_equals_reference_fixture():
    from sea_ice_drift_b200 import libdefor
    for c in cases():
        e1, e2, e3, a, p = libdefor.get_deformation_on_triangulation(c["x"], c["y"], c["u"], c["v"], c["tri"])
        scale = np.abs(c["e2"]).max()
        assert _close(e1, c["e1"], scale) and _close(e2, c["e2"], scale) and _close(e3, c["e3"], scale)
        assert np.max(np.abs(a / c["area"] - 1)) <= RTOL and np.max(np.abs(p / c["perim"] - 1)) <= RTOL
        xt, yt, ut, vt = [q[c["tri"]].T for q in (c["x"], c["y"], c["u"], c["v"])]
        f1, f2, f3 = libdefor.get_deformation_elems(xt, yt, ut, vt, c["a_user"])
        # caller-supplied areas: no hypot on the path, so these are bit-identical
        assert np.array_equal(f1, c["f1"]) and np.array_equal(f2, c["f2"]) and np.array_equal(f3, c["f3"])


@pytest.mark.gpu
def test_gpu_deformation_seeded_large_and_edge_cases():
    from sea_ice_drift_b200 import libdefor, _lib
    rng = np.random.default_rng(7)
    n = 90000
    x = rng.uniform(-4e5, 4e5, n); y = rng.uniform(-4e5, 4e5, n)
    u = rng.normal(0, 0.1, n); v = rng.normal(0, 0.1, n)
    e1, e2, e3, a, p, t = libdefor.get_deformation_nodes(x, y, u, v)
    w1, w2, w3, wa, wp = defor_oracle.deformation(x, y, u, v, t)
    identical = int(np.sum((a == wa) & (p == wp)))
    print("elements %d, area+perimeter bit-identical to NumPy: %d" % (len(t), identical))
    # the side lengths use a correctly rounded hypot (verified against exact rational arithmetic); glibc's, which
    # NumPy calls, is correctly rounded for ~99.3 % of arguments, so ~1.5 % of the elements differ in the last bits
    assert identical >= 0.97 * len(t)
    assert np.max(np.abs(a / wa - 1)) <= RTOL and np.max(np.abs(p / wp - 1)) <= RTOL
    # slivers amplify a 1-ulp side length: compare against the gradient scale of each element
    scale = np.maximum(np.abs(w2), 1e-30) * np.maximum(1.0, (wp * wp) / wa)
    for got, want in ((e1, w1), (e2, w2), (e3, w3)):
        assert np.all(np.abs(got - want) <= 1e-11 * scale)
    # empty input, bad index -> NaN element, degenerate (collinear) element -> division by zero area like NumPy
    ctx = _lib.default_context()
    assert all(o.size == 0 for o in ctx.deformation(x[:3], y[:3], u[:3], v[:3], np.zeros((0, 3), np.int32)))
    out = ctx.deformation(x[:3], y[:3], u[:3], v[:3], np.array([[0, 1, 7], [0, 1, 2]], np.int32))
    assert all(np.isnan(o[0]) and np.isfinite(o[1]) for o in out)
    xs = np.array([0.0, 1.0, 2.0]); zs = np.zeros(3)
    with np.errstate(all='ignore'):
        want = defor_oracle.deformation(xs, zs, xs, zs, np.array([[0, 1, 2]]))
    got = ctx.deformation(xs, zs, xs, zs, np.array([[0, 1, 2]], np.int32))
    for g_, w_ in zip(got, want):
        assert np.array_equal(g_, w_, equal_nan=True)
    with pytest.raises(ValueError):
        ctx.deformation(x[:3], y[:3], u[:3], v[:2], np.array([[0, 1, 2]], np.int32))
