"""Feature-tracking matcher: cv2.BFMatcher(NORM_HAMMING).knnMatch(k=2) as the reference
calls it (ftlib.py:92-99).  CPU part pins the oracle's ordering rule against OpenCV; GPU part checks the kernel."""
import numpy as np
import pytest
import cv2

from oracle import c_oracle as co


def descriptors(rng, n, dup=0):
    if dup:
        base = rng.integers(0, 256, (dup, 32), dtype=np.uint8)
        d = base[rng.integers(0, dup, n)].copy()
        flip = rng.random(n) < 0.3                       # some near-duplicates: distance-1 ties
        d[flip, rng.integers(0, 32, flip.sum())] ^= 1
        return d
    return rng.integers(0, 256, (n, 32), dtype=np.uint8)


def cv2_knn(d1, d2):
    m = cv2.BFMatcher(cv2.NORM_HAMMING).knnMatch(d1, d2, k=2)
    idx = np.full((len(d1), 2), -1, np.int32)
    dist = np.full((len(d1), 2), -1, np.int32)
    for q, pair in enumerate(m):
        for k, x in enumerate(pair):
            idx[q, k], dist[q, k] = x.trainIdx, int(x.distance)
    return idx, dist


CASES = [(1, 2, 0), (5, 1, 0), (130, 129, 0), (300, 257, 7), (257, 1000, 3), (1000, 2500, 40), (64, 5000, 0)]


@pytest.mark.parametrize("n1,n2,dup", CASES)
def test_oracle_matches_opencv_ordering(n1, n2, dup):
    rng = np.random.default_rng(n1 * 7 + n2)
    d1, d2 = descriptors(rng, n1, dup), descriptors(rng, n2, dup)
    idx, dist = co.knn_hamming2(d1, d2)
    ridx, rdist = cv2_knn(d1, d2)
    assert np.array_equal(idx, ridx) and np.array_equal(dist, rdist)


@pytest.mark.gpu
@pytest.mark.parametrize("n1,n2,dup", CASES + [(3000, 7000, 200), (20000, 3000, 0)])
def test_gpu_matcher_equals_opencv_and_oracle(gpu_ctx, n1, n2, dup):
    rng = np.random.default_rng(n1 * 7 + n2)
    d1, d2 = descriptors(rng, n1, dup), descriptors(rng, n2, dup)
    idx, dist = gpu_ctx.knn_hamming2(d1, d2)
    ridx, rdist = cv2_knn(d1, d2)
    assert np.array_equal(idx, ridx) and np.array_equal(dist, rdist)
    if n1 * n2 <= 3_000_000:
        oidx, odist = co.knn_hamming2(d1, d2)
        assert np.array_equal(idx, oidx) and np.array_equal(dist, odist)


@pytest.mark.gpu
def test_gpu_matcher_as_reference_plugin():
    """The class plugs into get_match_coords exactly like cv2.BFMatcher (ftlib.py:64-116)."""
    from sea_ice_drift_b200 import ftlib
    rng = np.random.default_rng(5)
    d1 = descriptors(rng, 800)
    d2 = descriptors(rng, 900)
    src = rng.permutation(800)[:500]                     # 500 true correspondences: 10 flipped bits each
    d2[:500] = d1[src]
    for row in range(500):
        for bit in rng.choice(256, 10, replace=False):
            d2[row, bit // 8] ^= np.uint8(1 << (bit % 8))
    kp1 = [cv2.KeyPoint(float(x), float(y), 1) for x, y in rng.uniform(0, 500, (800, 2))]
    kp2 = [cv2.KeyPoint(float(x), float(y), 1) for x, y in rng.uniform(0, 500, (900, 2))]
    fast = ftlib.get_match_coords(kp1, d1, kp2, d2, ratio_test=0.75)
    via_cv2 = ftlib.get_match_coords(kp1, d1, kp2, d2, matcher=cv2.BFMatcher, norm=cv2.NORM_HAMMING, ratio_test=0.75)
    assert len(fast[0]) > 10
    for a, b in zip(fast, via_cv2):
        assert np.array_equal(a, b)
    # object interface: same attributes the reference reads from cv2.DMatch
    m = ftlib.BFMatcher(ftlib.NORM_HAMMING).knnMatch(d1[:5], d2, k=2)
    ref = cv2.BFMatcher(cv2.NORM_HAMMING).knnMatch(d1[:5], d2, k=2)
    for (a0, a1), (b0, b1) in zip(m, ref):
        assert (a0.queryIdx, a0.trainIdx, a0.distance, a1.trainIdx, a1.distance) == \
               (b0.queryIdx, b0.trainIdx, b0.distance, b1.trainIdx, b1.distance)
    with pytest.raises(ValueError):
        ftlib.BFMatcher(4)
