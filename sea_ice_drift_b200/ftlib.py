"""GPU matcher for the reference's feature tracking (the caller side of the hot path).

``BFMatcher`` is a drop-in for the ``matcher`` plug-in kwarg of the reference's
``get_match_coords`` / ``feature_tracking`` (reference sea_ice_drift/ftlib.py:64-99):

    feature_tracking(n1, n2, matcher=sea_ice_drift_b200.ftlib.BFMatcher, ...)

ORB keypoint detection itself stays on OpenCV, as BASELINE.json's north_star says."""
import numpy as np

from . import _lib

NORM_HAMMING = 6            # == cv2.NORM_HAMMING


class DMatch(object):
    """The four attributes of cv2.DMatch the reference reads (ftlib.py:105-116)."""
    __slots__ = ("queryIdx", "trainIdx", "imgIdx", "distance")

    def __init__(self, queryIdx, trainIdx, distance):
        self.queryIdx, self.trainIdx, self.imgIdx, self.distance = queryIdx, trainIdx, 0, distance


class BFMatcher(object):
    """``cv2.BFMatcher(cv2.NORM_HAMMING)`` look-alike whose ``knnMatch(d1, d2, k=2)`` runs on the GPU."""

    def __init__(self, normType=NORM_HAMMING, crossCheck=False, device=None):
        if normType != NORM_HAMMING or crossCheck:
            raise ValueError("only NORM_HAMMING without cross-check is implemented")
        self._device = device

    def knn_arrays(self, queryDescriptors, trainDescriptors):
        return _lib.default_context(self._device).knn_hamming2(queryDescriptors, trainDescriptors)

    def knnMatch(self, queryDescriptors, trainDescriptors, k=2):
        if k != 2:
            raise ValueError("only k=2 is implemented")
        idx, dist = self.knn_arrays(queryDescriptors, trainDescriptors)
        return [[DMatch(q, int(i), float(d)) for i, d in zip(idx[q], dist[q]) if i >= 0] for q in range(len(idx))]


def _points(keypoints):
    if isinstance(keypoints, np.ndarray):
        return np.asarray(keypoints, dtype=np.float64).reshape(-1, 2)
    return np.array([kp.pt for kp in keypoints], dtype=np.float64).reshape(-1, 2)


def get_match_coords(keyPoints1, descriptors1, keyPoints2, descriptors2, matcher=BFMatcher, norm=NORM_HAMMING,
                     ratio_test=0.7, verbose=False, **kwargs):
    """Lowe-ratio-filtered matches as start / end coordinates ``x1, y1, x2, y2`` (reference ftlib.py:64-116).
    With the GPU matcher the ratio test runs on the (n, 2) arrays directly; any other ``matcher`` class is
    used through its ``knnMatch`` exactly as the reference does."""
    bf = matcher(norm)
    if isinstance(bf, BFMatcher):
        idx, dist = bf.knn_arrays(descriptors1, descriptors2)
        good = (idx[:, 1] >= 0) & (dist[:, 0].astype(np.float64) < ratio_test * dist[:, 1].astype(np.float64))
        q, t = np.nonzero(good)[0], idx[good, 0]
    else:
        pairs = [(m.queryIdx, m.trainIdx) for m, n in bf.knnMatch(descriptors1, descriptors2, k=2)
                 if m.distance < ratio_test * n.distance]
        q = np.array([p[0] for p in pairs], dtype=np.int64)
        t = np.array([p[1] for p in pairs], dtype=np.int64)
    if verbose:
        print('Ratio test %f found %d keypoints' % (ratio_test, len(q)))
    p1, p2 = _points(keyPoints1), _points(keyPoints2)
    return p1[q, 0], p1[q, 1], p2[t, 0], p2[t, 1]
