"""Hamming kNN matcher: GPU vs cv2.BFMatcher on the host, 20 000 x 20 000 ORB-sized descriptors
(the size in the reference's notebooks: 2.80 s, examples/drift_from_arrays.ipynb:144)."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, cv2
from sea_ice_drift_b200 import _lib
rng = np.random.default_rng(0)
n = int(sys.argv[1]) if len(sys.argv) > 1 else 20000
d1 = rng.integers(0, 256, (n, 32), dtype=np.uint8); d2 = rng.integers(0, 256, (n, 32), dtype=np.uint8)
ctx = _lib.Context(0)
ctx.knn_hamming2(d1[:256], d2[:256])
t = time.perf_counter(); idx, dist = ctx.knn_hamming2(d1, d2); t_gpu = time.perf_counter() - t
t = time.perf_counter(); idx, dist = ctx.knn_hamming2(d1, d2); t_gpu2 = time.perf_counter() - t
cv2.setNumThreads(os.cpu_count())
t = time.perf_counter(); m = cv2.BFMatcher(cv2.NORM_HAMMING).knnMatch(d1, d2, k=2); t_cv = time.perf_counter() - t
ridx = np.array([[a.trainIdx, b.trainIdx] for a, b in m]); rdist = np.array([[a.distance, b.distance] for a, b in m])
print("n=%d: GPU (host arrays in/out) %.2f ms (first %.2f ms), cv2.BFMatcher %d threads %.0f ms, speed-up %.0fx, identical=%s" % (
    n, t_gpu2 * 1e3, t_gpu * 1e3, cv2.getNumThreads(), t_cv * 1e3, t_cv / t_gpu2, np.array_equal(idx, ridx) and np.array_equal(dist, rdist)))
