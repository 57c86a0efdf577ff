"""TEST INFRASTRUCTURE ONLY -- CPU restatement (NumPy, float64) of the reference's element deformation,
``sea_ice_drift/libdefor.py``: get_deformation_on_triangulation (:50-99) and get_deformation_elems (:4-48).

Pinned by ``tests/golden/defor.npz`` (outputs of the unmodified reference functions on seeded inputs, generated
by ``oracle/make_golden_defor.py``); only tests may import it.  Written element-major (M x 3) with the same order
of floating-point operations as the reference's node-major code, so results are bit-identical."""
import numpy as np


def deformation(x, y, u, v, tri, area=None):
    """(e1, e2, e3, area, perimeter) per element.  tri: (M, 3) node indices."""
    x, y, u, v = [np.asarray(k, dtype=np.float64).ravel() for k in (x, y, u, v)]
    tri = np.asarray(tri, dtype=np.int64).reshape(-1, 3)
    X, Y, U, V = x[tri], y[tri], u[tri], v[tri]                    # (M, 3)
    nxt = [1, 2, 0]
    # sides node c -> node c+1 (reference: np.diff over the closed polygon, libdefor.py:86-88)
    side = np.hypot(X[:, nxt] - X, Y[:, nxt] - Y)
    perim = (side[:, 0] + side[:, 1]) + side[:, 2]                # libdefor.py:90
    s = perim / 2
    heron = np.sqrt(s * (s - side[:, 0]) * (s - side[:, 1]) * (s - side[:, 2]))   # libdefor.py:93
    a = heron if area is None else np.asarray(area, dtype=np.float64).ravel()
    # contour integrals over the sides (1,0), (2,1), (0,2) in that order (libdefor.py:36-40)
    ux = uy = vx = vy = 0.0
    for i0, i1 in ((1, 0), (2, 1), (0, 2)):
        us, vs = U[:, i0] + U[:, i1], V[:, i0] + V[:, i1]
        dy, dx = Y[:, i0] - Y[:, i1], X[:, i0] - X[:, i1]
        ux = ux + us * dy
        uy = uy - us * dx
        vx = vx + vs * dy
        vy = vy - vs * dx
    a2 = 2 * a
    ux, uy, vx, vy = ux / a2, uy / a2, vx / a2, vy / a2            # libdefor.py:42
    d1, d2 = ux - vy, uy + vx
    return ux + vy, np.sqrt(d1 * d1 + d2 * d2), vx - uy, a, perim  # libdefor.py:45-47
