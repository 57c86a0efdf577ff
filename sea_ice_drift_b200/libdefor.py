"""Deformation of the drift field and the quality filter that precedes it -- the immediate consumer of the
pattern-matching output.  Same functions and argument meaning as the reference's
``sea_ice_drift/libdefor.py``; the arithmetic runs in ``sid_deformation`` on the GPU (one thread per element).

The triangulation of ``get_deformation_nodes`` comes from ``scipy.spatial.Delaunay`` (the reference uses
``matplotlib.tri.Triangulation``, libdefor.py:133, which is not a dependency here): both are Qhull Delaunay
triangulations, so the element SET is the same for points in general position; element order may differ and
rows are oriented counter-clockwise like matplotlib's.
"""
import numpy as np

from . import _lib


def quality_mask(r, h, threshold=4.0):
    """High-quality vectors as selected in the reference's README.md:79: ``rpm * hpm > 4`` (NaN -> False)."""
    r = np.asarray(r, dtype=np.float64)
    h = np.asarray(h, dtype=np.float64)
    with np.errstate(invalid='ignore'):
        return np.nan_to_num(r * h, nan=-np.inf) > threshold


def get_deformation_elems(x, y, u, v, a):
    """Divergence, shear and vorticity (1/s) of M elements whose node values are given as 3 x M arrays and whose
    areas are ``a`` (reference libdefor.py:4-48)."""
    x, y, u, v = [np.asarray(k, dtype=np.float64) for k in (x, y, u, v)]
    if x.ndim != 2 or x.shape[0] != 3:
        raise ValueError("node values must be 3 x M arrays")
    m = x.shape[1]
    tri = np.arange(3 * m, dtype=np.int32).reshape(3, m).T        # element e uses flat nodes e, M+e, 2M+e
    e1, e2, e3, _, _ = _lib.default_context().deformation(x, y, u, v, tri, area=a)
    return e1, e2, e3


def get_deformation_on_triangulation(x, y, u, v, t):
    """(e1, e2, e3, area, perimeter) for the elements ``t`` (M x 3 node indices) of N nodes
    (reference libdefor.py:50-99)."""
    return _lib.default_context().deformation(x, y, u, v, t)


def triangulate(x, y):
    """Delaunay triangulation of the nodes: (M, 3) int32 indices, counter-clockwise rows."""
    from scipy.spatial import Delaunay
    x = np.ravel(np.asarray(x, dtype=np.float64))
    y = np.ravel(np.asarray(y, dtype=np.float64))
    t = Delaunay(np.column_stack([x, y])).simplices.astype(np.int32)
    cross = (x[t[:, 1]] - x[t[:, 0]]) * (y[t[:, 2]] - y[t[:, 0]]) - (x[t[:, 2]] - x[t[:, 0]]) * (y[t[:, 1]] - y[t[:, 0]])
    cw = cross < 0
    t[cw] = t[cw][:, [0, 2, 1]]
    return t


def get_deformation_nodes(x, y, u, v):
    """Triangulate the nodes and compute (e1, e2, e3, area, perimeter, triangles) (reference libdefor.py:101-137)."""
    t = triangulate(x, y)
    return get_deformation_on_triangulation(x, y, u, v, t) + (t,)
