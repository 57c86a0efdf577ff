"""Host-side helpers that `pattern_matching` needs either side of the GPU hot path.
They mirror the reference's helper names (reference sea_ice_drift/lib.py:139-201,
408-412) so that callers can switch imports; none of them is on the hot path."""
import numpy as np
from scipy.interpolate import griddata, LinearNDInterpolator


def _poly_design(x, y, order):
    cols = [np.ones(len(x)), x, y]
    if order > 1:
        cols += [x ** 2, y ** 2, x * y]
    if order > 2:
        cols += [x ** 3, y ** 3, x ** 2 * y, y ** 2 * x]
    return np.vstack(cols).T


def interpolation_poly(x1, y1, x2, y2, x1grd, y1grd, order=1, **kwargs):
    """Least-squares polynomial (order 1-3) map (x1, y1) -> (x2, y2), evaluated on
    the grid points (reference lib.py:139-177)."""
    design = _poly_design(np.asarray(x1), np.asarray(y1), order)
    coef_x = np.linalg.lstsq(design, x2, rcond=-1)[0]
    coef_y = np.linalg.lstsq(design, y2, rcond=-1)[0]
    gx, gy = np.asarray(x1grd), np.asarray(y1grd)
    grid_design = _poly_design(gx.flatten(), gy.flatten(), order)
    return (np.dot(grid_design, coef_x).reshape(gx.shape),
            np.dot(grid_design, coef_y).reshape(gx.shape))


def interpolation_near(x1, y1, x2, y2, x1grd, y1grd, method='linear', **kwargs):
    """Piecewise (Delaunay) interpolation of x2, y2 onto the grid points; NaN outside
    the convex hull of the keypoints (reference lib.py:179-201)."""
    src = np.array([y1, x1]).T
    dst = np.array([y1grd, x1grd]).T
    if method == 'linear' and src.ndim == 2 and src.shape[1] == 2:
        # what griddata(method='linear') does, but with ONE Delaunay triangulation shared by both value
        # sets (the reference triangulates the same keypoints twice); values are identical
        both = LinearNDInterpolator(src, np.column_stack([x2, y2]), fill_value=np.nan)(dst)
        return both[..., 0].T, both[..., 1].T      # value axis is the last one for 1-D and 2-D grids alike
    return (griddata(src, x2, dst, method=method).T,
            griddata(src, y2, dst, method=method).T)


def _fill_gpi(shape, gpi, data):
    """Scatter 1-D `data` (one value per True in `gpi`) into a NaN-filled array of
    `shape` (reference lib.py:408-412)."""
    full = np.full(int(np.prod(shape)), np.nan)
    full[gpi] = data
    return full.reshape(shape)
