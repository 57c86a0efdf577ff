"""TEST INFRASTRUCTURE ONLY -- golden vectors for the deformation step from the UNMODIFIED reference
(``/root/reference/sea_ice_drift/libdefor.py``), run in the build container.  Writes tests/golden/defor.npz.

    python oracle/make_golden_defor.py

The reference's get_deformation_nodes needs matplotlib's Triangulation (absent here); its two arithmetic
functions are pure NumPy and run unmodified on a scipy Delaunay triangulation stored in the fixture."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import ref_import  # noqa: E402


def field(seed, n_side, spacing):
    """Jittered grid of nodes (metres) with a smooth drift field (m/s) plus noise."""
    rng = np.random.default_rng(seed)
    gx, gy = np.meshgrid(np.arange(n_side) * spacing, np.arange(n_side) * spacing)
    x = (gx + rng.uniform(-0.3, 0.3, gx.shape) * spacing).ravel() - 1.5e5
    y = (gy + rng.uniform(-0.3, 0.3, gy.shape) * spacing).ravel() + 2.0e5
    L = n_side * spacing
    u = 0.05 * np.sin(2 * np.pi * x / L) + 0.02 * y / L + rng.normal(0, 2e-3, x.size)
    v = 0.04 * np.cos(2 * np.pi * y / L) - 0.03 * x / L + rng.normal(0, 2e-3, x.size)
    return x, y, u, v


def main():
    ref_import.load_reference()
    import sea_ice_drift.libdefor as ref
    from scipy.spatial import Delaunay
    out = {}
    for k, (seed, n_side, spacing) in enumerate([(0, 40, 4000.0), (1, 25, 800.0), (2, 3, 10000.0)]):
        x, y, u, v = field(seed, n_side, spacing)
        tri = Delaunay(np.column_stack([x, y])).simplices.astype(np.int32)
        e1, e2, e3, a, p = ref.get_deformation_on_triangulation(x, y, u, v, tri)
        xt, yt, ut, vt = [q[tri].T for q in (x, y, u, v)]
        a_user = a * np.random.default_rng(100 + seed).uniform(0.5, 2.0, a.size)      # caller-supplied areas
        f1, f2, f3 = ref.get_deformation_elems(xt, yt, ut, vt, a_user)
        for name, val in dict(x=x, y=y, u=u, v=v, tri=tri, e1=e1, e2=e2, e3=e3, area=a, perim=p,
                              a_user=a_user, f1=f1, f2=f2, f3=f3).items():
            out["c%d_%s" % (k, name)] = val
    path = os.path.join(ROOT, "tests", "golden", "defor.npz")
    np.savez_compressed(path, **out)
    print("wrote", path, os.path.getsize(path), "bytes")


if __name__ == "__main__":
    main()
