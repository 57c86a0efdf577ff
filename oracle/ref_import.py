"""TEST INFRASTRUCTURE ONLY -- loads the *unmodified* reference hot path.

Imports ``/root/reference/sea_ice_drift/pmlib.py`` (and ``lib.py``) after
registering empty stand-ins for the modules the reference imports at module
top but never touches on the pattern-matching path (``matplotlib``,
``osgeo.gdal``, ``nansat``; see reference ``pmlib.py:22-26``, ``lib.py:16-23``).

Source of the reference files, in this order: ``/root/reference`` (the build container) or the byte-for-byte
copy that ``oracle/build_ref.py`` placed under the git-ignored ``oracle/_ref/`` (which travels to the GPU box with
the snapshot).  It is used to (1) pin the restatements in ``oracle/``, (2) generate the golden vectors committed
under ``tests/golden/`` (``oracle/make_golden.py``), (3) run the REAL reference beside the CUDA path in the
``-m gpu`` parity tests and in ``bench.py``'s CPU legs.  Nothing in the product imports this file.
"""
import os
import sys
import types
import warnings

_HERE = os.path.dirname(os.path.abspath(__file__))
REFERENCE_ROOT = os.environ.get("SID_REFERENCE_ROOT", "/root/reference")
if not os.path.isfile(os.path.join(REFERENCE_ROOT, "sea_ice_drift", "pmlib.py")):
    REFERENCE_ROOT = os.path.join(_HERE, "_ref")          # the copy placed by oracle/build_ref.py


def reference_available():
    return os.path.isfile(os.path.join(REFERENCE_ROOT, "sea_ice_drift", "pmlib.py"))


def reference_kind():
    """'reference' when the unmodified reference can be loaded (here or on the GPU box), else None."""
    return "reference" if reference_available() else None


def _stub(name, **attrs):
    if name in sys.modules:
        return sys.modules[name]
    m = types.ModuleType(name)
    m.__dict__.update(attrs)
    sys.modules[name] = m
    return m


def load_reference():
    """Return the reference ``sea_ice_drift.pmlib`` module (unmodified source)."""
    if not reference_available():
        raise ImportError("reference tree not present at %s" % REFERENCE_ROOT)

    class _Dummy(object):
        def __init__(self, *a, **k):
            pass

    mpl = _stub("matplotlib")
    mpl.pyplot = _stub("matplotlib.pyplot")
    mpl.tri = _stub("matplotlib.tri", Triangulation=_Dummy)
    osgeo = _stub("osgeo")
    osgeo.gdal = _stub("osgeo.gdal")
    _stub("nansat", Nansat=_Dummy, Domain=_Dummy, NSR=_Dummy)
    if REFERENCE_ROOT not in sys.path:
        sys.path.insert(0, REFERENCE_ROOT)
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        import sea_ice_drift.pmlib as ref_pmlib
    return ref_pmlib
