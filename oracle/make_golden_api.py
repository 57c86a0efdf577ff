"""TEST INFRASTRUCTURE ONLY -- records the public signatures of the UNMODIFIED reference for the functions this
package mirrors (tests/golden/api_signatures.json), so the drop-in check also runs where /root/reference is absent.

    python oracle/make_golden_api.py
"""
import inspect
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import ref_import  # noqa: E402

MIRRORED = {
    "pmlib": ["get_hessian", "get_distance_to_nearest_keypoint", "get_initial_rotation", "get_template", "rotate_and_match",
              "use_mcc", "use_mcc_mp", "prepare_first_guess", "pattern_matching"],
    "lib": ["interpolation_poly", "interpolation_near", "_fill_gpi"],
    "ftlib": ["get_match_coords"],
    "libdefor": ["get_deformation_elems", "get_deformation_on_triangulation", "get_deformation_nodes"],
}


def describe(fn):
    out = []
    for p in inspect.signature(fn).parameters.values():
        d = p.default
        if d is inspect.Parameter.empty:
            default = None
        elif callable(d):
            default = "<callable>"              # plug-in defaults (cv2.matchTemplate, cv2.BFMatcher) are replaced on purpose
        else:
            default = repr(d)
        out.append([p.name, str(p.kind), default])
    return out


def main():
    import importlib
    ref_import.load_reference()
    table = {}
    for mod, names in MIRRORED.items():
        m = importlib.import_module("sea_ice_drift." + mod)
        for n in names:
            table["%s.%s" % (mod, n)] = describe(getattr(m, n))
    import sea_ice_drift.seaicedrift as rs
    for n in ("__init__", "get_drift_FT", "get_drift_PM"):
        table["seaicedrift.SeaIceDrift.%s" % n] = describe(getattr(rs.SeaIceDrift, n))
    path = os.path.join(ROOT, "tests", "golden", "api_signatures.json")
    json.dump(table, open(path, "w"), indent=1, sort_keys=True)
    print("wrote", path, len(table), "signatures")


if __name__ == "__main__":
    main()
