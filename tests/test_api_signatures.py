"""Drop-in check: every mirrored function takes the reference's parameters (names, kinds, order, defaults).
Fixture: tests/golden/api_signatures.json, recorded from the unmodified reference by oracle/make_golden_api.py."""
import importlib
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle.make_golden_api import describe  # noqa: E402


def test_mirrored_signatures_equal_reference():
    table = json.load(open(os.path.join(ROOT, "tests", "golden", "api_signatures.json")))
    assert len(table) >= 19
    for key, want in table.items():
        parts = key.split(".")
        obj = importlib.import_module("sea_ice_drift_b200." + parts[0])
        for p in parts[1:]:
            obj = getattr(obj, p)
        got = describe(obj)
        assert [g[:2] for g in got] == [w[:2] for w in want], key
        for g, w in zip(got, want):
            if w[2] == "<callable>":
                continue            # plug-in default (cv2.matchTemplate / cv2.BFMatcher): ours selects the GPU implementation
            assert g[2] == w[2], (key, g, w)
