"""The C-ABI library: builds for sm_100a, loads, exports every symbol that
include/sid_b200.h declares, and fails loudly without a GPU.  CPU only."""
import ctypes
import os
import re

import pytest

from sea_ice_drift_b200 import _lib, _build

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_symbols():
    text = open(os.path.join(ROOT, "include", "sid_b200.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(sid_[a-z_0-9]+)\s*\(", text)))


def test_header_symbols_all_exported():
    lib = _lib.load_library()
    names = declared_symbols()
    assert len(names) >= 14
    for name in names:
        assert hasattr(lib, name), "missing export " + name
    assert sorted(_lib.EXPORTS) == names


def test_library_is_sm100a_in_tree():
    assert os.path.dirname(_build.LIB) == os.path.join(ROOT, "sea_ice_drift_b200")
    assert "arch=compute_100a,code=sm_100a" in " ".join(_build.FLAGS)
    assert b"sm_100a" in _lib.load_library().sid_version()


def test_no_cpu_fallback_without_device():
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    with pytest.raises(_lib.SidError):
        _lib.Context(0)
    import numpy as np
    import sea_ice_drift_b200 as sid
    img = np.ones((64, 64), np.uint8)
    with pytest.raises(_lib.SidError):
        sid.use_mcc_batch([30.], [30.], [30.], [30.], [5.], img, img, 9, 0.0)


def test_product_never_imports_oracle():
    pkg = os.path.join(ROOT, "sea_ice_drift_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                text = open(os.path.join(dirpath, f)).read()
                assert "oracle" not in text.replace("no CPU fallback", ""), f


def test_null_and_bad_arguments_return_codes():
    lib = _lib.load_library()
    assert lib.sid_create(None, 0) == -1
    assert lib.sid_last_error(None) == b"null context"
    assert lib.sid_launch_count(None) == 0
    assert lib.sid_synchronize(None) == -1
    assert lib.sid_set_stream(None, None) == -1
