"""TEST INFRASTRUCTURE ONLY -- recipe that places the UNMODIFIED reference next to the oracle.

The reference (nansencenter/sea_ice_drift) is pure Python, so "building" it for the GPU box means making its
source files travel: ``oracle/_ref/`` is git-ignored (the repository never holds reference sources) but NOT
gpurun-ignored, so whatever this recipe puts there rides along with the snapshot exactly like our own built
``.so`` files.  ``build()`` copies ``/root/reference/sea_ice_drift/*.py`` byte for byte into
``oracle/_ref/sea_ice_drift/`` and writes ``MANIFEST.json`` with the sha256 of every file, so a reader can check
that what ran on the GPU box is the unmodified reference.  Where ``/root/reference`` is absent (the GPU box) the
recipe does nothing and the previously placed copy is used as is.

Used by: ``oracle/ref_import.py`` (loads the reference with matplotlib / osgeo / nansat stubbed),
``bench.py --impl reference`` and the ``cpu_baseline`` leg (``kind: "reference"``), and the ``-m gpu`` parity tests
that compare the CUDA path with the reference's own ``pmlib.use_mcc_mp`` at BASELINE.json's sizes.
Nothing under ``sea_ice_drift_b200/`` imports it.
"""
import hashlib
import json
import os
import shutil

HERE = os.path.dirname(os.path.abspath(__file__))
REF_SRC = os.environ.get("SID_REFERENCE_ROOT", "/root/reference")
REF_DST = os.path.join(HERE, "_ref")
FILES = ("__init__.py", "pmlib.py", "lib.py", "ftlib.py", "libdefor.py", "seaicedrift.py")


def _sha(path):
    with open(path, "rb") as f:
        return hashlib.sha256(f.read()).hexdigest()


def build():
    """Copy the reference package (unmodified) into oracle/_ref; returns the destination or None."""
    src_pkg = os.path.join(REF_SRC, "sea_ice_drift")
    dst_pkg = os.path.join(REF_DST, "sea_ice_drift")
    if not os.path.isfile(os.path.join(src_pkg, "pmlib.py")):
        return REF_DST if os.path.isfile(os.path.join(dst_pkg, "pmlib.py")) else None
    os.makedirs(dst_pkg, exist_ok=True)
    manifest = {}
    for name in FILES:
        src = os.path.join(src_pkg, name)
        if not os.path.isfile(src):
            continue
        dst = os.path.join(dst_pkg, name)
        if not os.path.isfile(dst) or _sha(dst) != _sha(src):
            shutil.copyfile(src, dst)
        manifest[name] = _sha(dst)
    with open(os.path.join(REF_DST, "MANIFEST.json"), "w") as f:
        json.dump({"source": "nansencenter/sea_ice_drift (unmodified files of /root/reference/sea_ice_drift)",
                   "sha256": manifest}, f, indent=1, sort_keys=True)
    return REF_DST


def available():
    return os.path.isfile(os.path.join(REF_DST, "sea_ice_drift", "pmlib.py"))


if __name__ == "__main__":
    print(build())
