"""Compile the CUDA library IN-TREE for sm_100a: sea_ice_drift_b200/libsid_b200.so.

The shared object is git-ignored but travels to the GPU box with the snapshot.
"""
import os
import subprocess
import sys

PKG = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(PKG, "csrc")
LIB = os.path.join(PKG, "libsid_b200.so")
NVCC = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-std=c++17", "-lineinfo",
         "-Xcompiler", "-fPIC", "-shared", "-cudart", "static"]


def sources():
    return [os.path.join(CSRC, f) for f in sorted(os.listdir(CSRC)) if f.endswith((".cu", ".cuh"))] + \
           [os.path.join(os.path.dirname(PKG), "include", "sid_b200.h")]


def is_stale():
    if not os.path.exists(LIB):
        return True
    t = os.path.getmtime(LIB)
    return any(os.path.getmtime(s) > t for s in sources())


def build(force=False, verbose=False, out=None, extra=()):
    """``out`` / ``extra``: build an experimental variant next to the product library, e.g.
    ``build(force=True, out='/tmp/libsid_variant.so', extra=['-DSOME_SWITCH'])`` for an A/B run."""
    if not force and not is_stale() and out is None:
        return LIB
    target = out or LIB
    # several ranks of one torchrun may find the library stale at the same moment: one builds (file lock), the others
    # wait and re-check; the compiler writes to a temporary name and the result is moved into place atomically, so a
    # concurrent loader never sees a half-written file
    import fcntl
    with open(target + ".lock", "w") as lock:
        fcntl.flock(lock, fcntl.LOCK_EX)
        if not force and out is None and not is_stale():
            return LIB
        tmp = "%s.tmp.%d" % (target, os.getpid())
        cmd = [NVCC] + FLAGS + list(extra) + (["-Xptxas", "-v"] if verbose else []) + \
              ["-o", tmp, os.path.join(CSRC, "sid_api.cu")]
        res = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
        if verbose or res.returncode:
            sys.stderr.write(res.stdout)
        if res.returncode:
            if os.path.exists(tmp):
                os.remove(tmp)
            raise RuntimeError("nvcc failed (%d)" % res.returncode)
        os.replace(tmp, target)
    return target


if __name__ == "__main__":
    print(build(force=True, verbose="-v" in sys.argv))
