"""Build the plain-C oracle (test infrastructure) into oracle/_build/libsid_oracle.so."""
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
OUT_DIR = os.path.join(HERE, "_build")
LIB = os.path.join(OUT_DIR, "libsid_oracle.so")


def build(force=False):
    src = os.path.join(HERE, "mcc_oracle.c")
    if (not force and os.path.exists(LIB)
            and os.path.getmtime(LIB) >= os.path.getmtime(src)):
        return LIB
    os.makedirs(OUT_DIR, exist_ok=True)
    cmd = ["gcc", "-O2", "-fPIC", "-shared", "-std=c11",
           "-ffp-contract=off", "-fno-fast-math", "-o", LIB, src, "-lm"]
    subprocess.check_call(cmd)
    return LIB


if __name__ == "__main__":
    print(build(force="--force" in sys.argv))
