"""Build recipe for the oracle's C restatement -- test infrastructure.

    python -m oracle.build        ->  oracle/_build/libofo.so

``-fno-builtin`` keeps gcc from folding ``pow(x, 2.0)`` / ``sqrt`` so libm is
called exactly as CPython calls it; ``-ffp-contract=off`` forbids FMA fusion.
The reference itself is pure Python (nothing to compile), so there is no
``oracle/_ref`` binary: the "compiled reference" leg of the task does not apply.
"""
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
OUT_DIR = os.path.join(HERE, "_build")
LIB = os.path.join(OUT_DIR, "libofo.so")


def build(force=False):
    src = os.path.join(HERE, "step_c.c")
    if (not force and os.path.exists(LIB)
            and os.path.getmtime(LIB) >= os.path.getmtime(src)):
        return LIB
    os.makedirs(OUT_DIR, exist_ok=True)
    cmd = ["gcc", "-O2", "-std=gnu11", "-fPIC", "-shared", "-fno-builtin", "-ffp-contract=off",
           "-fopenmp", "-o", LIB, src, "-lm"]
    subprocess.check_call(cmd)
    return LIB


if __name__ == "__main__":
    print(build(force="--force" in sys.argv))
