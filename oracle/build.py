"""Build the CPU oracle (test infrastructure) into oracle/_build/libdfjk_oracle.so.

The reference file as a whole (lib3index/dfhelper.cc) cannot be compiled here: it includes
libmints/basisset.h -> <libint2/shell.h>, and Libint2 is neither installed nor vendored
(SURVEY.md 8c).  The functions ON the J/K path never touch those headers, though: oracle/ref_build.py
slices them out of the reference file at build time and compiles them into oracle/_ref/libref_dfjk.so,
against which this restatement is checked bit for bit (tests/test_reference_slice.py).
"""
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
OUT = os.path.join(HERE, "_build")
LIB = os.path.join(OUT, "libdfjk_oracle.so")


def build(force: bool = False) -> str:
    src = os.path.join(HERE, "dfjk_oracle.c")
    os.makedirs(OUT, exist_ok=True)
    if not force and os.path.exists(LIB) and os.path.getmtime(LIB) >= os.path.getmtime(src):
        return LIB
    cmd = ["gcc", "-O3", "-march=x86-64-v3", "-fopenmp", "-shared", "-fPIC", "-std=gnu11", "-o", LIB, src, "-ldl", "-lm"]
    subprocess.check_call(cmd)
    return LIB


if __name__ == "__main__":
    print(build(force="--force" in sys.argv))
