"""Build the CPU oracle (test infrastructure) into oracle/_build/libdfjk_oracle.so.

The reference path itself (lib3index/dfhelper.cc) cannot be compiled here: it includes
libmints/basisset.h -> <libint2/shell.h>, and Libint2 is neither installed nor vendored
(SURVEY.md 8c).  So there is no oracle/_ref; the restatement in dfjk_oracle.c is the oracle.
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
