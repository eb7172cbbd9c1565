"""Build glue/harness/libglue_harness.so: the psi4-side glue (glue/B200MemDFJK.cc, unmodified) linked against the
harness stand-ins of psi4's classes (glue/harness/) and libb200jk.so, so the glue can be EXECUTED on a GPU box that has
no psi4 (tests/test_glue_harness.py).  Test infrastructure; plain g++."""
import os
import subprocess

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
LIB = os.path.join(HERE, "libglue_harness.so")
SOURCES = [os.path.join(ROOT, "glue", "B200MemDFJK.cc"), os.path.join(HERE, "harness.cc")]


def _deps():
    out = list(SOURCES) + [os.path.join(ROOT, "glue", "B200MemDFJK.h"), os.path.join(ROOT, "include", "b200jk.h")]
    for d, _, fs in os.walk(os.path.join(HERE, "include")):
        out += [os.path.join(d, f) for f in fs]
    return out


def build(force: bool = False) -> str:
    if not force and os.path.exists(LIB) and all(os.path.getmtime(f) <= os.path.getmtime(LIB) for f in _deps()):
        return LIB
    eng = os.path.join(ROOT, "psi4_b200")
    cmd = ["g++", "-std=c++17", "-O2", "-Wall", "-Werror", "-fPIC", "-shared", "-I" + os.path.join(ROOT, "include"),
           "-I" + os.path.join(ROOT, "glue"), "-I" + os.path.join(HERE, "include"), *SOURCES, "-L" + eng, "-lb200jk",
           "-Wl,-rpath," + eng, "-Wl,-rpath,$ORIGIN/../../psi4_b200", "-o", LIB]
    subprocess.check_call(cmd)
    return LIB


if __name__ == "__main__":
    print(build(force=True))
