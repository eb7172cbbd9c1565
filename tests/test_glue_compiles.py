"""Row f3: the psi4-side glue (glue/B200MemDFJK.{h,cc}, glue/export_b200jk.cc) type-checks against
the reference's REAL headers (libfock/jk.h, lib3index/dfhelper.h, libmints/matrix.h, liboptions, pybind11.h) and
include/b200jk.h.  psi4 itself cannot be built in this image (Libint2 / Eigen / LibXC absent), so this is a
g++ -fsyntax-only pass with two declaration-only stand-ins (glue/compile_check/) for <eigen3/Eigen/Core> and
<libint2/shell.h>; it proves every protected member, virtual signature and C-ABI call the glue uses exists with
that type in the reference tree.  Skipped where /root/reference is absent (the GPU box)."""
import os
import shutil
import subprocess
import sysconfig

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF = "/root/reference/psi4"

pytestmark = pytest.mark.skipif(not os.path.isdir(REF) or shutil.which("g++") is None,
                                reason="needs the reference checkout and g++")


def syntax_check(src, extra=()):
    cmd = ["g++", "-std=c++17", "-Wall", "-Werror", "-fsyntax-only", "-I" + os.path.join(ROOT, "include"),
           "-I" + os.path.join(REF, "include"), "-I" + os.path.join(REF, "src"),
           "-I" + os.path.join(ROOT, "glue", "compile_check"), "-I" + os.path.join(ROOT, "glue"), *extra,
           os.path.join(ROOT, "glue", src)]
    r = subprocess.run(cmd, capture_output=True, text=True)
    assert r.returncode == 0, r.stderr[-4000:]


def test_glue_subclass_type_checks_against_reference_headers():
    syntax_check("B200MemDFJK.cc")


def test_glue_pybind_export_type_checks():
    pybind11 = pytest.importorskip("pybind11")
    syntax_check("export_b200jk.cc", ["-I" + pybind11.get_include(), "-I" + sysconfig.get_paths()["include"]])
