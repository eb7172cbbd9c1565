"""Build oracle/_ref/libref_dfjk.so: the reference's OWN hot-path functions, compiled from where they lie.

psi4 as a whole cannot be built in this image (Libint2 / LibXC / gau2grid / Eigen are absent, and
lib3index/dfhelper.cc pulls their headers in through libmints), but the functions on the MEM_DF J/K path never touch
them: they are loops over std::vector tables and three BLAS calls.  This recipe slices exactly those member-function
definitions out of /root/reference/psi4/src/psi4/lib3index/dfhelper.cc at build time (located by signature, copied
byte for byte into a translation unit that exists only in memory and is piped to g++ -- no reference source is ever
written into the repository tree; only the .so files land in the git-ignored oracle/_ref/), declares them in a stand-in `class DFHelper` whose members carry the reference's names and types
(ref_members.inl), and compiles them against ref_shim.h together with the reference's libqt BLAS wrappers
(libqt/blas_intfc23.cc, blas_intfc.cc -- whole files, unmodified).  Result: the arithmetic the restatement is checked against is
the reference's own object code.

  Qshell_blocks_for_JK_build, contract_metric_AO_core_symm, first_transform_pQq, build_JK, compute_JK,
  compute_J_symm, fill, compute_J, compute_J_combined, compute_K, compute_wK,
  and the table-building half of prepare_sparsity (everything after the Libint2 loop)

Skipped (returns None) when /root/reference is absent, e.g. on the GPU box: the prebuilt .so travels with the repo.
"""
import os
import re
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
REF_SRC = "/root/reference/psi4/src/psi4/lib3index/dfhelper.cc"
OUT = os.path.join(HERE, "_ref")
LIB = os.path.join(OUT, "libref_dfjk.so")
FUNCTIONS = ["Qshell_blocks_for_JK_build", "contract_metric_AO_core_symm", "first_transform_pQq", "build_JK", "compute_JK",
             "compute_J_symm", "fill", "compute_J", "compute_J_combined", "compute_K", "compute_wK"]


def slice_function(text: str, name: str) -> tuple[str, str]:
    """(declaration, definition) of DFHelper::name: from its signature line to the closing brace in column 0."""
    m = re.search(r"^[^\n/]*\bDFHelper::" + re.escape(name) + r"\(", text, re.M)
    if not m:
        raise RuntimeError(f"DFHelper::{name} not found in {REF_SRC}")
    start = m.start()
    end = text.index("\n}\n", start) + 3
    definition = text[start:end]
    signature = definition[:definition.index("{")].strip()
    declaration = signature.replace("DFHelper::", "", 1) + ";"
    return declaration, definition


def slice_sparsity_tables(text: str) -> str:
    """Second half of DFHelper::prepare_sparsity (dfhelper.cc:370-420): everything from the tolerance to the end of the
    function -- mask rule and all index tables -- as the body of a member that receives the Schwarz maxima the first
    half would have computed with Libint2."""
    start = text.index("void DFHelper::prepare_sparsity()")
    a = text.index("    // => Prepare screening/indexing data <=", start)
    b = text.index("\n}\n", a)
    body = text[a:b].replace('timer_off("DFH: sparsity prep");', "")
    return ("void DFHelper::prepare_sparsity_tables(std::vector<double>& shell_max_vals, std::vector<double>& fun_max_vals,\n"
            "                                       double max_val) {\n" + body + "\n}\n")


# libqt's row-major-over-Fortran BLAS/LAPACK wrappers compile as whole files, straight from the reference (FC_SYMBOL == 2:
# lower-case symbol + underscore, what psi4's CMake detects for gfortran-style BLAS); their unused Fortran externs stay
# unresolved, so the libraries are linked and loaded with lazy binding
LIBQT = "/root/reference/psi4/src/psi4/libqt"
LIBQT_FLAGS = ["-DFC_SYMBOL=2", "-I/root/reference/psi4/include", "-I/root/reference/psi4/src", "-Wl,-z,lazy"]
REF_MATRIX_SRC = "/root/reference/psi4/src/psi4/libmints/matrix.cc"
LIB_MATRIX = os.path.join(OUT, "libref_matrix.so")


def build_matrix(force: bool = False):
    """oracle/_ref/libref_matrix.so: Matrix::power (libmints/matrix.cc:2370-2424) compiled from the reference."""
    if not os.path.exists(REF_MATRIX_SRC):
        return LIB_MATRIX if os.path.exists(LIB_MATRIX) else None
    deps = [REF_MATRIX_SRC, os.path.join(HERE, "ref_matrix_shim.h"), os.path.join(HERE, "ref_matrix_entry.inl"),
            os.path.abspath(__file__)]
    if not force and os.path.exists(LIB_MATRIX) and all(os.path.getmtime(LIB_MATRIX) >= os.path.getmtime(d) for d in deps):
        return LIB_MATRIX
    os.makedirs(OUT, exist_ok=True)
    text = open(REF_MATRIX_SRC).read()
    start = text.index("Dimension Matrix::power(double alpha, double cutoff) {")
    end = text.index("\n}\n", start) + 3
    src = ('#include "ref_matrix_shim.h"\nnamespace psi {\n' + text[start:end] + '}  // namespace psi\n'
           '#include "ref_matrix_entry.inl"\n')
    cmd = ["g++", "-O2", "-shared", "-fPIC", "-std=c++17", "-w", "-I" + HERE, "-o", LIB_MATRIX, *LIBQT_FLAGS, "-x", "c++", "-",
           os.path.join(LIBQT, "lapack_intfc.cc"), os.path.join(LIBQT, "blas_intfc.cc"),
           os.path.join(LIBQT, "blas_intfc23.cc"), "-ldl"]
    subprocess.run(cmd, input=src.encode(), check=True)
    return LIB_MATRIX


def build(force: bool = False):
    if not os.path.exists(REF_SRC):
        return LIB if os.path.exists(LIB) else None
    deps = [REF_SRC, os.path.join(HERE, "ref_shim.h"), os.path.join(HERE, "ref_members.inl"),
            os.path.join(HERE, "ref_entry.inl"), os.path.abspath(__file__)]
    if not force and os.path.exists(LIB) and all(os.path.getmtime(LIB) >= os.path.getmtime(d) for d in deps):
        return LIB
    os.makedirs(OUT, exist_ok=True)
    text = open(REF_SRC).read()
    decls, defs = zip(*(slice_function(text, f) for f in FUNCTIONS))
    src = ["// generated in memory by oracle/ref_build.py from " + REF_SRC + "; piped to g++, never written to disk",
           '#include "ref_shim.h"', "namespace psi {", "class DFHelper {", "   public:", '#include "ref_members.inl"']
    src += ["    " + " ".join(d.split()) for d in decls]
    src += ["    void prepare_sparsity_tables(std::vector<double>& shell_max_vals, std::vector<double>& fun_max_vals, "
            "double max_val);", "};", ""]
    src += list(defs) + [slice_sparsity_tables(text), "}  // namespace psi", '#include "ref_entry.inl"', ""]
    cmd = ["g++", "-O2", "-march=x86-64-v3", "-fopenmp", "-shared", "-fPIC", "-std=c++17", "-w", "-I" + HERE, "-o", LIB,
           *LIBQT_FLAGS, "-x", "c++", "-", os.path.join(LIBQT, "blas_intfc23.cc"), os.path.join(LIBQT, "blas_intfc.cc"), "-ldl"]
    subprocess.run(cmd, input="\n".join(src).encode(), check=True)
    return LIB


if __name__ == "__main__":
    print(build(force="--force" in sys.argv))
    print(build_matrix(force="--force" in sys.argv))
