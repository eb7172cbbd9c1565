"""Row f3: the psi4-side glue, EXECUTED.

psi4 cannot be built in this image, so glue/B200MemDFJK.cc (the file a psi4 maintainer would compile, unmodified) is
linked against stand-ins of the few psi4 classes it touches (glue/harness/: JK / MemDFJK / DFHelper / Matrix / Options
with the reference's member names, types, virtuals and access, and the C1 branch of JK::compute restated from
libfock/jk.cc:595-681) and driven the way psi4's SCF driver drives a JK object:

    build (JK::build_JK's MEM_DF branch, jk.cc:143-148) -> set_do_wK / set_omega -> initialize() ->
    [C_left().clear(); push_back; compute(); J()/K()/wK()] x iterations with a CHANGING matrix count -> finalize()

CPU part (here): the stand-in declarations are probed against the reference's REAL headers, so the harness cannot drift
from what the glue meets inside psi4.  GPU part (-m gpu): every J / K / wK the glue returns is compared with the oracle."""
import ctypes as ct
import os
import shutil
import subprocess
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF = "/root/reference/psi4"
HARNESS = os.path.join(ROOT, "glue", "harness")


def _harness_build():
    import importlib.util

    spec = importlib.util.spec_from_file_location("glue_harness_build", os.path.join(HARNESS, "build.py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


def _syntax(includes):
    cmd = ["g++", "-std=c++17", "-fsyntax-only", "-I" + os.path.join(ROOT, "include"), *["-I" + i for i in includes],
           os.path.join(HARNESS, "probe_types.cc")]
    r = subprocess.run(cmd, capture_output=True, text=True)
    assert r.returncode == 0, r.stderr[-3000:]


@pytest.mark.skipif(not os.path.isdir(REF) or shutil.which("g++") is None, reason="needs the reference checkout and g++")
def test_harness_declarations_agree_with_the_reference_headers():
    """One probe translation unit (member types via static_assert, the overridden virtuals, every setter the glue's
    factory calls) compiles against the reference's headers and against the harness stand-ins alike."""
    _syntax([os.path.join(REF, "include"), os.path.join(REF, "src"), os.path.join(ROOT, "glue", "compile_check")])
    _syntax([os.path.join(HARNESS, "include")])


@pytest.mark.skipif(shutil.which("g++") is None, reason="needs g++")
def test_harness_library_builds_and_exports_the_driver():
    lib = _harness_build().build()
    syms = subprocess.run(["nm", "-D", "--defined-only", lib], capture_output=True, text=True).stdout
    for s in ("harness_build", "harness_initialize", "harness_compute", "harness_inject", "harness_header"):
        assert s in syms


def _lib():
    L = ct.CDLL(_harness_build().build())
    L.harness_last_error.restype = ct.c_char_p
    L.harness_build.restype = ct.c_void_p
    L.harness_build.argtypes = [ct.c_int, ct.c_int, ct.c_int, ct.c_int]
    L.harness_destroy.argtypes = [ct.c_void_p]
    dp, dpp = ct.POINTER(ct.c_double), ct.POINTER(ct.POINTER(ct.c_double))
    L.harness_inject.argtypes = [ct.c_int, ct.POINTER(ct.c_ubyte), ct.c_size_t, dp, dp, dp, ct.c_double]
    L.harness_set_tasks.argtypes = [ct.c_void_p, ct.c_int, ct.c_int, ct.c_int, ct.c_double]
    L.harness_initialize.argtypes = [ct.c_void_p]
    L.harness_finalize.argtypes = [ct.c_void_p]
    L.harness_condition.restype = ct.c_double
    L.harness_condition.argtypes = [ct.c_void_p]
    L.harness_cutoff.restype = ct.c_double
    L.harness_cutoff.argtypes = [ct.c_void_p]
    L.harness_pinned.argtypes = [ct.c_void_p]
    L.harness_tensors_on_host.argtypes = [ct.c_void_p]
    L.harness_compute.argtypes = [ct.c_void_p, ct.c_int, ct.c_int, ct.POINTER(ct.c_int), dpp, dpp, dpp, dpp, dpp]
    L.harness_header.argtypes = [ct.c_void_p, ct.c_char_p, ct.c_size_t]
    L.harness_set_option_double.argtypes = [ct.c_char_p, ct.c_double]
    L.harness_set_option_str.argtypes = [ct.c_char_p, ct.c_char_p]
    L.harness_timer_calls.argtypes = [ct.c_char_p]
    L.harness_last_launches.argtypes = [ct.c_void_p]
    return L


def _ok(L, rc):
    assert rc == 0, L.harness_last_error().decode()


def _ptrs(arrs):
    dp = ct.POINTER(ct.c_double)
    return (dp * len(arrs))(*[a.ctypes.data_as(dp) for a in arrs])


def _compute(L, s, n, Cl, Cr, want):
    nmat = len(Cl)
    nocc = (ct.c_int * nmat)(*[c.shape[1] for c in Cl])
    out = {k: [np.full((n, n), np.nan) for _ in range(nmat)] for k in "JKW"}
    _ok(L, L.harness_compute(s, nmat, n, nocc, _ptrs(Cl), _ptrs(Cr) if Cr is not None else None,
                             _ptrs(out["J"]) if "J" in want else None, _ptrs(out["K"]) if "K" in want else None,
                             _ptrs(out["W"]) if "W" in want else None))
    return out


@pytest.mark.gpu
def test_glue_runs_an_scf_like_sequence_and_matches_the_oracle(oracle):
    from psi4_b200 import DFHelper

    L = _lib()
    rng = np.random.default_rng(21)
    n, a, omega = 75, 113, 0.4
    r = rng.random((n, n))
    keep = (r + r.T) < 1.45
    np.fill_diagonal(keep, True)
    d = DFHelper(n, a)
    d.prepare_sparsity(keep=keep)

    def packed():
        B = rng.standard_normal((a, n, n)) * 0.1
        return d.pack(B + B.transpose(0, 2, 1))

    P, M1, W = packed(), packed(), packed()
    keep8 = np.ascontiguousarray(keep, dtype=np.uint8)
    dp = ct.POINTER(ct.c_double)
    _ok(L, L.harness_inject(n, keep8.ctypes.data_as(ct.POINTER(ct.c_ubyte)), P.size, P.ctypes.data_as(dp),
                            M1.ctypes.data_as(dp), W.ctypes.data_as(dp), omega))
    sp = oracle.Sparsity(keep8, a)

    # the factory reads psi4's options exactly as JK::build_JK does for MEM_DF (condition_ was left at 1e-12)
    L.harness_reset_options()
    s = L.harness_build(n, a, 1, 1)
    assert s, L.harness_last_error().decode()
    assert L.harness_condition(s) == 1.0e-10  # DF_FITTING_CONDITION default, read_options.cc:1734
    assert L.harness_cutoff(s) == 1.0e-12
    L.harness_destroy(s)
    L.harness_set_option_double(b"DF_FITTING_CONDITION", 1.0e-8)
    L.harness_set_option_double(b"INTS_TOLERANCE", 1.0e-9)
    s = L.harness_build(n, a, 1, 1)
    assert L.harness_condition(s) == 1.0e-8 and L.harness_cutoff(s) == 1.0e-9
    L.harness_destroy(s)
    L.harness_set_option_str(b"SCREENING", b"NONE")
    s = L.harness_build(n, a, 1, 1)
    assert L.harness_cutoff(s) == 0.0
    L.harness_destroy(s)
    L.harness_reset_options()

    # ---- an SCF-like life cycle ----
    s = L.harness_build(n, a, 1, 1)
    _ok(L, L.harness_initialize(s))                 # MemDFJK::preiterations + upload
    assert L.harness_tensors_on_host(s) == 0        # release_host: DFHelper's copy is gone once HBM has it
    assert L.harness_timer_calls(b"JK: B200 upload") >= 1
    buf = ct.create_string_buffer(4096)
    _ok(L, L.harness_header(s, buf, 4096))
    hdr = buf.value.decode()
    assert "MemDFJK: Density-Fitted J/K Matrices" in hdr and "B200 DF-JK engine" in hdr and "GPUs (Q shards)" in hdr

    def check(Cl, Cr, want="JK", m1=None, w=None):
        got = _compute(L, s, n, Cl, Cr, want)
        D = [x @ (x if Cr is None else y).T for x, y in zip(Cl, Cl if Cr is None else Cr)]
        Jo, Ko, Wo, _ = oracle.build_JK(sp, P, Cl, Cr, D=D, do_wK="W" in want, m1Ppq=m1, wPpq=w)
        for i in range(len(Cl)):
            assert np.abs(got["J"][i] - Jo[i]).max() < 1e-10
            assert np.abs(got["K"][i] - Ko[i]).max() < 1e-10
            if "W" in want:
                assert np.abs(got["W"][i] - Wo[i]).max() < 1e-10
        return got

    C1 = [rng.standard_normal((n, 9))]
    g1 = check(C1, None)                                   # RHF-like: one symmetric density
    assert L.harness_pinned(s) == 3                        # D, J, K of the one matrix are page-locked
    launches = L.harness_last_launches(s)
    assert launches > 0
    C2 = [rng.standard_normal((n, 9)), rng.standard_normal((n, 6))]
    check(C2, None)                                        # UHF-like: matrix count changes -> psi4 re-creates D/J/K
    assert L.harness_pinned(s) == 6
    R3 = [rng.standard_normal((n, 5)) for _ in range(3)]
    C3 = [rng.standard_normal((n, 5))] * 3
    check(C3, R3)                                          # response-like: three right-hand sides, one C_left
    assert L.harness_pinned(s) == 9
    g1b = check(C1, None)                                  # back to one matrix: same bits as the first build
    assert L.harness_pinned(s) == 3
    assert np.array_equal(g1["J"][0], g1b["J"][0]) and np.array_equal(g1["K"][0], g1b["K"][0])
    assert L.harness_timer_calls(b"JK: B200 build") == 4

    # a second initialize() on the same object is legal in the reference (it recomputes); same results afterwards
    _ok(L, L.harness_initialize(s))
    g1c = check(C1, None)
    assert np.array_equal(g1["K"][0], g1c["K"][0])
    _ok(L, L.harness_finalize(s))
    L.harness_destroy(s)

    # ---- range-separated exchange: set_do_wK / set_omega between build and initialize (scf_iterator.py:112-135) ----
    s = L.harness_build(n, a, 1, 0)
    _ok(L, L.harness_set_tasks(s, 1, 1, 1, omega))
    _ok(L, L.harness_initialize(s))
    assert L.harness_tensors_on_host(s) == 1               # release_host = False keeps DFHelper's copy
    check(C1, None, "JKW", M1, W)                          # hermitivitized wK (MemDFJK.cc:104-110)
    check(C2, [rng.standard_normal(c.shape) for c in C2], "JKW", M1, W)
    L.harness_destroy(s)

    # ---- errors surface as PSIEXCEPTION text, never as a silent CPU build ----
    s = L.harness_build(n, a, 1, 1)
    rc = L.harness_compute(s, 1, n, (ct.c_int * 1)(9), _ptrs(C1), None, None, None, None)  # compute() before initialize()
    assert rc != 0 and b"B200MemDFJK" in L.harness_last_error()
    L.harness_destroy(s)
