"""ctypes front end of the CPU oracle + an independent dense-einsum second opinion.

TEST INFRASTRUCTURE ONLY: imported by tests/, __graft_entry__.smoke() and bench.py's
cpu_baseline / --impl reference legs.  Nothing under psi4_b200/ may import this module.

Reference anchors (paths relative to /root/reference/psi4/src/psi4/):
  tables   lib3index/dfhelper.cc:299-420      layout  :1274-1276, :1666-1677
  J        :3162-3223 (symmetric), :3230-3285 half-transform :2162-2186
  K        :3350-3377                          wK      :3378-3438
  outer    libfock/MemDFJK.cc:97-111, libfock/jk.cc:314-354, :595-681
The dense oracle mirrors tests/python/3-index-transforms/input.py (einsum definitions).
"""
from __future__ import annotations

import ctypes as ct
import glob
import os
import sys

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, _HERE)
import build as _build  # noqa: E402

_lib = None
_sz = ct.c_size_t
_dp = ct.POINTER(ct.c_double)
_szp = ct.POINTER(ct.c_size_t)


def find_openblas() -> str:
    import scipy

    root = os.path.join(os.path.dirname(os.path.dirname(scipy.__file__)), "scipy.libs")
    hits = sorted(glob.glob(os.path.join(root, "libscipy_openblas-*.so")))
    if not hits:
        raise RuntimeError("LP64 OpenBLAS bundled with scipy not found under " + root)
    return hits[0]


def lib():
    global _lib
    if _lib is None:
        path = _build.build()
        L = ct.CDLL(path)
        L.oracle_init.argtypes = [ct.c_char_p]
        L.oracle_init.restype = ct.c_int
        L.oracle_blas_config.restype = ct.c_char_p
        L.oracle_set_blas_threads.argtypes = [ct.c_int]
        L.oracle_get_blas_threads.restype = ct.c_int
        L.oracle_max_threads.restype = ct.c_int
        L.oracle_schwarz_mask.argtypes = [_sz, _dp, ct.c_double, ct.POINTER(ct.c_ubyte)]
        L.oracle_prepare_sparsity.argtypes = [_sz, _sz, ct.POINTER(ct.c_ubyte), _szp, _szp, _szp, _szp, _szp, _szp]
        L.oracle_pack_pQq.argtypes = [_sz, _sz, _dp, _szp, _szp, _szp, _dp]
        L.oracle_build_JK.argtypes = [
            _sz, _sz, ct.c_int, _szp, _szp, _szp, _szp, _szp, _dp, _dp, _dp, ct.c_int,
            ct.POINTER(_dp), ct.POINTER(_dp), ct.POINTER(ct.c_int), ct.POINTER(_dp),
            ct.POINTER(_dp), ct.POINTER(_dp), ct.POINTER(_dp), ct.c_int, ct.c_int, ct.c_int, ct.c_int, _sz,
        ]
        L.oracle_build_JK.restype = ct.c_int
        L.oracle_last_timings.argtypes = [_dp]
        L.oracle_compute_D.argtypes = [_sz, ct.c_int, _dp, _dp, _dp]
        L.oracle_synth_fill.argtypes = [_sz, _sz, _sz, ct.c_uint64, _dp, _szp, _szp, _szp, _dp]
        L.oracle_synth_rowblock.argtypes = [_sz, _sz, ct.c_uint64, _dp, ct.POINTER(ct.c_ubyte), _sz, _dp]
        L.oracle_synth_dq.argtypes = [_sz, _sz, ct.c_uint64, _dp, ct.POINTER(ct.c_ubyte), _dp, _dp]
        rc = L.oracle_init(find_openblas().encode())
        if rc:
            raise RuntimeError(f"oracle_init failed rc={rc}")
        _lib = L
    return _lib


_ref = None
_JK_ARGTYPES = [
    _sz, _sz, ct.c_int, _szp, _szp, _szp, _szp, _szp, _dp, _dp, _dp, ct.c_int,
    ct.POINTER(_dp), ct.POINTER(_dp), ct.POINTER(ct.c_int), ct.POINTER(_dp),
    ct.POINTER(_dp), ct.POINTER(_dp), ct.POINTER(_dp), ct.c_int, ct.c_int, ct.c_int, ct.c_int, _sz,
]


def ref_lib():
    """oracle/_ref/libref_dfjk.so: the reference's own DFHelper functions compiled from /root/reference by
    oracle/ref_build.py (None when neither the reference checkout nor a prebuilt library is present)."""
    global _ref
    if _ref is None:
        import ref_build

        path = ref_build.build()
        if path is None:
            return None
        L = ct.CDLL(path, mode=os.RTLD_LAZY)  # libqt's unused Fortran externs stay unresolved
        L.ref_init.argtypes = [ct.c_char_p]
        L.ref_build_JK.argtypes = _JK_ARGTYPES
        L.ref_build_JK.restype = ct.c_int
        L.ref_contract_metric_AO_core_symm.argtypes = [_sz, _sz, ct.c_int, _szp, _szp, _szp, _szp, _szp, _szp, _dp, _dp, _dp,
                                                       _sz, _sz]
        L.ref_contract_metric_AO_core_symm.restype = ct.c_int
        if L.ref_init(find_openblas().encode()):
            raise RuntimeError("ref_init failed")
        lib()  # the BLAS thread controls live in the restatement's library (same OpenBLAS instance)
        _ref = L
    return _ref


def ref_matrix_power(A, alpha, cutoff):
    """Matrix::power (libmints/matrix.cc:2370-2424) as compiled from the reference: (A^alpha, eigenvalues kept)."""
    import ref_build

    path = ref_build.build_matrix()
    if path is None:
        raise RuntimeError("oracle/_ref/libref_matrix.so is not available")
    L = ct.CDLL(path, mode=os.RTLD_LAZY)
    L.refm_init.argtypes = [ct.c_char_p]
    L.refm_power.argtypes = [_dp, ct.c_int, ct.c_double, ct.c_double]
    if L.refm_init(find_openblas().encode()):
        raise RuntimeError("refm_init failed")
    M = np.array(A, dtype=np.float64, order="C", copy=True)
    kept = L.refm_power(_d(M), M.shape[0], float(alpha), float(cutoff))
    if kept < 0:
        raise RuntimeError(f"refm_power rc={kept}")
    return M, kept


def ref_sparsity_tables(fun_max_vals, naux, cutoff=1e-12, pshell_aggs=None):
    """The table-building half of the reference's own prepare_sparsity (oracle/_ref) on given Schwarz maxima.
    Returns a dict of the reference's member arrays."""
    L = ref_lib()
    if L is None:
        raise RuntimeError("oracle/_ref/libref_dfjk.so is not available")
    f = np.ascontiguousarray(fun_max_vals, dtype=np.float64)
    n = f.shape[0]
    aggs = np.arange(n + 1) if pshell_aggs is None else np.asarray(pshell_aggs, dtype=np.int64)
    ps = len(aggs) - 1
    shell = np.zeros((ps, ps))
    for i in range(ps):
        for j in range(ps):
            shell[i, j] = f[aggs[i]:aggs[i + 1], aggs[j]:aggs[j + 1]].max()
    out = {k: np.zeros(sz, dtype=np.uintp) for k, sz in
           (("schwarz_fun_index_", n * n), ("small_skips_", n + 1), ("big_skips_", n + 1), ("symm_small_skips_", n),
            ("symm_ignored_columns_", n), ("symm_big_skips_", n + 1), ("schwarz_shell_mask_", ps * ps))}
    L.ref_prepare_sparsity_tables.argtypes = [_sz, _sz, _sz, ct.c_double, _dp, _dp, ct.c_double] + [_szp] * 7
    rc = L.ref_prepare_sparsity_tables(n, naux, ps, cutoff, _d(shell), _d(f), float(f.max()), *[_s(v) for v in out.values()])
    if rc:
        raise RuntimeError(f"ref_prepare_sparsity_tables rc={rc}")
    return out


def _d(a):
    return a.ctypes.data_as(_dp)


def _s(a):
    return a.ctypes.data_as(_szp)


class Sparsity:
    """Index tables of DFHelper::prepare_sparsity for a boolean pair mask (dfhelper.cc:377-416)."""

    def __init__(self, keep: np.ndarray, naux: int):
        keep = np.ascontiguousarray(keep, dtype=np.uint8)
        n = keep.shape[0]
        assert keep.shape == (n, n)
        self.nbf, self.naux = n, int(naux)
        self.keep = keep
        self.fun_index = np.zeros(n * n, dtype=np.uintp)
        self.small_skips = np.zeros(n + 1, dtype=np.uintp)
        self.big_skips = np.zeros(n + 1, dtype=np.uintp)
        self.symm_small_skips = np.zeros(n, dtype=np.uintp)
        self.symm_ignored_columns = np.zeros(n, dtype=np.uintp)
        self.symm_big_skips = np.zeros(n + 1, dtype=np.uintp)
        lib().oracle_prepare_sparsity(
            n, self.naux, keep.ctypes.data_as(ct.POINTER(ct.c_ubyte)), _s(self.fun_index), _s(self.small_skips),
            _s(self.big_skips), _s(self.symm_small_skips), _s(self.symm_ignored_columns), _s(self.symm_big_skips))

    @property
    def packed_size(self) -> int:
        return int(self.big_skips[self.nbf])

    def with_naux(self, naux: int) -> "Sparsity":
        return Sparsity(self.keep, naux)


def schwarz_mask(fun_max_vals: np.ndarray, cutoff: float) -> np.ndarray:
    f = np.ascontiguousarray(fun_max_vals, dtype=np.float64)
    n = f.shape[0]
    keep = np.zeros((n, n), dtype=np.uint8)
    lib().oracle_schwarz_mask(n, _d(f), float(cutoff), keep.ctypes.data_as(ct.POINTER(ct.c_ubyte)))
    return keep


def pack_pQq(sp: Sparsity, dense_Qmn: np.ndarray) -> np.ndarray:
    dense = np.ascontiguousarray(dense_Qmn, dtype=np.float64)
    assert dense.shape == (sp.naux, sp.nbf, sp.nbf)
    out = np.zeros(sp.packed_size, dtype=np.float64)
    lib().oracle_pack_pQq(sp.nbf, sp.naux, _d(dense), _s(sp.fun_index), _s(sp.small_skips), _s(sp.big_skips), _d(out))
    return out


def synth_fill(sp: Sparsity, q0: int, nq: int, seed: int, amp: np.ndarray) -> np.ndarray:
    """Packed synthetic tensor rows Q in [q0,q0+nq); sp must have naux == nq."""
    assert sp.naux == nq
    amp = np.ascontiguousarray(amp, dtype=np.float64)
    out = np.zeros(sp.packed_size, dtype=np.float64)
    lib().oracle_synth_fill(sp.nbf, q0, nq, ct.c_uint64(seed), _d(amp), _s(sp.fun_index), _s(sp.small_skips),
                            _s(sp.big_skips), _d(out))
    return out


def synth_rowblock(keep, naux, seed, amp, m):
    """Dense B[:, m, :] (naux, nbf) of the synthetic tensor, regenerated from the counter hash."""
    keep = np.ascontiguousarray(keep, dtype=np.uint8)
    amp = np.ascontiguousarray(amp, dtype=np.float64)
    n = keep.shape[0]
    out = np.empty((naux, n))
    lib().oracle_synth_rowblock(n, naux, ct.c_uint64(seed), _d(amp), keep.ctypes.data_as(ct.POINTER(ct.c_ubyte)), m, _d(out))
    return out


def synth_dq(keep, naux, seed, amp, D):
    """d[q] = sum_mn B(q,m,n) D[m,n] over the whole synthetic tensor (one pass over naux*nbf^2 hash evaluations)."""
    keep = np.ascontiguousarray(keep, dtype=np.uint8)
    amp = np.ascontiguousarray(amp, dtype=np.float64)
    D = np.ascontiguousarray(D, dtype=np.float64)
    d = np.empty(naux)
    lib().oracle_synth_dq(keep.shape[0], naux, ct.c_uint64(seed), _d(amp), keep.ctypes.data_as(ct.POINTER(ct.c_ubyte)),
                          _d(D), _d(d))
    return d


def contract_metric_AO_core_symm(sp: Sparsity, Qpq_sym: np.ndarray, metp: np.ndarray, Ppq: np.ndarray | None = None,
                                 begin: int = 0, end: int | None = None, nthreads=None, impl="port") -> np.ndarray:
    """dfhelper.cc:1653-1678 on the host: fitted + mirrored rows [begin, end] written into Ppq (allocated if None).
    Qpq_sym is relative to symm_big_skips[begin]."""
    L = lib()
    if "_cm" not in L.__dict__:
        L.oracle_contract_metric_AO_core_symm.argtypes = [_sz, _sz, ct.c_int, _szp, _szp, _szp, _szp, _szp, _szp, _dp, _dp,
                                                          _dp, _sz, _sz]
        L.oracle_contract_metric_AO_core_symm.restype = ct.c_int
        L.__dict__["_cm"] = True
    end = sp.nbf - 1 if end is None else end
    if Ppq is None:
        Ppq = np.zeros(sp.packed_size)
    q = np.ascontiguousarray(Qpq_sym, dtype=np.float64)
    m = np.ascontiguousarray(metp, dtype=np.float64)
    fn = L.oracle_contract_metric_AO_core_symm
    if impl == "ref":
        if ref_lib() is None:
            raise RuntimeError("oracle/_ref/libref_dfjk.so is not available")
        fn = ref_lib().ref_contract_metric_AO_core_symm
    rc = fn(sp.nbf, sp.naux, nthreads or L.oracle_max_threads(), _s(sp.fun_index),
                                               _s(sp.small_skips), _s(sp.big_skips), _s(sp.symm_small_skips),
                                               _s(sp.symm_ignored_columns), _s(sp.symm_big_skips), _d(q), _d(Ppq), _d(m),
                                               begin, end)
    if rc:
        raise RuntimeError("oracle_contract_metric_AO_core_symm failed")
    return Ppq


def _ptrs(arrs):
    return (_dp * len(arrs))(*[_d(a) for a in arrs])


def build_JK(sp: Sparsity, Ppq, Cleft, Cright=None, D=None, do_J=True, do_K=True, do_wK=False, m1Ppq=None, wPpq=None,
             nthreads=None, q_block=0, impl="port"):
    """The reference's MemDFJK::compute_JK on host arrays.  Cright=None => lr_symmetric (jk.cc:597-602).
    impl="port": the C restatement (dfjk_oracle.c); impl="ref": the reference's own functions (oracle/_ref).

    Returns (J, K, wK, timings) lists of (nbf,nbf) arrays (None where not tasked)."""
    L = lib()
    if impl == "ref" and ref_lib() is None:
        raise RuntimeError("oracle/_ref/libref_dfjk.so is not available (no reference checkout, nothing prebuilt)")
    n = sp.nbf
    nmat = len(Cleft)
    lr = Cright is None
    Cl = [np.ascontiguousarray(c, dtype=np.float64).reshape(n, -1) for c in Cleft]
    Cr = Cl if lr else [np.ascontiguousarray(c, dtype=np.float64).reshape(n, -1) for c in Cright]
    nocc = (ct.c_int * nmat)(*[c.shape[1] for c in Cl])
    for a, b in zip(Cl, Cr):
        if a.shape != b.shape:
            raise ValueError("JK: C_left/C_right MO zip index size mismatch!")  # jk.cc:612-615
    if D is None:
        D = [compute_D(a, b) for a, b in zip(Cl, Cr)]
    D = [np.ascontiguousarray(d, dtype=np.float64) for d in D]
    J = [np.zeros((n, n)) for _ in range(nmat)]
    K = [np.zeros((n, n)) for _ in range(nmat)]
    wK = [np.zeros((n, n)) for _ in range(nmat)]
    nthreads = nthreads or L.oracle_max_threads()
    null = ct.cast(None, _dp)
    import time

    t0 = time.perf_counter()
    rc = (ref_lib().ref_build_JK if impl == "ref" else L.oracle_build_JK)(
        n, sp.naux, nthreads, _s(sp.fun_index), _s(sp.small_skips), _s(sp.big_skips), _s(sp.symm_small_skips),
        _s(sp.symm_ignored_columns), _d(Ppq), _d(m1Ppq) if m1Ppq is not None else null,
        _d(wPpq) if wPpq is not None else null, nmat, _ptrs(Cl), _ptrs(Cr), nocc, _ptrs(D), _ptrs(J), _ptrs(K),
        _ptrs(wK), int(do_J), int(do_K), int(do_wK), int(lr), int(q_block))
    wall = time.perf_counter() - t0
    if rc:
        raise RuntimeError(f"{'ref' if impl == 'ref' else 'oracle'}_build_JK rc={rc}")
    t = np.zeros(4)
    if impl != "ref":
        L.oracle_last_timings(_d(t))
    return (J if do_J else None, K if do_K else None, wK if do_wK else None,
            {"J": t[0], "K": t[1], "wK": t[2], "total": wall})


def compute_D(Cl, Cr):
    """jk.cc:351  D = Cl Cr^T."""
    Cl = np.ascontiguousarray(Cl, dtype=np.float64)
    Cr = np.ascontiguousarray(Cr, dtype=np.float64)
    n, o = Cl.shape
    D = np.zeros((n, n))
    lib().oracle_compute_D(n, o, _d(Cl), _d(Cr), _d(D))
    return D


# ---------------------------------------------------------------------------------------------
# Second opinion: dense einsum on the *masked* dense tensor.
# ---------------------------------------------------------------------------------------------
def dense_JK(dense_Qmn, keep, Cleft, Cright=None, D=None, dense_m1=None, dense_w=None):
    B = np.asarray(dense_Qmn) * np.asarray(keep, dtype=np.float64)[None, :, :]
    Cr_list = Cleft if Cright is None else Cright
    Js, Ks, wKs = [], [], []
    for i, (Cl, Cr) in enumerate(zip(Cleft, Cr_list)):
        Di = Cl @ Cr.T if D is None else D[i]
        if Cright is None:
            # symmetric path reads only the upper triangle of D (dfhelper.cc:3188)
            Di = np.triu(Di) + np.triu(Di, 1).T
        d = np.einsum("Qmn,mn->Q", B, Di)
        Js.append(np.einsum("Qmn,Q->mn", B, d))
        T1 = np.einsum("Qmn,ni->mQi", B, Cl)
        T2 = T1 if Cright is None else np.einsum("Qmn,ni->mQi", B, Cr)
        Ks.append(np.einsum("mQi,nQi->mn", T1, T2))
        if dense_m1 is not None:
            k = np.asarray(keep, dtype=np.float64)[None]
            W1 = np.einsum("Qmn,ni->mQi", dense_m1 * k, Cl)
            W2 = np.einsum("Qmn,ni->mQi", dense_w * k, Cr)
            w = np.einsum("mQi,nQi->mn", W1, W2)
            if Cright is None:
                w = 0.5 * (w + w.T)
            wKs.append(w)
    return Js, Ks, wKs
