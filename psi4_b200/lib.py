"""ctypes binding of libb200jk.so (include/b200jk.h) -- the same stub a psi4 maintainer would
write in C++.  Fails loudly if the CUDA library is missing or cannot run:
there is no CPU path in this package."""
from __future__ import annotations

import ctypes as ct
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libb200jk.so")

NCCL_ID_BYTES = 128
TENSOR_PPQ, TENSOR_M1PPQ, TENSOR_WPPQ = 0, 1, 2
ERR_NAMES = {1: "INVALID", 2: "CUDA", 3: "OOM", 4: "NCCL", 5: "NODEVICE"}

_dp = ct.POINTER(ct.c_double)
_dpp = ct.POINTER(_dp)
_szp = ct.POINTER(ct.c_size_t)


class Stats(ct.Structure):
    _fields_ = [
        ("ms_total", ct.c_double), ("ms_j", ct.c_double), ("ms_half", ct.c_double), ("ms_kgemm", ct.c_double),
        ("ms_allreduce", ct.c_double), ("ms_h2d", ct.c_double), ("ms_d2h", ct.c_double),
        ("j_bytes", ct.c_double), ("half_flops", ct.c_double), ("half_bytes", ct.c_double),
        ("kgemm_flops", ct.c_double), ("launches", ct.c_uint64), ("hbm_tensor_bytes", ct.c_uint64),
        ("hbm_work_bytes", ct.c_uint64), ("n_shards", ct.c_int), ("q_begin", ct.c_int), ("q_end", ct.c_int),
        ("reduce_kind", ct.c_int), ("kgemm_kind", ct.c_int), ("kgemm_moduli", ct.c_int), ("half_kind", ct.c_int),
        ("ms_half_i8", ct.c_double * 4), ("ms_kgemm_i8", ct.c_double * 3), ("half_i8_ops", ct.c_double),
        ("half_i8_plane_bytes", ct.c_double), ("half_i8_convert_bytes", ct.c_double), ("kgemm_i8_ops", ct.c_double),
        ("half_moduli", ct.c_int), ("half_i8_chunks", ct.c_int), ("half_i8_cached", ct.c_int), ("half_i8_resident_rows", ct.c_int),
    ]

    def as_dict(self):
        out = {}
        for k, _ in self._fields_:
            v = getattr(self, k)
            out[k] = list(v) if hasattr(v, "__len__") else v
        return out


class B200JKError(RuntimeError):
    """Engine error; the psi4 glue would rethrow this as PSIEXCEPTION."""

    def __init__(self, code, msg):
        super().__init__(f"b200jk error {code} ({ERR_NAMES.get(code, '?')}): {msg}")
        self.code = code


_lib = None

# name -> (restype, argtypes); the full exported surface of include/b200jk.h
SIGNATURES = {
    "b200jk_create": (ct.c_int, [ct.POINTER(ct.c_void_p), ct.c_int, ct.POINTER(ct.c_int)]),
    "b200jk_create_rank": (ct.c_int, [ct.POINTER(ct.c_void_p), ct.c_int, ct.c_int, ct.c_int, ct.c_void_p]),
    "b200jk_nccl_unique_id": (ct.c_int, [ct.c_void_p]),
    "b200jk_destroy": (None, [ct.c_void_p]),
    "b200jk_set_layout": (ct.c_int, [ct.c_void_p, ct.c_size_t, ct.c_size_t, _szp, _szp, _szp]),
    "b200jk_upload": (ct.c_int, [ct.c_void_p, ct.c_int, _dp]),
    "b200jk_upload_rows": (ct.c_int, [ct.c_void_p, ct.c_int, ct.c_size_t, ct.c_size_t, _dp]),
    "b200jk_compute": (ct.c_int, [ct.c_void_p, ct.c_int, _dpp, _dpp, ct.POINTER(ct.c_int), _dpp, _dpp, _dpp, _dpp,
                                  ct.c_int, ct.c_int, ct.c_int]),
    "b200jk_compute_device": (ct.c_int, [ct.c_void_p, ct.c_int, _dpp, _dpp, ct.POINTER(ct.c_int), _dpp, _dpp, _dpp,
                                         _dpp, ct.c_int, ct.c_int, ct.c_int]),
    "b200jk_get_stats": (ct.c_int, [ct.c_void_p, ct.POINTER(Stats)]),
    "b200jk_last_error": (ct.c_char_p, [ct.c_void_p]),
    "b200jk_hbm_estimate": (ct.c_int, [ct.c_void_p, ct.c_size_t, ct.c_int, ct.POINTER(ct.c_uint64)]),
    "b200jk_set_work_budget": (ct.c_int, [ct.c_void_p, ct.c_uint64]),
    "b200jk_set_kgemm": (ct.c_int, [ct.c_void_p, ct.c_int, ct.c_int]),
    "b200jk_set_half": (ct.c_int, [ct.c_void_p, ct.c_int, ct.c_int]),
    "b200jk_fill_synthetic": (ct.c_int, [ct.c_void_p, ct.c_int, ct.c_uint64, _dp]),
    "b200jk_download_rows": (ct.c_int, [ct.c_void_p, ct.c_int, ct.c_size_t, ct.c_size_t, ct.c_size_t, _dp]),
    "b200jk_dev_alloc": (ct.c_int, [ct.c_void_p, ct.c_size_t, ct.POINTER(ct.c_void_p)]),
    "b200jk_dev_free": (ct.c_int, [ct.c_void_p, ct.c_void_p]),
    "b200jk_dev_copy": (ct.c_int, [ct.c_void_p, ct.c_void_p, ct.c_void_p, ct.c_size_t, ct.c_int]),
    "b200jk_fp64_peak": (ct.c_int, [ct.c_void_p, ct.c_int, ct.c_double, _dp]),
    "b200jk_register_host": (ct.c_int, [ct.c_void_p, ct.c_void_p, ct.c_size_t]),
    "b200jk_unregister_host": (ct.c_int, [ct.c_void_p, ct.c_void_p]),
    "b200jk_set_metric": (ct.c_int, [ct.c_void_p, _dp]),
    "b200jk_fit_rows": (ct.c_int, [ct.c_void_p, ct.c_int, ct.c_size_t, ct.c_size_t, _dp]),
    "b200jk_fit_stats": (ct.c_int, [ct.c_void_p, _dp, _dp]),
    "b200jk_grad_begin": (ct.c_int, [ct.c_void_p, ct.c_int, _dpp, ct.POINTER(ct.c_int), _dp, _dp]),
    "b200jk_grad_vectors": (ct.c_int, [ct.c_void_p, _dp, _dp]),
    "b200jk_grad_rows": (ct.c_int, [ct.c_void_p, ct.c_size_t, ct.c_size_t, _dp]),
    "b200jk_grad_end": (ct.c_int, [ct.c_void_p]),
    "b200jk_matrix_power": (ct.c_int, [ct.c_void_p, ct.c_size_t, _dp, ct.c_double, ct.c_double, _dp,
                                        ct.POINTER(ct.c_int), _dp]),
}


def load():
    """dlopen libb200jk.so.  Raises if it has not been built -- never falls back to CPU code."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise ImportError(
                f"{LIB_PATH} not found: build it with `python -m psi4_b200.build` (or __graft_entry__.build()). "
                "psi4_b200 has no CPU fallback.")
        L = ct.CDLL(LIB_PATH)
        for name, (res, args) in SIGNATURES.items():
            fn = getattr(L, name)  # AttributeError if the symbol is missing
            fn.restype = res
            fn.argtypes = args
        _lib = L
    return _lib


def _d(a):
    return a.ctypes.data_as(_dp)


def _ptr_array(arrs):
    if arrs is None:
        return None
    return (_dp * len(arrs))(*[_d(a) for a in arrs])


class Engine:
    """Thin owner of one b200jk_t handle."""

    def __init__(self, ngpu: int = 1, devices=None, *, rank=None, world=None, device=None, nccl_id: bytes | None = None):
        self.L = load()
        self.h = ct.c_void_p()
        if rank is None:
            devs = None if devices is None else (ct.c_int * ngpu)(*devices)
            rc = self.L.b200jk_create(ct.byref(self.h), ngpu, devs)
        else:
            buf = ct.create_string_buffer(nccl_id, NCCL_ID_BYTES) if nccl_id else None
            rc = self.L.b200jk_create_rank(ct.byref(self.h), int(device or 0), int(rank), int(world), buf)
        self._check(rc)
        self.nbf = self.naux = 0

    @staticmethod
    def nccl_unique_id() -> bytes:
        buf = ct.create_string_buffer(NCCL_ID_BYTES)
        rc = load().b200jk_nccl_unique_id(buf)
        if rc:
            raise B200JKError(rc, "ncclGetUniqueId failed")
        return buf.raw

    def _check(self, rc):
        if rc:
            msg = self.L.b200jk_last_error(self.h) if self.h else b"creation failed"
            raise B200JKError(rc, (msg or b"").decode())

    def close(self):
        if getattr(self, "h", None):
            self.L.b200jk_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def set_layout(self, nbf, naux, small_skips, big_skips, fun_index):
        ss = np.ascontiguousarray(small_skips, dtype=np.uintp)
        bs = np.ascontiguousarray(big_skips, dtype=np.uintp)
        fi = np.ascontiguousarray(fun_index, dtype=np.uintp).ravel()
        self._check(self.L.b200jk_set_layout(self.h, nbf, naux, ss.ctypes.data_as(_szp), bs.ctypes.data_as(_szp),
                                             fi.ctypes.data_as(_szp)))
        self.nbf, self.naux = int(nbf), int(naux)
        self._small_skips, self._big_skips = ss, bs
        # symm_big_skips_ (dfhelper.cc:413-416): offsets of the symmetric-packed blocks fit_rows takes
        mi = np.triu(fi.reshape(int(nbf), int(nbf)) > 0).sum(axis=1)
        self._symm_big_skips = np.concatenate(([0], np.cumsum(mi * int(naux)))).astype(np.uintp)

    def upload(self, which, packed):
        p = np.ascontiguousarray(packed, dtype=np.float64)
        if p.size != int(self._big_skips[self.nbf]):
            raise B200JKError(1, f"packed tensor has {p.size} doubles, layout says {int(self._big_skips[self.nbf])}")
        self._check(self.L.b200jk_upload(self.h, which, _d(p)))

    def upload_rows(self, which, m0, m1, rows):
        p = np.ascontiguousarray(rows, dtype=np.float64)
        need = int(self._big_skips[m1] - self._big_skips[m0])
        if p.size != need:
            raise B200JKError(1, f"row block has {p.size} doubles, layout says {need}")
        self._check(self.L.b200jk_upload_rows(self.h, which, m0, m1, _d(p)))

    def set_metric(self, metric):
        """naux x naux metric power for fit_rows (None: no contraction)."""
        if metric is None:
            self._check(self.L.b200jk_set_metric(self.h, None))
            return
        m = np.ascontiguousarray(metric, dtype=np.float64)
        assert m.shape == (self.naux, self.naux)
        self._check(self.L.b200jk_set_metric(self.h, _d(m)))

    def fit_rows(self, which, m0, m1, sym_rows):
        """Unfitted symmetric-packed rows [m0, m1) -> fitted, mirrored rows of tensor `which` (on the device)."""
        p = np.ascontiguousarray(sym_rows, dtype=np.float64)
        need = int(self._symm_big_skips[m1] - self._symm_big_skips[m0])
        if p.size < need:  # psi4 hands over one reused block buffer sized for its largest block (dfhelper.cc:553-585)
            raise B200JKError(1, f"symmetric-packed block has {p.size} doubles, layout needs {need}")
        self._check(self.L.b200jk_fit_rows(self.h, which, m0, m1, _d(p)))

    def fit_stats(self) -> dict:
        ms, fl = ct.c_double(), ct.c_double()
        self._check(self.L.b200jk_fit_stats(self.h, ct.byref(ms), ct.byref(fl)))
        return {"ms_gemm": ms.value, "flops": fl.value,
                "tflops": fl.value / (ms.value * 1e-3) / 1e12 if ms.value else 0.0}

    def fill_synthetic(self, which, seed, amp):
        a = np.ascontiguousarray(amp, dtype=np.float64)
        assert a.shape == (self.nbf, self.nbf)
        self._check(self.L.b200jk_fill_synthetic(self.h, which, ct.c_uint64(seed), _d(a)))

    def download_rows(self, which, m, q0, q1):
        sp = int(self._small_skips[m])
        out = np.zeros((q1 - q0, sp))
        self._check(self.L.b200jk_download_rows(self.h, which, m, q0, q1, _d(out)))
        return out

    def _outputs(self, tag, nmat, reuse):
        n = self.nbf
        if not reuse:
            return [np.empty((n, n)) for _ in range(nmat)]
        cache = self.__dict__.setdefault("_out_cache", {})
        key = (tag, nmat, n)
        if key not in cache:
            cache[key] = [np.zeros((n, n)) for _ in range(nmat)]  # zeros: pages touched once, here
            for a in cache[key]:
                self.register_host(a)  # persistent result matrices: the engine DMAs straight into them
        return cache[key]

    def register_host(self, arr: np.ndarray):
        """b200jk_register_host: page-lock a persistent caller array (D, J, K, wK) for direct DMA.  The array must
        outlive the engine or be unregistered first."""
        self._check(self.L.b200jk_register_host(self.h, ct.c_void_p(arr.ctypes.data), arr.nbytes))
        self.__dict__.setdefault("_registered", []).append(arr)  # keep it alive

    def unregister_host(self, arr: np.ndarray):
        self._check(self.L.b200jk_unregister_host(self.h, ct.c_void_p(arr.ctypes.data)))

    def compute(self, Cl, Cr, D, do_J=True, do_K=True, do_wK=False, reuse_outputs=False, fetch=True):
        """Host-operand build.  Cl/Cr: lists of (nbf, nocc_i) arrays (Cr None => lr_symmetric);
        D: list of (nbf,nbf).  Returns (J, K, wK) lists (None where untasked).

        reuse_outputs: write into persistent per-engine result matrices, as psi4's JK does (J_/K_ are allocated
        once in JK::allocate_JK, jk.cc:355-389, and overwritten by every compute(); callers re-fetch J()[i],
        jk.h:149-159).  Otherwise fresh arrays are returned.
        fetch=False (rank mode, rank != 0 only): take part in the build and the all-reduce, bring nothing home."""
        n = self.nbf
        nmat = len(Cl) if Cl is not None else len(D)
        if not n:
            raise B200JKError(1, "b200jk_compute before b200jk_set_layout")
        Cl_ = None if Cl is None else [np.ascontiguousarray(c, dtype=np.float64).reshape(n, -1) for c in Cl]
        Cr_ = None if Cr is None else [np.ascontiguousarray(c, dtype=np.float64).reshape(n, -1) for c in Cr]
        D_ = None if D is None else [np.ascontiguousarray(d, dtype=np.float64).reshape(n, n) for d in D]
        nocc = (ct.c_int * nmat)(*([c.shape[1] for c in Cl_] if Cl_ is not None else [0] * nmat))
        J = self._outputs("J", nmat, reuse_outputs) if do_J and fetch else None
        K = self._outputs("K", nmat, reuse_outputs) if do_K and fetch else None
        wK = self._outputs("wK", nmat, reuse_outputs) if do_wK and fetch else None
        rc = self.L.b200jk_compute(self.h, nmat, _ptr_array(Cl_), _ptr_array(Cr_), nocc, _ptr_array(D_),
                                   _ptr_array(J), _ptr_array(K), _ptr_array(wK), int(do_J), int(do_K), int(do_wK))
        self._check(rc)
        return J, K, wK

    # -- device-resident operands (kernel-only timing) --
    def dev_alloc(self, nbytes) -> int:
        p = ct.c_void_p()
        self._check(self.L.b200jk_dev_alloc(self.h, nbytes, ct.byref(p)))
        return p.value

    def dev_free(self, p):
        self._check(self.L.b200jk_dev_free(self.h, ct.c_void_p(p)))

    def dev_put(self, arr) -> int:
        a = np.ascontiguousarray(arr, dtype=np.float64)
        p = self.dev_alloc(a.nbytes)
        self._check(self.L.b200jk_dev_copy(self.h, ct.c_void_p(p), a.ctypes.data_as(ct.c_void_p), a.nbytes, 1))
        return p

    def dev_get(self, p, shape):
        out = np.empty(shape)
        self._check(self.L.b200jk_dev_copy(self.h, out.ctypes.data_as(ct.c_void_p), ct.c_void_p(p), out.nbytes, 2))
        return out

    def compute_device(self, dCl, dCr, nocc, dD, dJ, dK, dwK, do_J=True, do_K=True, do_wK=False):
        nmat = len(nocc)

        def arr(ps):
            if ps is None:
                return None
            return ct.cast((ct.c_void_p * nmat)(*ps), _dpp)

        no = (ct.c_int * nmat)(*nocc)
        self._check(self.L.b200jk_compute_device(self.h, nmat, arr(dCl), arr(dCr), no, arr(dD), arr(dJ), arr(dK),
                                                 arr(dwK), int(do_J), int(do_K), int(do_wK)))

    # -- DF-JK gradient intermediates (scfgrad/jk_grad.cc) --
    def grad_begin(self, C, Dt, Jm12):
        """C: [Ca_occ] (restricted) or [Ca_occ, Cb_occ]; Dt: total density; Jm12: the J^-1/2 the tensor was fitted with."""
        n = self.nbf
        Cs = [np.ascontiguousarray(c, dtype=np.float64).reshape(n, -1) for c in C]
        nocc = (ct.c_int * len(Cs))(*[c.shape[1] for c in Cs])
        Dt = np.ascontiguousarray(Dt, dtype=np.float64)
        J = np.ascontiguousarray(Jm12, dtype=np.float64)
        assert Dt.shape == (n, n) and J.shape == (self.naux, self.naux)
        self._check(self.L.b200jk_grad_begin(self.h, len(Cs), _ptr_array(Cs), nocc, _d(Dt), _d(J)))

    def grad_vectors(self):
        d, V = np.empty(self.naux), np.empty((self.naux, self.naux))
        self._check(self.L.b200jk_grad_vectors(self.h, _d(d), _d(V)))
        return d, V

    def grad_rows(self, a0, a1):
        out = np.empty((a1 - a0, self.nbf, self.nbf))
        self._check(self.L.b200jk_grad_rows(self.h, a0, a1, _d(out)))
        return out

    def grad_end(self):
        self._check(self.L.b200jk_grad_end(self.h))

    def matrix_power(self, A, alpha, cutoff, with_info=False):
        """Matrix::power (libmints/matrix.cc:2370-2424) on the device; returns the powered matrix (and, with_info, the
        number of eigenvalues kept and the device milliseconds)."""
        a = np.ascontiguousarray(A, dtype=np.float64)
        assert a.ndim == 2 and a.shape[0] == a.shape[1]
        out = np.empty_like(a)
        kept, ms = ct.c_int(), ct.c_double()
        self._check(self.L.b200jk_matrix_power(self.h, a.shape[0], _d(a), float(alpha), float(cutoff), _d(out),
                                               ct.byref(kept), ct.byref(ms)))
        return (out, kept.value, ms.value) if with_info else out

    def stats(self) -> dict:
        s = Stats()
        self._check(self.L.b200jk_get_stats(self.h, ct.byref(s)))
        return s.as_dict()

    def hbm_estimate(self, max_nocc, do_wK=False) -> int:
        v = ct.c_uint64()
        self._check(self.L.b200jk_hbm_estimate(self.h, max_nocc, int(do_wK), ct.byref(v)))
        return v.value

    def set_work_budget(self, nbytes):
        self._check(self.L.b200jk_set_work_budget(self.h, nbytes))

    def set_half(self, arm="auto", moduli=0):
        """Arm of the half transform: "auto" | "dmma" | "i8" (INT8 tensor cores by residues, `moduli` 6..13, default 12)."""
        self._check(self.L.b200jk_set_half(self.h, {"auto": 0, "dmma": 1, "i8": 2}[arm], int(moduli)))

    def set_kgemm(self, arm="auto", moduli=0):
        """Arm of the K GEMM: "auto" | "dmma" (FP64 tensor pipe) | "i8" (INT8 tensor cores by residues, `moduli` 6..13)."""
        self._check(self.L.b200jk_set_kgemm(self.h, {"auto": 0, "dmma": 1, "i8": 2}[arm], int(moduli)))

    def fp64_peak(self, kind, seconds=1.0) -> dict:
        v = (ct.c_double * 4)()
        self._check(self.L.b200jk_fp64_peak(self.h, kind, float(seconds), v))
        return {"burst_tflops": v[0], "sustained_tflops": v[1], "burst_sm_mhz": v[2], "sustained_sm_mhz": v[3]}
