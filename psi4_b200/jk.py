"""Host-side mirror of psi4's JK / MemDFJK interface over the B200 engine.

Same names, argument meaning and error behaviour as the reference classes, so a psi4 user's
code (and the reference's own tests, e.g. tests/pytests/test_dfjk.py) reads unchanged:

    jk = MemDFJK(dfh, tensors...)      # reference: JK.build_JK(primary, aux)  (libfock/jk.cc:72-234)
    jk.initialize()                     # MemDFJK::preiterations               (libfock/MemDFJK.cc:71-96)
    jk.C_left_add(C); jk.C_right_add(C)
    jk.compute()                        # JK::compute                           (libfock/jk.cc:595-681)
    jk.J(), jk.K(), jk.wK(), jk.D()

Matrices are numpy (nbf, n) float64 arrays -- the C1 case of psi4's SharedMatrix
(MemDFJK::C1() is true, jk.h:1148, so the reference also works in the C1 AO basis).
All arithmetic of the build runs in libb200jk.so; there is no CPU path here.
"""
from __future__ import annotations

import numpy as np

from . import lib as _lib
from .dfhelper import DFHelper


class PsiException(RuntimeError):
    """Stand-in for PSIEXCEPTION (libpsi4util/exception.h:48)."""


class JK:
    """State and driver of libfock/jk.h:232-589 (C1 only)."""

    def __init__(self, nbf: int):
        # JK::common_init, jk.cc:248-277
        self.nbf_ = int(nbf)
        self.print_ = 1
        self.debug_ = 0
        self.bench_ = 0
        self.memory_ = 32000000
        self.omp_nthread_ = 1
        self.cutoff_ = 1.0e-12
        self.do_J_ = True
        self.do_K_ = True
        self.do_wK_ = False
        self.wcombine_ = False
        self.lr_symmetric_ = False
        self.omega_ = 0.0
        self.omega_alpha_ = 1.0
        self.omega_beta_ = 0.0
        self.C_left_: list[np.ndarray] = []
        self.C_right_: list[np.ndarray] = []
        self.D_: list[np.ndarray] = []
        self.J_: list[np.ndarray] = []
        self.K_: list[np.ndarray] = []
        self.wK_: list[np.ndarray] = []

    # ---- knobs, export_fock.cc:58-80 ----
    def set_print(self, v): self.print_ = int(v)
    def set_debug(self, v): self.debug_ = int(v)
    def set_bench(self, v): self.bench_ = int(v)
    def set_cutoff(self, v): self.cutoff_ = float(v)
    def get_cutoff(self): return self.cutoff_
    def set_memory(self, v): self.memory_ = int(v)
    def set_omp_nthread(self, v): self.omp_nthread_ = int(v)
    def set_do_J(self, v): self.do_J_ = bool(v)
    def set_do_K(self, v): self.do_K_ = bool(v)
    def set_do_wK(self, v): self.do_wK_ = bool(v)
    def get_do_wK(self): return self.do_wK_
    def set_omega(self, v): self.omega_ = float(v)
    def get_omega(self): return self.omega_
    def set_omega_alpha(self, v): self.omega_alpha_ = float(v)
    def get_omega_alpha(self): return self.omega_alpha_
    def set_omega_beta(self, v): self.omega_beta_ = float(v)
    def get_omega_beta(self): return self.omega_beta_
    def get_wcombine(self): return self.wcombine_

    def set_wcombine(self, wcombine):
        # jk.cc:683-688
        self.wcombine_ = bool(wcombine)
        if wcombine:
            raise PsiException("To combine exchange terms, use MemDFJK\n")

    # ---- operand lists, export_fock.cc:83-95 ----
    def C_left(self): return self.C_left_
    def C_right(self): return self.C_right_
    def C_clear(self):
        self.C_left_.clear()
        self.C_right_.clear()
    def C_add(self, C):
        self.C_left_.append(C)
        self.C_right_.append(C)
    def C_left_add(self, C): self.C_left_.append(C)
    def C_right_add(self, C): self.C_right_.append(C)
    def J(self): return self.J_
    def K(self): return self.K_
    def wK(self): return self.wK_
    def D(self): return self.D_

    def initialize(self): self.preiterations()
    def finalize(self): self.postiterations()
    def basisset(self): return getattr(self, "primary_", None)  # jk.h:417

    def computed_shells_per_iter(self, n_let: str | None = None):
        """jk.cc:239-246.  Only the integral-direct algorithms tally shell n-lets (DirectJK.cc:955-957,
        CompositeJK.cc:320-338); MemDFJK never fills the map, so it is empty here as well."""
        tally = self.__dict__.setdefault("computed_shells_per_iter_", {})
        return tally if n_let is None else tally[n_let]

    @staticmethod
    def build_JK(primary, auxiliary, do_wK: bool = False, doubles: int = 0, scf_type: str = "MEM_DF", **options):
        """JK::build_JK (jk.cc:72-234, export_fock.cc:50-57) for psi4_b200 basis sets (integrals.BasisSet).
        scf_type "MEM_DF": a MemDFJK over the B200 engine (jk.cc:143-150).  scf_type "DF": the automatic choice of
        jk.cc:206-229 -- the exact MEM_DF estimate from the Schwarz mask is compared with `doubles`; the reference then
        falls back to DISK_DF, which does not exist behind the engine (its answer to "does not fit" is more Q shards),
        so that case raises.  options: cutoff, condition, ngpu, omega, fit_on_device."""
        from . import scf
        from .integrals import MintsHelper

        scf_type = scf_type.upper()
        if scf_type not in ("MEM_DF", "DF"):
            raise PsiException("JK::build_JK: only SCF_TYPE MEM_DF (or DF resolving to it) is served by the B200 engine")
        if scf_type == "DF":
            est = DFHelper(primary.nbf(), auxiliary.nbf())
            est.set_schwarz_cutoff(options.get("cutoff", 1.0e-12))
            est.set_do_wK(do_wK)
            est.set_Qshell_max(max(auxiliary.shell_nfunction(s) for s in range(auxiliary.nshell())))
            est.prepare_sparsity(MintsHelper(primary.molecule(), primary).schwarz_function_maxima())
            need = est.get_core_size(1)
            if not need < doubles:  # jk.cc:213
                raise PsiException(f"MemDFJK Memory: AOs need {need * 8 / 1024.0 ** 3:.3f} GiB; user supplied "
                                   f"{doubles * 8 / 1024.0 ** 3:.3f} GiB.  The reference would switch to DiskDFJK here "
                                   "(jk.cc:219-225); behind the B200 engine shard the tensor over more GPUs instead")
        jk = scf.build_jk(primary.molecule(), primary, auxiliary, do_wK=do_wK, **options)
        if doubles:
            jk.set_memory(doubles)
        return jk

    def compute(self):
        """JK::compute, jk.cc:595-681 (C1 branch: no USO2AO/AO2USO work when nirrep == 1)."""
        if len(self.C_left_) and not len(self.C_right_):
            self.lr_symmetric_ = True
            C_right = self.C_left_
        else:
            self.lr_symmetric_ = False
            C_right = self.C_right_
        if len(C_right) != len(self.C_left_):
            raise PsiException("JK: C_left/C_right irrep mismatch!")
        Cl, Cr = [], []
        for a, b in zip(self.C_left_, C_right):
            a = np.ascontiguousarray(a, dtype=np.float64)
            b = np.ascontiguousarray(b, dtype=np.float64)
            if a.ndim != 2 or a.shape[0] != self.nbf_ or b.ndim != 2 or b.shape[0] != self.nbf_:
                raise PsiException("JK: Input orbital irrep mismatch!")
            if a.shape[1] != b.shape[1]:
                raise PsiException("JK: C_left/C_right MO zip index size mismatch!")  # jk.cc:612-615
            Cl.append(a)
            Cr.append(b)
        # compute_D, jk.cc:314-354 -- O(N^2 o) host GEMM exactly where the reference does it
        self.D_ = [a @ b.T for a, b in zip(Cl, Cr)]
        self._Cl, self._Cr = Cl, (None if self.lr_symmetric_ else Cr)
        self.compute_JK()

    # pure virtuals of jk.h:341-345,394-398
    def preiterations(self): raise NotImplementedError
    def compute_JK(self): raise NotImplementedError
    def postiterations(self): pass
    def C1(self): return True
    def name(self): raise NotImplementedError
    def memory_estimate(self): raise NotImplementedError
    def print_header(self): raise NotImplementedError


class MemDFJK(JK):
    """libfock/jk.h:1124-1212 + MemDFJK.cc over the CUDA engine.

    The constructor takes what MemDFJK's DFHelper would have produced on the host after
    prepare_sparsity()/prepare_AO_core(): the tables (a psi4_b200.DFHelper) and the packed
    tensors.  From basis sets use psi4_b200.scf.build_jk(), the analogue of JK::build_JK.
    """

    def __init__(self, dfh: DFHelper, Ppq=None, m1Ppq=None, wPpq=None, *, ngpu: int = 1, devices=None,
                 rank=None, world=None, device=None, nccl_id=None, synthetic=None, unfitted=None, provider=None,
                 fit_block: int = 64):
        super().__init__(dfh.nbf_)
        if not dfh.sparsity_prepared_:
            raise PsiException("MemDFJK: DFHelper sparsity not prepared")
        self.dfh_ = dfh
        self.condition_ = 1.0e-12  # jk.h:1142
        self._Ppq, self._m1Ppq, self._wPpq = Ppq, m1Ppq, wPpq
        self._synthetic = synthetic  # (seed, amp) -> device-side fill, bench / large tests only
        # unfitted: tensor id -> (symmetric-packed unfitted integrals, metric power or None) -> fitted on the device;
        # the older triple (sym, metric, rows per block) for the Ppq tensor alone is still understood
        if unfitted is not None and not isinstance(unfitted, dict):
            sym, metric, fit_block = unfitted
            unfitted = {_lib.TENSOR_PPQ: (sym, metric)}
        self._unfitted = unfitted
        self._fit_block = int(fit_block)
        # provider(cutoff, condition, omega, do_wK) -> the host part of DFHelper::initialize (scf.DFTensors), run by
        # preiterations() with the knob values current THEN (MemDFJK.cc:71-96), not with those of construction time
        self._provider = provider
        self._engine_args = dict(ngpu=ngpu, devices=devices, rank=rank, world=world, device=device, nccl_id=nccl_id)
        self.engine: _lib.Engine | None = None

    def dfh(self): return self.dfh_
    def name(self): return "MemDFJK"
    def set_condition(self, c): self.condition_ = float(c)

    def set_wcombine(self, wcombine):
        # dfhelper.h:164-169: hard-disabled in the reference
        if wcombine:
            raise PsiException("MemDFJK: wcombine is not supported (disabled in the reference, dfhelper.h:164-169)")
        self.wcombine_ = False

    def memory_estimate(self) -> int:
        """MemDFJK.cc:65-69 -> DFHelper::get_core_size (doubles)."""
        self.dfh_.set_do_wK(self.do_wK_)
        return self.dfh_.get_core_size(self.omp_nthread_)

    def preiterations(self):
        """MemDFJK.cc:71-96: configure, then move the in-core tensors into HBM (Q-sharded)."""
        if self.engine is not None:
            return
        self.engine = _lib.Engine(**self._engine_args)
        if self._provider is not None:
            if self.do_wK_ and not self.omega_ > 0.0:
                raise PsiException("MemDFJK: wK tasked but omega is not set (JK::set_omega)")
            # the metric powers (Matrix::power, the O(naux^3) step of the setup) run on the device too
            t = self._provider(self.cutoff_, self.condition_, self.omega_, self.do_wK_, power=self.engine.matrix_power)
            self.dfh_ = t.dfh
            self._Ppq, self._m1Ppq, self._wPpq, self._unfitted = t.Ppq, t.m1Ppq, t.wPpq, t.unfitted
            self.mints_ = t.mints
            self.Jm12_ = t.Jm12
        d = self.dfh_
        d.set_do_wK(self.do_wK_)
        self.engine.set_layout(d.nbf_, d.naux_, d.small_skips_, d.big_skips_, d.schwarz_fun_index_)
        if self._synthetic is not None:
            seed, amp = self._synthetic
            self.engine.fill_synthetic(_lib.TENSOR_PPQ, seed, amp)
        elif self._unfitted is not None:
            # on-device fitting: the p-blocked loop of prepare_AO_core / prepare_AO_wK_core (dfhelper.cc:566-585,
            # :660-692) with the metric contraction + mirror copy (contract_metric_AO_core_symm, :1653-1678) done by
            # the engine; wPpq_ takes no metric (:688-692), m1Ppq_ the full inverse (:642-650)
            if self.do_wK_ and not all(w in self._unfitted for w in (_lib.TENSOR_M1PPQ, _lib.TENSOR_WPPQ)):
                raise PsiException("MemDFJK: do_wK requires the m1Ppq and wPpq tensors")
            for which in sorted(self._unfitted):
                if which != _lib.TENSOR_PPQ and not self.do_wK_:
                    continue
                sym, metric = self._unfitted[which]
                self.engine.set_metric(metric)
                for m0 in range(0, d.nbf_, self._fit_block):
                    m1 = min(d.nbf_, m0 + self._fit_block)
                    self.engine.fit_rows(which, m0, m1, sym[int(d.symm_big_skips_[m0]):int(d.symm_big_skips_[m1])])
            self._unfitted = None
        else:
            if self._Ppq is None:
                raise PsiException("MemDFJK: no Ppq tensor supplied")
            self.engine.upload(_lib.TENSOR_PPQ, self._Ppq)
            if self.do_wK_:
                # dfhelper.cc:199-206: wK needs the two extra in-core tensors
                if self._m1Ppq is None or self._wPpq is None:
                    raise PsiException("MemDFJK: do_wK requires the m1Ppq and wPpq tensors")
                self.engine.upload(_lib.TENSOR_M1PPQ, self._m1Ppq)
                self.engine.upload(_lib.TENSOR_WPPQ, self._wPpq)
        # the host copies may be released now (the engine does not retain host pointers)
        self._Ppq = self._m1Ppq = self._wPpq = None

    def max_nocc(self) -> int:
        # MemDFJK.cc:133-139
        return max([c.shape[1] for c in self._Cl], default=0)

    def compute_JK(self):
        """MemDFJK.cc:97-111: zero(), DFHelper::build_JK, hermitivitize(wK) -- all inside the engine."""
        if self.engine is None:
            raise PsiException("MemDFJK: compute() before initialize()")
        if not len(self._Cl):
            self.J_, self.K_, self.wK_ = [], [], []
            return
        try:
            J, K, wK = self.engine.compute(self._Cl, self._Cr, self.D_, self.do_J_, self.do_K_, self.do_wK_,
                                           reuse_outputs=True)
        except _lib.B200JKError as e:
            raise PsiException(str(e)) from e
        n = self.nbf_
        zeros = lambda: [np.zeros((n, n)) for _ in self._Cl]  # noqa: E731  (allocate_JK, jk.cc:355-389)
        self.J_ = J if J is not None else zeros()
        self.K_ = K if K is not None else zeros()
        self.wK_ = wK if wK is not None else zeros()

    def postiterations(self):
        if self.engine is not None:
            self.engine.close()
            self.engine = None

    def stats(self) -> dict:
        return self.engine.stats()

    def print_header(self, out=None) -> str:
        """MemDFJK.cc:113-132, extended with the device placement."""
        d = self.dfh_
        lines = ["  ==> MemDFJK: Density-Fitted J/K Matrices <==", ""]
        lines.append("    J tasked:           %11s" % ("Yes" if self.do_J_ else "No"))
        lines.append("    K tasked:           %11s" % ("Yes" if self.do_K_ else "No"))
        lines.append("    wK tasked:          %11s" % ("Yes" if self.do_wK_ else "No"))
        if self.do_wK_:
            lines.append("    Omega:              %11.3E" % self.omega_)
        lines.append("    OpenMP threads:     %11d" % self.omp_nthread_)
        lines.append("    Memory [MiB]:       %11d" % ((self.memory_ * 8) // (1024 * 1024)))
        lines.append("    Algorithm:          %11s" % "HBM-Core")
        lines.append("    Schwarz Cutoff:     %11.0E" % self.cutoff_)
        lines.append("    Mask sparsity (%%):  %11.4f" % (100.0 * d.ao_sparsity()))
        lines.append("    Fitting Condition:  %11.0E" % self.condition_)
        if self.engine is not None:
            st = self.engine.stats()
            lines.append("    GPUs (Q shards):    %11d" % st["n_shards"])
            lines.append("    HBM tensors [MiB]:  %11d" % (st["hbm_tensor_bytes"] // (1024 * 1024)))
        text = "\n".join(lines) + "\n"
        if self.print_ and out is not None:
            out.write(text)
        return text
