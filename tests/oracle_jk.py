"""Test-only JK object backed by the CPU oracle (oracle/dfjk_oracle.c) with the psi4 JK method surface, so the
same SCF driver can run on the restatement of the reference and on the CUDA engine."""
import numpy as np

import dfjk_oracle as oracle
from psi4_b200.jk import JK


class OracleJK(JK):
    def __init__(self, dfh, Ppq, m1Ppq=None, wPpq=None, impl="port"):
        super().__init__(dfh.nbf_)
        self.impl = impl  # "port": the C restatement; "ref": the reference's own functions (oracle/_ref)
        self.dfh_ = dfh
        self.sp = oracle.Sparsity(dfh.keep_.astype(np.uint8), dfh.naux_)
        self.Ppq, self.m1Ppq, self.wPpq = Ppq, m1Ppq, wPpq

    def name(self):
        return "OracleMemDFJK"

    def preiterations(self):
        pass

    def compute_JK(self):
        J, K, wK, _ = oracle.build_JK(self.sp, self.Ppq, self._Cl, self._Cr, D=self.D_, do_J=self.do_J_, do_K=self.do_K_,
                                      do_wK=self.do_wK_, m1Ppq=self.m1Ppq, wPpq=self.wPpq, impl=self.impl)
        n = self.nbf_
        z = lambda: [np.zeros((n, n)) for _ in self._Cl]  # noqa: E731
        self.J_, self.K_, self.wK_ = J or z(), K or z(), wK or z()
