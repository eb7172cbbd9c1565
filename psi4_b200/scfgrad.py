"""Caller side of the DF-SCF gradient: host mirrors of psi4's DFJKGrad (scfgrad/jk_grad.cc) and of
the assembly in scfgrad/scf_grad.cc:115-310, over the engine's gradient entry points (b200jk_grad_*).

    jkg = DFJKGrad(jk, aux)            # reference: JKGrad::build_JKGrad(1, mints)     (jk_grad.cc:79-120)
    jkg.set_Ca(C); jkg.set_Cb(C); jkg.set_Da(D); jkg.set_Db(D); jkg.set_Dt(2 D)
    jkg.compute_gradient()             # DFJKGrad::compute_gradient                    (jk_grad.cc:175-293)
    jkg.gradients()["Coulomb"], ["Exchange"]

What runs where: every tensor contraction (c_A, (A|ij), the metric contractions, V_AB, the AO back-transform) runs in
libb200jk.so on the tensor already resident in HBM; the derivative integrals (A|B)^x and (A|mn)^x and their final dot
products stay on the host, as they stay with Libint2 in psi4.  This package has no analytic derivative-integral code:
the host front end differentiates its own integrals by a five-point central stencil in the nuclear coordinates (exact
integrals, O(h^4) stencil error, ~1e-10 at h = 2e-3 bohr) -- good for the small-molecule anchors the reference's tests
hold, not a production path.
"""
from __future__ import annotations

import numpy as np

from .integrals import BasisSet, MintsHelper, Molecule
from .jk import PsiException


def _displaced(mol: Molecule, atom: int, xyz: int, h: float) -> Molecule:
    x = mol.xyz.copy()
    x[atom, xyz] += h
    return Molecule(list(mol.symbols), x, mol.charge)


def derivative_integrals(mol: Molecule, primary_name: str, aux_name: str, h: float = 2.0e-3, puream=None):
    """Total nuclear-coordinate derivatives of S, T + V, (A|B) and (A|mn): arrays indexed [atom, xyz, ...].
    Five-point stencil f' = [f(-2h) - 8 f(-h) + 8 f(h) - f(2h)] / 12h of the host front end's integrals."""
    coef = {-2: 1.0 / 12.0, -1: -8.0 / 12.0, 1: 8.0 / 12.0, 2: -1.0 / 12.0}
    nat = len(mol.symbols)
    out = None
    for a in range(nat):
        for x in range(3):
            acc = None
            for k, c in coef.items():
                m2 = _displaced(mol, a, x, k * h)
                P, A = BasisSet.build(m2, primary_name, puream), BasisSet.build(m2, aux_name)
                mints = MintsHelper(m2, P)
                S, T, V = mints.one_electron()
                vals = [S, T + V, mints.metric(A), mints.three_center(A), np.array(m2.nuclear_repulsion())]
                acc = [c / h * v for v in vals] if acc is None else [p + c / h * v for p, v in zip(acc, vals)]
            if out is None:
                out = [np.zeros((nat, 3) + v.shape) for v in acc]
            for o, v in zip(out, acc):
                o[a, x] = v
    return dict(dS=out[0], dH=out[1], dAB=out[2], dAmn=out[3], dEnuc=out[4])


class DFJKGrad:
    """scfgrad/jk_grad.h: DFJKGrad over the engine.  `jk` is an initialized psi4_b200.MemDFJK (its tensor is resident);
    `Jm12` the metric power it was fitted with; `deriv` the derivative integrals (dict with dAB, dAmn)."""

    def __init__(self, jk, Jm12: np.ndarray, deriv: dict):
        if jk.engine is None:
            raise PsiException("DFJKGrad: the JK object is not initialized")
        self.jk, self.Jm12, self.deriv = jk, Jm12, deriv
        self.do_J_, self.do_K_, self.do_wK_ = True, True, False
        self.Ca_ = self.Cb_ = self.Da_ = self.Db_ = self.Dt_ = None
        self.gradients_ = {}
        self.block_rows = 32

    def set_Ca(self, C): self.Ca_ = C
    def set_Cb(self, C): self.Cb_ = C
    def set_Da(self, D): self.Da_ = D
    def set_Db(self, D): self.Db_ = D
    def set_Dt(self, D): self.Dt_ = D
    def set_do_J(self, v): self.do_J_ = bool(v)
    def set_do_K(self, v): self.do_K_ = bool(v)

    def set_do_wK(self, v):
        if v:
            raise PsiException("DFJKGrad: Exchange,LR gradients are not served by the B200 engine yet")

    def gradients(self): return self.gradients_

    def compute_gradient(self):
        """jk_grad.cc:175-293."""
        if not (self.do_J_ or self.do_K_):
            return
        if any(x is None for x in (self.Ca_, self.Cb_, self.Da_, self.Db_, self.Dt_)):
            raise PsiException("Occupation/Density not set")  # jk_grad.cc:178
        restricted = self.Ca_ is self.Cb_  # jk_grad.cc:302: pointer equality
        eng = self.jk.engine
        eng.grad_begin([self.Ca_] if restricted else [self.Ca_, self.Cb_], self.Dt_, self.Jm12)
        d, V = eng.grad_vectors()
        dAB, dAmn = self.deriv["dAB"], self.deriv["dAmn"]
        naux = dAB.shape[-1]
        nat = dAB.shape[0]
        if self.do_J_:
            # metric_grad "Coulomb" (mintshelper.cc:2462-2470) + build_Amn_x_terms (jk_grad.cc:1078-1088)
            self.gradients_["Coulomb"] = (-0.5 * np.einsum("axAB,A,B->ax", dAB, d, d)
                                          + np.einsum("axAmn,A,mn->ax", dAmn, d, self.Dt_, optimize=True))
        if self.do_K_:
            g = -0.5 * np.einsum("axAB,AB->ax", dAB, V)
            for a0 in range(0, naux, self.block_rows):  # the auxiliary-shell blocks of jk_grad.cc:993-1001
                a1 = min(naux, a0 + self.block_rows)
                Kmn = eng.grad_rows(a0, a1)
                g = g + np.einsum("axAmn,Amn->ax", dAmn[:, :, a0:a1], Kmn, optimize=True)
            self.gradients_["Exchange"] = g
        eng.grad_end()
        assert self.gradients_[next(iter(self.gradients_))].shape == (nat, 3)


def scf_gradient_rhf(rhf, jkgrad, deriv: dict) -> dict:
    """SCFDeriv::compute_gradient for RHF (scfgrad/scf_grad.cc:115-310): Nuclear + Core + Overlap + Coulomb + Exchange.
    `rhf` is a converged psi4_b200.scf.RHF; `jkgrad` anything with the DFJKGrad surface (the engine-backed class above or
    an oracle-backed one in the tests)."""
    nd = rhf.ndocc
    Cocc = np.ascontiguousarray(rhf.C[:, :nd])
    Da = Cocc @ Cocc.T
    Dt = 2.0 * Da
    eps = rhf.eps[:nd]
    W = 2.0 * (Cocc * eps) @ Cocc.T  # scf_grad.cc:205-218 (alpha + beta)
    terms = {
        "Nuclear": deriv["dEnuc"].copy(),
        "Core": np.einsum("axmn,mn->ax", deriv["dH"], Dt),                 # core_hamiltonian_grad(Dt)
        "Overlap": -np.einsum("axmn,mn->ax", deriv["dS"], W),              # overlap_grad(W) scaled by -1 (:219-220)
    }
    jkgrad.set_Ca(Cocc)
    jkgrad.set_Cb(Cocc)
    jkgrad.set_Da(Da)
    jkgrad.set_Db(Da)
    jkgrad.set_Dt(Dt)
    jkgrad.set_do_J(True)
    jkgrad.set_do_K(True)
    jkgrad.compute_gradient()
    g = jkgrad.gradients()
    terms["Coulomb"] = g["Coulomb"]
    terms["Exchange"] = -1.0 * g["Exchange"]  # scale(-alpha), alpha = 1 for Hartree-Fock (:266)
    terms["Total"] = sum(terms[k] for k in ("Nuclear", "Core", "Overlap", "Coulomb", "Exchange"))
    return terms
