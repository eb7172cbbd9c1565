"""Minimal DF-SCF driver around the JK interface: the caller side of the drop-in
boundary, written the way psi4's own tests drive a JK object (tests/psi4numpy/rhf/input.py:78-137).

  build_jk(mol, primary, aux)   analogue of JK::build_JK(primary, aux) with SCF_TYPE=MEM_DF
                                (libfock/jk.cc:143-150): runs the host part of DFHelper::initialize
                                (dfhelper.cc:149-215: metric power, Schwarz mask, (A|mn), fitting, packing)
                                and hands the packed tensor to the CUDA engine.
  RHF / UHF                     energy expressions and Fock builds of libscf_solver (rhf.cc:187-365,
                                uhf.cc:184-424) on top of jk.compute(); core guess + DIIS.

Any object with the JK method surface can be passed as `jk` (the tests inject an oracle-backed one to pin the
CPU restatement against the same published energies); the default is the B200 engine and nothing else.
"""
from __future__ import annotations

import numpy as np

from .dfhelper import DFHelper
from .integrals import BasisSet, MintsHelper, Molecule
from .jk import MemDFJK


def matrix_power(A: np.ndarray, alpha: float, cutoff: float) -> np.ndarray:
    """Matrix::power (libmints/matrix.cc:2370-2424): V diag(lambda^alpha) V^T, eigenvalues with
    |lambda| < cutoff*max|lambda| dropped when alpha < 0."""
    w, V = np.linalg.eigh(A)
    max_a = max(abs(w[0]), abs(w[-1]))
    out = np.zeros_like(w)
    for i, a in enumerate(w):
        if alpha < 0.0 and abs(a) < cutoff * max_a:
            out[i] = 0.0
        else:
            with np.errstate(all="ignore"):
                v = np.power(a, alpha)
            out[i] = v if np.isfinite(v) else 0.0
    return (V * out) @ V.T


class DFTensors:
    """Host part of DFHelper::initialize for the in-core STORE method (dfhelper.cc:149-215, :514-588)."""

    def __init__(self, mol: Molecule, primary: BasisSet, aux: BasisSet, cutoff: float = 1e-12, condition: float = 1e-10,
                 do_wK: bool = False, fit_on_device: bool = False, omega: float = 0.0, power=None):
        """power(A, alpha, cutoff): Matrix::power implementation; default the host one (numpy), the engine-backed JK
        passes the device one (b200jk_matrix_power)."""
        matrix_power_ = power or matrix_power
        if do_wK and not omega > 0.0:
            raise ValueError("do_wK needs omega > 0 (JK::set_omega)")
        mints = MintsHelper(mol, primary)
        self.mints = mints
        self.dfh = DFHelper(primary.nbf(), aux.nbf())
        self.dfh.set_schwarz_cutoff(cutoff)
        self.dfh.prepare_blocking([primary.shell_nfunction(s) for s in range(primary.nshell())],
                                  [aux.shell_nfunction(s) for s in range(aux.nshell())])  # :84-103
        self.dfh.prepare_sparsity(fun_max_vals=mints.schwarz_function_maxima())   # prepare_sparsity :299-420
        metric = mints.metric(aux)                                                 # prepare_metric :1462-1476
        self.Jm12 = matrix_power_(metric, -0.5, condition)                         # compute_metric :1491-1517
        Amn = mints.three_center(aux)                                              # :1284-1347
        self.Ppq = self.dense = self.m1Ppq = self.wPpq = None
        self.unfitted = None  # tensor id -> (symmetric-packed unfitted integrals, metric power or None)
        Jm1 = Wmn = None
        if do_wK:
            # prepare_AO_wK_core :589-699 -- m1Ppq_ = J^-1 (A|mn) (wmpower_ = -1.0), wPpq_ = (A|erf(omega r)/r|mn) unfitted
            self.dfh.set_do_wK(True)
            Jm1 = matrix_power_(metric, -1.0, condition)
            Wmn = mints.three_center(aux, omega)
        if fit_on_device:
            # hand the unfitted n >= m half to the engine (b200jk_fit_rows); metric contraction + mirror run on the GPU.
            # With wK: the same (A|mn) contracted with J^-1 for m1Ppq_ (:642-650, :678) and the erf-attenuated
            # integrals scattered + mirrored with no metric for wPpq_ (:688-692).
            sym = self.dfh.pack_symm(Amn)
            self.unfitted = {0: (sym, self.Jm12)}
            if do_wK:
                self.unfitted[1] = (sym, Jm1)
                self.unfitted[2] = (self.dfh.pack_symm(Wmn), None)
        else:
            # contract_metric_AO_core_symm :1653-1678  (B = J^-1/2 (A|mn)) on the host, then pack to pQq
            B = np.tensordot(self.Jm12, Amn, axes=([1], [0]))
            self.dense = B
            self.Ppq = self.dfh.pack(B)
            if do_wK:
                self.m1Ppq = self.dfh.pack(np.tensordot(Jm1, Amn, axes=([1], [0])))
                self.wPpq = self.dfh.pack(Wmn)


def build_jk(mol: Molecule, primary: BasisSet, aux: BasisSet, *, cutoff: float = 1e-12, condition: float = 1e-10,
             ngpu: int = 1, jk_factory=None, fit_on_device: bool = False, fit_block: int = 64, do_wK: bool = False,
             omega: float = 0.0):
    """JK::build_JK analogue.  jk_factory(dfh, Ppq) may construct another JK implementation (tests).
    fit_on_device: the engine contracts the metric itself (b200jk_set_metric / b200jk_fit_rows), fed in blocks of
    fit_block basis functions like the p-blocked loop of prepare_AO_core.

    As in the reference, the object is only CONFIGURED here: the integrals, the metric power and the packed tensors
    are produced by jk.initialize() (MemDFJK::preiterations pushes cutoff / condition / omega / do_wK into DFHelper and
    only then calls dfh_->initialize(), MemDFJK.cc:71-96), so set_cutoff / set_condition / set_omega / set_do_wK
    between build and initialize take effect -- the order psi4's own initialize_jk uses (scf_iterator.py:112-135).
    A jk_factory (an oracle-backed JK in the tests) takes ready-made host tensors and is therefore built eagerly."""
    if jk_factory is not None:
        t = DFTensors(mol, primary, aux, cutoff, condition, fit_on_device=False, do_wK=do_wK, omega=omega)
        jk = jk_factory(t.dfh, t.Ppq, t.m1Ppq, t.wPpq) if do_wK else jk_factory(t.dfh, t.Ppq)
        jk.mints_ = t.mints
    else:
        def provider(cutoff, condition, omega, do_wK, power=None):
            return DFTensors(mol, primary, aux, cutoff, condition, do_wK=do_wK, fit_on_device=fit_on_device, omega=omega,
                             power=power)

        # the tables of the build-time cutoff size nbf / memory_estimate(); initialize() rebuilds them from the knobs
        dfh = DFHelper(primary.nbf(), aux.nbf())
        dfh.set_schwarz_cutoff(cutoff)
        mints = MintsHelper(mol, primary)
        dfh.prepare_blocking([primary.shell_nfunction(s) for s in range(primary.nshell())],
                             [aux.shell_nfunction(s) for s in range(aux.nshell())])
        dfh.prepare_sparsity(fun_max_vals=mints.schwarz_function_maxima())
        jk = MemDFJK(dfh, ngpu=ngpu, provider=provider, fit_block=fit_block)
        jk.mints_ = mints
    if do_wK:
        jk.set_do_wK(True)
        jk.set_omega(omega)
    jk.set_cutoff(cutoff)
    if hasattr(jk, "set_condition"):
        jk.set_condition(condition)
    jk.primary_ = primary
    return jk


class _DIIS:
    def __init__(self, max_vecs=10):
        self.F, self.E, self.max = [], [], max_vecs

    def add(self, F, e):
        self.F.append(F.copy())
        self.E.append(e.copy())
        if len(self.F) > self.max:
            self.F.pop(0)
            self.E.pop(0)

    def extrapolate(self):
        n = len(self.F)
        B = -np.ones((n + 1, n + 1))
        B[n, n] = 0.0
        for i in range(n):
            for j in range(n):
                B[i, j] = np.vdot(self.E[i], self.E[j])
        scale = np.abs(B[:n, :n]).max()
        if scale > 0:
            B[:n, :n] /= scale
        rhs = np.zeros(n + 1)
        rhs[n] = -1.0
        c = np.linalg.lstsq(B, rhs, rcond=None)[0][:n]
        return sum(ci * Fi for ci, Fi in zip(c, self.F))


class DIIS:
    """psi4.p4util.solvers.DIIS (psi4/driver/p4util/solvers.py:203-322): same pruning policy, same scaled
    Pulay matrix and pseudo-inverse, so iteration tables produced with it are comparable digit for digit."""

    def __init__(self, max_vec: int = 6, removal_policy: str = "OLDEST"):
        self.error, self.state = [], []
        self.max_vec = max_vec
        self.removal_policy = removal_policy.upper()
        if self.removal_policy not in ("LARGEST", "OLDEST"):
            raise ValueError("DIIS: removal_policy must either be oldest or largest.")

    def add(self, state, error):
        self.error.append(np.array(error, copy=True))
        self.state.append(np.array(state, copy=True))

    def extrapolate(self):
        n = len(self.state)
        if n == 0:
            raise ValueError("DIIS: No previous vectors.")
        if n == 1:
            return self.state[0]
        if n > self.max_vec:
            pos = 0 if self.removal_policy == "OLDEST" else int(np.argmax([np.sqrt(np.mean(x ** 2)) for x in self.error]))
            del self.state[pos]
            del self.error[pos]
            n -= 1
        B = np.empty((n + 1, n + 1))
        B[-1, :] = 1
        B[:, -1] = 1
        B[-1, -1] = 0
        for i, e1 in enumerate(self.error):
            for j, e2 in enumerate(self.error):
                if j <= i:
                    B[i, j] = B[j, i] = np.vdot(e1, e2)
        resid = np.zeros(n + 1)
        resid[-1] = 1
        if np.any(np.diag(B)[:-1] <= 0.0):
            S = np.ones(n + 1)
        else:
            S = np.diag(B).copy()
            S[:-1] **= -0.5
            S[-1] = 1
        B *= S[:, None] * S
        ci = matrix_power(B, -1.0, 1.0e-12) @ resid
        ci *= S
        return sum(c * s for c, s in zip(ci[:-1], self.state))


def rhf_jk_loop(mol: Molecule, primary: BasisSet, jk, ndocc: int, maxiter=12, E_conv=1.0e-6, D_conv=1.0e-5):
    """The hand-written RHF loop of the reference's tests/psi4numpy/rhf/input.py:32-137, statement for statement,
    on a JK object: core guess, A = S^-1/2 (power(-0.5, 1e-16)), DIIS(max_vec=3, "largest").
    Returns the list of (energy, dE, dRMS) per iteration."""
    mints = getattr(jk, "mints_", None) or MintsHelper(mol, primary)
    S, T, V = mints.one_electron()
    H = T + V
    A = matrix_power(S, -0.5, 1.0e-16)

    def build_orbitals(diag):
        Fp = A.T @ diag @ A
        _, Cp = np.linalg.eigh(Fp)
        C = A @ Cp
        Cocc = np.ascontiguousarray(C[:, :ndocc])
        return C, Cocc, Cocc @ Cocc.T

    C, Cocc, D = build_orbitals(H)
    Enuc = mol.nuclear_repulsion()
    Eold = 0.0
    diis = DIIS(max_vec=3, removal_policy="largest")
    table = []
    for it in range(1, maxiter + 1):
        jk.C_left_add(Cocc)
        jk.compute()
        jk.C_clear()
        F = H + 2.0 * jk.J()[0] - jk.K()[0]
        e = A @ (F @ D @ S - S @ D @ F) @ A
        diis.add(F, e)
        E = float(np.vdot(F + H, D)) + Enuc
        drms = float(np.sqrt(np.mean(e ** 2)))
        table.append((E, E - Eold, drms))
        if abs(E - Eold) < E_conv and drms < D_conv:
            break
        Eold = E
        C, Cocc, D = build_orbitals(diis.extrapolate())
    return table


class RHF:
    """Closed-shell SCF on a JK object.  D = Cocc Cocc^T (no factor 2, rhf.cc:276-291); G = 2J - K
    (rhf.cc:215-242); E = Enuc + D.(H + F) (rhf.cc:308-365)."""

    def __init__(self, mol: Molecule, primary: BasisSet, jk, e_convergence=1e-10, d_convergence=1e-8, maxiter=100):
        self.mol, self.primary, self.jk = mol, primary, jk
        self.e_conv, self.d_conv, self.maxiter = e_convergence, d_convergence, maxiter
        mints = getattr(jk, "mints_", None) or MintsHelper(mol, primary)
        self.S, T, V = mints.one_electron()
        self.H = T + V
        self.Enuc = mol.nuclear_repulsion()
        self.ndocc = mol.nelectron() // 2
        self.X = matrix_power(self.S, -0.5, 1e-10)  # symmetric orthogonalisation (hf.cc:708)
        self.iterations = []

    def _diag(self, F):
        e, C2 = np.linalg.eigh(self.X @ F @ self.X)
        return e, self.X @ C2

    def compute_energy(self) -> float:
        jk = self.jk
        eps, C = self._diag(self.H)  # core guess
        Cocc = np.ascontiguousarray(C[:, : self.ndocc])
        diis = _DIIS()
        Eold = 0.0
        for it in range(self.maxiter):
            jk.C_clear()
            jk.C_left_add(Cocc)  # C_right empty => lr_symmetric (jk.cc:597-602)
            jk.compute()
            J, K = jk.J()[0], jk.K()[0]
            D = Cocc @ Cocc.T
            F = self.H + 2.0 * J - K
            E = self.Enuc + float(np.sum(D * (self.H + F)))
            grad = self.X @ (F @ D @ self.S - self.S @ D @ F) @ self.X
            drms = float(np.sqrt(np.mean(grad ** 2)))
            self.iterations.append((E, E - Eold, drms))
            if abs(E - Eold) < self.e_conv and drms < self.d_conv:
                break
            Eold = E
            diis.add(F, grad)
            Fx = diis.extrapolate() if it >= 1 else F
            eps, C = self._diag(Fx)
            Cocc = np.ascontiguousarray(C[:, : self.ndocc])
        self.energy, self.eps, self.C, self.F, self.D, self.J, self.K = E, eps, C, F, D, J, K
        self.eps = np.linalg.eigvalsh(self.X @ F @ self.X)
        return E


class UHF:
    """Unrestricted SCF: both spins in one jk.compute() (uhf.cc:195-200); Jtot = Ja + Jb;
    Fa = H + Jtot - Ka; E = Enuc + 1/2 [ (Da+Db).H + Da.Fa + Db.Fb ] (uhf.cc:361-424)."""

    def __init__(self, mol: Molecule, primary: BasisSet, jk, multiplicity=1, e_convergence=1e-10, d_convergence=1e-8,
                 maxiter=200):
        self.mol, self.primary, self.jk = mol, primary, jk
        self.e_conv, self.d_conv, self.maxiter = e_convergence, d_convergence, maxiter
        mints = getattr(jk, "mints_", None) or MintsHelper(mol, primary)
        self.S, T, V = mints.one_electron()
        self.H = T + V
        self.Enuc = mol.nuclear_repulsion()
        ne = mol.nelectron()
        self.na = (ne + multiplicity - 1) // 2
        self.nb = ne - self.na
        self.X = matrix_power(self.S, -0.5, 1e-10)
        self.iterations = []

    def compute_energy(self) -> float:
        X, S, H, jk = self.X, self.S, self.H, self.jk

        def diag(F):
            e, C2 = np.linalg.eigh(X @ F @ X)
            return X @ C2

        C = diag(H)
        Ca, Cb = np.ascontiguousarray(C[:, : self.na]), np.ascontiguousarray(C[:, : self.nb])
        da, db = _DIIS(), _DIIS()
        Eold = 0.0
        for it in range(self.maxiter):
            jk.C_clear()
            jk.C_left_add(Ca)
            jk.C_left_add(Cb)
            jk.compute()
            Jt = jk.J()[0] + jk.J()[1]
            Fa, Fb = H + Jt - jk.K()[0], H + Jt - jk.K()[1]
            Da, Db = Ca @ Ca.T, Cb @ Cb.T
            E = self.Enuc + 0.5 * float(np.sum((Da + Db) * H) + np.sum(Da * Fa) + np.sum(Db * Fb))
            ga = X @ (Fa @ Da @ S - S @ Da @ Fa) @ X
            gb = X @ (Fb @ Db @ S - S @ Db @ Fb) @ X
            drms = float(np.sqrt(0.5 * (np.mean(ga ** 2) + np.mean(gb ** 2))))
            self.iterations.append((E, E - Eold, drms))
            if abs(E - Eold) < self.e_conv and drms < self.d_conv:
                break
            Eold = E
            da.add(Fa, np.concatenate([ga.ravel(), gb.ravel()]))
            db.add(Fb, np.concatenate([ga.ravel(), gb.ravel()]))
            if it >= 1:
                Fa, Fb = da.extrapolate(), db.extrapolate()
            Ca = np.ascontiguousarray(diag(Fa)[:, : self.na])
            Cb = np.ascontiguousarray(diag(Fb)[:, : self.nb])
        self.energy = E
        return E


class ROHF:
    """Restricted open-shell SCF on a JK object, the call pattern of ROHF::form_G (libscf_solver/rohf.cc:935-966): ONE
    jk.compute() over C_left = [C_docc, C_socc] (two matrices with different column counts, C_right empty), then
    Ga = 2 J[0] + J[1] - (K[0] + K[1]),  Gb = 2 J[0] + J[1] - K[0].  Orbitals from the Guest-Saunders effective Fock
    matrix in the MO basis (rohf.cc:330-400: closed-open block Fb, open-virtual block Fa, average elsewhere);
    E = Enuc + 1/2 [ (Da+Db).H + Da.Fa + Db.Fb ]."""

    def __init__(self, mol: Molecule, primary: BasisSet, jk, multiplicity=3, e_convergence=1e-10, d_convergence=1e-8,
                 maxiter=200):
        self.mol, self.primary, self.jk = mol, primary, jk
        self.e_conv, self.d_conv, self.maxiter = e_convergence, d_convergence, maxiter
        mints = getattr(jk, "mints_", None) or MintsHelper(mol, primary)
        self.S, T, V = mints.one_electron()
        self.H = T + V
        self.Enuc = mol.nuclear_repulsion()
        ne = mol.nelectron()
        self.na = (ne + multiplicity - 1) // 2
        self.nb = ne - self.na
        self.X = matrix_power(self.S, -0.5, 1e-10)
        self.iterations = []

    def compute_energy(self) -> float:
        X, S, H, jk, nb, na = self.X, self.S, self.H, self.jk, self.nb, self.na
        _, C2 = np.linalg.eigh(X @ H @ X)  # core guess
        C = X @ C2
        diis = _DIIS()
        Eold = 0.0
        for it in range(self.maxiter):
            Cd, Cs = np.ascontiguousarray(C[:, :nb]), np.ascontiguousarray(C[:, nb:na])
            jk.C_clear()
            jk.C_left_add(Cd)
            jk.C_left_add(Cs)
            jk.compute()
            J, K = jk.J(), jk.K()
            G = 2.0 * J[0] + J[1]
            Fa, Fb = H + G - (K[0] + K[1]), H + G - K[0]
            Dc, Do = Cd @ Cd.T, Cs @ Cs.T
            Da, Db = Dc + Do, Dc
            E = self.Enuc + 0.5 * float(np.sum((Da + Db) * H) + np.sum(Da * Fa) + np.sum(Db * Fb))
            # effective Fock matrix in the current MO basis
            fa, fb = C.T @ Fa @ C, C.T @ Fb @ C
            feff = 0.5 * (fa + fb)
            feff[:nb, nb:na] = fb[:nb, nb:na]
            feff[nb:na, :nb] = fb[nb:na, :nb]
            feff[nb:na, na:] = fa[nb:na, na:]
            feff[na:, nb:na] = fa[na:, nb:na]
            # orbital gradient: the occupied-virtual type blocks (closed-open, closed-virtual, open-virtual)
            g = np.zeros_like(feff)
            g[:nb, nb:] = feff[:nb, nb:]
            g[nb:na, na:] = feff[nb:na, na:]
            g = g - g.T
            drms = float(np.sqrt(np.mean(g ** 2)))
            self.iterations.append((E, E - Eold, drms))
            if abs(E - Eold) < self.e_conv and drms < self.d_conv:
                break
            Eold = E
            # DIIS on the effective Fock matrix expressed in the fixed orthogonal AO basis (C^-1 = C^T S)
            Ci = C.T @ S
            F_ao, g_ao = Ci.T @ feff @ Ci, Ci.T @ g @ Ci
            diis.add(X @ F_ao @ X, X @ g_ao @ X)
            Fx = diis.extrapolate() if it >= 1 else X @ F_ao @ X
            _, C2 = np.linalg.eigh(Fx)
            C = X @ C2
        self.energy, self.C = E, C
        return E


class CUHF(UHF):
    """Constrained UHF (libscf_solver/cuhf.cc): the JK call pattern of UHF (both spin blocks in one jk.compute()), with
    the spin-polarisation part of the Fock matrices projected so that core-virtual mixing vanishes in the basis of the
    natural orbitals of the charge density (cuhf.cc:213-283).  Converges to the ROHF energy (tests/scf5 compares the
    CUHF runs with the ROHF references)."""

    def compute_energy(self) -> float:
        X, S, H, jk, na, nb = self.X, self.S, self.H, self.jk, self.na, self.nb
        Shalf = matrix_power(S, 0.5, 1e-10)

        def diag(F):
            _, C2 = np.linalg.eigh(X @ F @ X)
            return X @ C2

        C = diag(H)
        Ca, Cb = np.ascontiguousarray(C[:, :na]), np.ascontiguousarray(C[:, :nb])
        da, db = _DIIS(), _DIIS()
        Eold = 0.0
        for it in range(self.maxiter):
            jk.C_clear()
            jk.C_left_add(Ca)
            jk.C_left_add(Cb)
            jk.compute()
            Jt = jk.J()[0] + jk.J()[1]
            Fa, Fb = H + Jt - jk.K()[0], H + Jt - jk.K()[1]
            Da, Db = Ca @ Ca.T, Cb @ Cb.T
            E = self.Enuc + 0.5 * float(np.sum((Da + Db) * H) + np.sum(Da * Fa) + np.sum(Db * Fb))
            # natural orbitals of the charge density in the orthogonal basis: occupations 1 (core), 1/2 (active), 0 (virtual)
            occ, U = np.linalg.eigh(Shalf @ (0.5 * (Da + Db)) @ Shalf)
            U = U[:, np.argsort(-occ)]
            Fp = X @ (0.5 * (Fa + Fb)) @ X
            Fm = U.T @ (X @ (0.5 * (Fa - Fb)) @ X) @ U
            Fm[:nb, na:] = 0.0
            Fm[na:, :nb] = 0.0
            Fm = U @ Fm @ U.T
            Xi = Shalf  # orthogonal-basis Fock -> AO basis: F = S^1/2 F' S^1/2
            Fa, Fb = Xi @ (Fp + Fm) @ Xi, Xi @ (Fp - Fm) @ Xi
            ga = X @ (Fa @ Da @ S - S @ Da @ Fa) @ X
            gb = X @ (Fb @ Db @ S - S @ Db @ Fb) @ X
            drms = float(np.sqrt(0.5 * (np.mean(ga ** 2) + np.mean(gb ** 2))))
            self.iterations.append((E, E - Eold, drms))
            if abs(E - Eold) < self.e_conv and drms < self.d_conv:
                break
            Eold = E
            err = np.concatenate([ga.ravel(), gb.ravel()])
            da.add(Fa, err)
            db.add(Fb, err)
            if it >= 1:
                Fa, Fb = da.extrapolate(), db.extrapolate()
            Ca = np.ascontiguousarray(diag(Fa)[:, :na])
            Cb = np.ascontiguousarray(diag(Fb)[:, :nb])
        self.energy = E
        return E
