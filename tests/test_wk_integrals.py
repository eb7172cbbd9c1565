"""Range-separated exchange with REAL integrals (row a14 / a7 of SURVEY.md 8a): the host front end's
(A|erf(omega r)/r|mn) -- what IntegralFactory::erf_eri gives DFHelper::prepare_AO_wK_core (dfhelper.cc:589-699) --
and the three tensors Ppq_ / m1Ppq_ / wPpq_ built from it, driven through compute_wK (dfhelper.cc:3378-3438).

The reference holds no wK numbers that do not also need an XC functional (tests/dft-omega), so the integrals are
pinned by two exact identities instead:
  * convolution: erf(w r)/r is 1/r smeared with a normalised Gaussian of exponent w^2, and a solid-harmonic Gaussian
    r^l Y_lm exp(-c r^2) smeared that way is (c'/c)^(l+3/2) r^l Y_lm exp(-c' r^2) with c' = c w^2 / (c + w^2): the
    attenuated integrals equal plain Coulomb integrals over an auxiliary basis with rescaled primitives;
  * limits: w -> infinity gives the Coulomb integrals, and then wK == K exactly in exact arithmetic because
    sum_Q [J^-1 (Q|mi)] (Q|nj) = sum_Q [J^-1/2 (Q|mi)] [J^-1/2 (Q|nj)]."""
import ctypes as ct

import numpy as np
import pytest

from psi4_b200 import scf
from psi4_b200 import integrals as I
from psi4_b200.integrals import BasisSet, MintsHelper, Molecule


def water():
    return Molecule.from_zmat_h2o(0.96, 104.5)


def coulomb_with_rescaled_aux(mints, aux, omega):
    """(A|mn) over the auxiliary basis with every primitive c -> c' = c w^2/(c+w^2), coefficient x (c'/c)^(l+3/2)."""
    P = mints.primary
    e2 = aux.exps * omega * omega / (aux.exps + omega * omega)
    lprim = np.repeat(aux.l, aux.nprim)
    c2 = aux.coefs * (e2 / aux.exps) ** (lprim + 1.5)
    p = lambda a, t: a.ctypes.data_as(ct.POINTER(t))  # noqa: E731
    args = [len(aux.shells), p(aux.xyz, ct.c_double), p(aux.l, ct.c_int), p(aux.nprim, ct.c_int), p(aux.poff, ct.c_int),
            p(e2, ct.c_double), p(c2, ct.c_double)]
    out = np.zeros((aux.ncart, P.ncart, P.ncart))
    assert I._ints().ints_three_center(*args, *P._args(), p(out, ct.c_double)) == 0
    t = np.einsum("Aa,amn->Amn", aux.U, out, optimize=True)
    t = np.einsum("Mm,Amn->AMn", P.U, t, optimize=True)
    return np.einsum("Nn,AMn->AMN", P.U, t, optimize=True)


@pytest.mark.parametrize("omega", [0.2, 0.4, 1.5])
def test_erf_three_center_equals_coulomb_over_smeared_aux(omega):
    mol = water()
    P, A = BasisSet.build(mol, "cc-pvdz"), BasisSet.build(mol, "cc-pvdz-jkfit")  # s,p,d orbital / s,p,d,f fitting shells
    mints = MintsHelper(mol, P)
    got = mints.three_center(A, omega)
    ref = coulomb_with_rescaled_aux(mints, A, omega)
    assert np.abs(ref).max() > 0.1
    assert np.abs(got - ref).max() < 1e-12
    full = mints.three_center(A)
    assert np.abs(got - full).max() > 1e-3  # the attenuation is not a no-op at these omegas


def test_erf_limits():
    mol = water()
    P, A = BasisSet.build(mol, "cc-pvdz"), BasisSet.build(mol, "cc-pvdz-jkfit")
    mints = MintsHelper(mol, P)
    full = mints.three_center(A)
    assert np.abs(mints.three_center(A, 1.0e7) - full).max() < 1e-10
    small = mints.three_center(A, 1.0e-4)
    assert np.abs(small).max() < 2.0e-4 * np.abs(full).max() * 10


def run_wk(factory, omega):
    mol = water()
    P, A = BasisSet.build(mol, "cc-pvdz"), BasisSet.build(mol, "cc-pvdz-jkfit")
    jk = scf.build_jk(mol, P, A, jk_factory=factory, do_wK=True, omega=omega)
    jk.initialize()
    rhf_orbitals = np.linalg.qr(np.random.default_rng(3).standard_normal((P.nbf(), 5)))[0]
    jk.C_left_add(rhf_orbitals)
    jk.compute()
    return mol, P, A, jk, rhf_orbitals


def dense_wk(mol, P, A, C, omega, condition=1e-10):
    """wK_mn = sum_ij,AB (mi|A) [J^-1]_AB (B|erf|nj) straight from dense integrals."""
    mints = MintsHelper(mol, P)
    Amn, Wmn = mints.three_center(A), mints.three_center(A, omega)
    Jinv = scf.matrix_power(mints.metric(A), -1.0, condition)
    L = np.einsum("AB,Bmn,ni->Ami", Jinv, Amn, C, optimize=True)
    R = np.einsum("Amn,ni->Ami", Wmn, C, optimize=True)
    wK = np.einsum("Ami,Ani->mn", L, R, optimize=True)
    return 0.5 * (wK + wK.T)  # lr_symmetric: hermitivitized (MemDFJK.cc:104-110)


def oracle_wk_factory():
    from oracle_jk import OracleJK

    return lambda dfh, Ppq, m1, w: OracleJK(dfh, Ppq, m1, w)


def test_oracle_wk_from_real_integrals(oracle):
    mol, P, A, jk, C = run_wk(oracle_wk_factory(), 0.4)
    assert np.abs(jk.wK()[0] - dense_wk(mol, P, A, C, 0.4)).max() < 1e-10
    # w -> infinity: wK == K
    _, _, _, jk2, C2 = run_wk(oracle_wk_factory(), 1.0e7)
    assert np.abs(jk2.K()[0]).max() > 0.1
    assert np.abs(jk2.wK()[0] - jk2.K()[0]).max() < 1e-9


@pytest.mark.gpu
def test_gpu_wk_from_real_integrals(oracle):
    mol, P, A, jk, C = run_wk(None, 0.4)
    ref = dense_wk(mol, P, A, C, 0.4)
    assert np.abs(jk.wK()[0] - ref).max() < 1e-10
    _, _, _, jo, _ = run_wk(oracle_wk_factory(), 0.4)
    for got, want in ((jk.J()[0], jo.J()[0]), (jk.K()[0], jo.K()[0]), (jk.wK()[0], jo.wK()[0])):
        assert np.abs(got - want).max() < 1e-10
    jk.finalize()


@pytest.mark.gpu
def test_gpu_wk_with_on_device_fitting(oracle):
    """prepare_AO_wK_core (dfhelper.cc:589-699) driven through the engine's fitting entry points: Ppq_ with J^-1/2,
    m1Ppq_ with J^-1 (b200jk_set_metric + b200jk_fit_rows) and wPpq_ with no metric (b200jk_set_metric(NULL))."""
    mol = water()
    P, A = BasisSet.build(mol, "cc-pvdz"), BasisSet.build(mol, "cc-pvdz-jkfit")
    jk = scf.build_jk(mol, P, A, do_wK=True, omega=0.4, fit_on_device=True, fit_block=7)
    jk.initialize()
    C = np.linalg.qr(np.random.default_rng(3).standard_normal((P.nbf(), 5)))[0]
    jk.C_left_add(C)
    jk.compute()
    assert np.abs(jk.wK()[0] - dense_wk(mol, P, A, C, 0.4)).max() < 1e-10
    _, _, _, jo, _ = run_wk(oracle_wk_factory(), 0.4)
    for got, want in ((jk.J()[0], jo.J()[0]), (jk.K()[0], jo.K()[0]), (jk.wK()[0], jo.wK()[0])):
        assert np.abs(got - want).max() < 1e-10
    jk.finalize()


@pytest.mark.gpu
def test_gpu_knobs_set_between_build_and_initialize_take_effect(oracle):
    """The reference configures after build and before initialize (scf_iterator.py:112-135: build_JK, set_do_wK /
    set_omega / ..., initialize; MemDFJK::preiterations pushes the knobs into DFHelper first, MemDFJK.cc:71-96).
    build_JK(do_wK=True) without an omega must not fail, and the omega set afterwards is the one wK uses."""
    from psi4_b200 import JK

    mol = water()
    P, A = BasisSet.build(mol, "cc-pvdz"), BasisSet.build(mol, "cc-pvdz-jkfit")
    jk = JK.build_JK(P, A, do_wK=True)  # no omega yet
    jk.set_omega(0.4)
    jk.set_condition(1e-10)
    jk.initialize()
    C = np.linalg.qr(np.random.default_rng(3).standard_normal((P.nbf(), 5)))[0]
    jk.C_left_add(C)
    jk.compute()
    assert np.abs(jk.wK()[0] - dense_wk(mol, P, A, C, 0.4)).max() < 1e-10
    assert "Omega:                4.000E-01" in jk.print_header()
    jk.finalize()
    # wK tasked, omega never set: initialize() refuses (the reference would build erf integrals with omega = 0)
    from psi4_b200 import PsiException

    jk2 = JK.build_JK(P, A, do_wK=True)
    with pytest.raises(PsiException):
        jk2.initialize()
