"""CPU tests (-m "not gpu"): the oracle restatement vs the independent dense-einsum oracle, the
host-side table logic vs the oracle's restatement of prepare_sparsity, and the C-ABI surface."""
import ctypes as ct
import os
import re

import numpy as np
import pytest

from psi4_b200 import DFHelper, lib as b2lib

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def random_mask(rng, n, density):
    r = rng.random((n, n))
    keep = (r + r.T) * 0.5 < density
    np.fill_diagonal(keep, True)
    return keep


def sym_tensor(rng, a, n):
    b = rng.standard_normal((a, n, n))
    return b + b.transpose(0, 2, 1)


@pytest.mark.parametrize("density", [1.0, 0.6, 0.15])
def test_tables_match_reference_restatement(oracle, density):
    rng = np.random.default_rng(3)
    n, a = 37, 11
    keep = random_mask(rng, n, density)
    sp = oracle.Sparsity(keep, a)
    d = DFHelper(n, a)
    d.prepare_sparsity(keep=keep)
    assert np.array_equal(d.schwarz_fun_index_.ravel(), sp.fun_index)
    assert np.array_equal(d.small_skips_, sp.small_skips)
    assert np.array_equal(d.big_skips_, sp.big_skips)
    assert np.array_equal(d.symm_small_skips_, sp.symm_small_skips)
    assert np.array_equal(d.symm_ignored_columns_, sp.symm_ignored_columns)
    assert np.array_equal(d.symm_big_skips_, sp.symm_big_skips)
    B = sym_tensor(rng, a, n)
    assert np.array_equal(d.pack(B), oracle.pack_pQq(sp, B))
    assert np.array_equal(d.unpack(d.pack(B)), B * keep[None])


def test_schwarz_mask_rule(oracle):
    # dfhelper.cc:371-386: keep iff value >= cutoff^2 / max
    rng = np.random.default_rng(5)
    n = 20
    v = 10.0 ** rng.uniform(-30, 0, (n, n))
    v = np.maximum(v, v.T)
    np.fill_diagonal(v, 1.0)
    keep_o = oracle.schwarz_mask(v, 1e-12).astype(bool)
    d = DFHelper(n, 3)
    d.set_schwarz_cutoff(1e-12)
    d.prepare_sparsity(fun_max_vals=v)
    assert np.array_equal(d.keep_, keep_o)
    assert 0 < keep_o.sum() < n * n
    d0 = DFHelper(n, 3)
    d0.set_schwarz_cutoff(0.0)  # SCREENING=NONE, jk.cc:60-61
    d0.prepare_sparsity(fun_max_vals=v)
    assert d0.keep_.all() and d0.ao_sparsity() == 0.0


@pytest.mark.parametrize("lr", [True, False])
@pytest.mark.parametrize("density", [1.0, 0.4])
@pytest.mark.parametrize("q_block", [0, 7])
def test_oracle_vs_dense_einsum(oracle, lr, density, q_block):
    rng = np.random.default_rng(11)
    n, a = 24, 29
    keep = random_mask(rng, n, density)
    sp = oracle.Sparsity(keep, a)
    B = sym_tensor(rng, a, n)
    P = oracle.pack_pQq(sp, B)
    Cl = [rng.random((n, 5)), rng.random((n, 9)), rng.random((n, 0))]
    Cr = None if lr else [rng.random((n, 5)), rng.random((n, 9)), rng.random((n, 0))]
    J, K, _, _ = oracle.build_JK(sp, P, Cl, Cr, q_block=q_block, nthreads=3)
    Jd, Kd, _ = oracle.dense_JK(B, keep, Cl, Cr)
    for i in range(3):
        assert np.abs(J[i] - Jd[i]).max() < 1e-10
        assert np.abs(K[i] - Kd[i]).max() < 1e-10
    assert not K[2].any()  # nocc == 0 => K untouched (dfhelper.cc:3354-3357)


@pytest.mark.parametrize("lr", [True, False])
def test_oracle_wK(oracle, lr):
    rng = np.random.default_rng(13)
    n, a = 18, 15
    keep = random_mask(rng, n, 0.7)
    sp = oracle.Sparsity(keep, a)
    B, M1, W = (sym_tensor(rng, a, n) for _ in range(3))
    P, P1, PW = (oracle.pack_pQq(sp, x) for x in (B, M1, W))
    Cl = [rng.random((n, 4))]
    Cr = None if lr else [rng.random((n, 4))]
    J, K, wK, _ = oracle.build_JK(sp, P, Cl, Cr, do_wK=True, m1Ppq=P1, wPpq=PW)
    Jd, Kd, wKd = oracle.dense_JK(B, keep, Cl, Cr, dense_m1=M1, dense_w=W)
    assert np.abs(wK[0] - wKd[0]).max() < 1e-10
    assert np.abs(K[0] - Kd[0]).max() < 1e-10
    if lr:
        assert np.array_equal(wK[0], wK[0].T)  # hermitivitize, MemDFJK.cc:104-110


def test_synth_generator_is_symmetric_and_sliceable(oracle):
    n, a = 16, 12
    rng = np.random.default_rng(1)
    keep = random_mask(rng, n, 0.5)
    amp = rng.random((n, n))
    amp = amp + amp.T
    full = oracle.synth_fill(oracle.Sparsity(keep, a), 0, a, 99, amp)
    d = DFHelper(n, a)
    d.prepare_sparsity(keep=keep)
    dense = d.unpack(full)
    assert np.array_equal(dense, dense.transpose(0, 2, 1))
    assert np.abs(dense).max() <= np.abs(amp).max() and np.abs(dense).max() > 0
    part = oracle.synth_fill(oracle.Sparsity(keep, 5), 4, 5, 99, amp)
    d5 = DFHelper(n, 5)
    d5.prepare_sparsity(keep=keep)
    assert np.array_equal(d5.unpack(part), dense[4:9])


def test_abi_exports_every_declared_symbol():
    """The C-ABI library loads on a CPU-only box and exports everything include/b200jk.h declares."""
    hdr = open(os.path.join(ROOT, "include", "b200jk.h")).read()
    declared = set(re.findall(r"\b(b200jk_[a-z0-9_]+)\s*\(", hdr))
    declared.discard("b200jk_t")
    assert len(declared) >= 18
    L = b2lib.load()
    for name in sorted(declared):
        assert hasattr(L, name), f"{name} declared in b200jk.h but not exported"
        assert name in b2lib.SIGNATURES, f"{name} has no ctypes signature in psi4_b200/lib.py"


def test_engine_refuses_without_gpu():
    import torch

    if torch.cuda.is_available():
        pytest.skip("GPU present")
    with pytest.raises(b2lib.B200JKError) as e:
        b2lib.Engine(1)
    assert e.value.code == 5  # B200JK_ERR_NODEVICE: no CPU fallback


@pytest.mark.parametrize("basis", ["cc-pvdz", "cc-pv5z"])
def test_memory_estimate_matches_reference_integers(basis):
    """tests/pytests/test_jkmemory.py:44,49 -- jk.memory_estimate() of MEM_DF for five Ar atoms on a line
    (1 590 520 and 57 020 770 doubles): pins prepare_sparsity's mask (Schwarz integrals up to h functions, cutoff
    1e-12, dfhelper.cc:371-386), big_skips_ (:390-397), naux_ / Qshell_max_ and DFHelper::get_core_size (:216-236)
    as exact integers."""
    import json
    import os

    from psi4_b200 import MemDFJK
    from psi4_b200.dfhelper import DFHelper
    from psi4_b200.integrals import BasisSet, MintsHelper, Molecule, basis_shape

    a = json.load(open(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "reference_anchors.json")))["jkmemory_ar5"]
    mol = Molecule.from_angstrom([g[0] for g in a["geometry_angstrom"]], [g[1:] for g in a["geometry_angstrom"]])
    P = BasisSet.build(mol, basis)
    naux, qshell_max = basis_shape(mol, basis + "-jkfit", puream=True)
    dfh = DFHelper(P.nbf(), naux)
    dfh.set_Qshell_max(qshell_max)
    dfh.prepare_sparsity(MintsHelper(mol, P).schwarz_function_maxima())
    jk = MemDFJK(dfh)           # no tensor, no device: memory_estimate is host logic (MemDFJK.cc:65-69)
    jk.set_omp_nthread(1)       # "ref valid for -n1 only", test_jkmemory.py:35
    jk.set_do_wK(False)
    assert jk.name() == "MemDFJK"
    assert jk.memory_estimate() == a["mem_df_estimate_doubles"][basis]


def test_build_jk_df_autoselect_follows_reference_threshold():
    """jk.cc:206-229: SCF_TYPE=DF keeps MemDFJK iff memory_estimate() < doubles (strict).  Five Ar atoms / cc-pVDZ need
    1 590 520 doubles (test_jkmemory.py:44): one double more is enough, exactly that many is not."""
    import sys

    sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
    from oracle_jk import OracleJK
    from psi4_b200.integrals import BasisSet, Molecule
    from psi4_b200.jk import JK, PsiException

    mol = Molecule.from_angstrom(["Ar"] * 5, [[0, 0, z] for z in (0.0, 5.0, 15.0, 25.0, 35.0)])
    P, A = BasisSet.build(mol, "cc-pvdz"), BasisSet.build(mol, "cc-pvdz-jkfit")
    with pytest.raises(PsiException, match="DiskDFJK"):
        JK.build_JK(P, A, doubles=1590520, scf_type="DF")
    jk = JK.build_JK(P, A, doubles=1590521, scf_type="DF", jk_factory=lambda dfh, Ppq: OracleJK(dfh, Ppq))
    assert jk.basisset() is P and jk.memory_ == 1590521
    with pytest.raises(PsiException):
        JK.build_JK(P, A, scf_type="PK")


def test_pshell_blocking_tiles_the_primary_shells():
    """pshell_blocks_for_AO_build (dfhelper.cc:700-762): blocks are consecutive, cover every primary shell once, each
    respects the memory constraint it was cut by, and a budget below one shell's need throws like the reference."""
    rng = np.random.default_rng(5)
    psh = [1, 1, 3, 3, 5, 1, 3, 5, 7, 1, 3]
    n, a = sum(psh), 40
    r = rng.random((n, n))
    keep = (r + r.T) < 1.2
    np.fill_diagonal(keep, True)
    d = DFHelper(n, a)
    d.prepare_blocking(psh, [1, 3, 5, 7, 9, 15])
    assert d.Qshell_max_ == 15 and int(d.pshell_aggs_[-1]) == n
    d.prepare_sparsity(keep=keep)
    sh = np.repeat(np.arange(len(psh)), psh)
    want = np.zeros((len(psh), len(psh)), bool)
    for i in range(n):
        for j in range(n):
            want[sh[i], sh[j]] |= keep[i, j]
    assert np.array_equal(d.schwarz_shell_mask_, want)
    full = int(d.big_skips_[n])
    whole = int(d.symm_big_skips_[n])
    steps, largest, block = d.pshell_blocks_for_AO_build(full + 2 * whole)
    assert steps == [(0, len(psh) - 1)] and largest == whole and block == n
    mem = full + 2 * whole // 3
    steps, largest, block = d.pshell_blocks_for_AO_build(mem)
    assert len(steps) > 1 and steps[0][0] == 0 and steps[-1][1] == len(psh) - 1
    for (a0, a1), (b0, b1) in zip(steps, steps[1:]):
        assert b0 == a1 + 1
    for s0, s1 in steps:
        cost = int(d.symm_big_skips_[int(d.pshell_aggs_[s1 + 1])] - d.symm_big_skips_[int(d.pshell_aggs_[s0])])
        assert 2 * cost + full <= mem and cost <= largest
    with pytest.raises(MemoryError):
        d.pshell_blocks_for_AO_build(full + 10)
