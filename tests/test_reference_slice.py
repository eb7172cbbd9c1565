"""The restatement (oracle/dfjk_oracle.c) against the reference's OWN object code.

oracle/ref_build.py slices the member functions of the MEM_DF J/K path out of /root/reference/.../dfhelper.cc at build
time (build_JK, compute_JK, compute_J_symm, compute_J, compute_K, compute_wK, first_transform_pQq,
Qshell_blocks_for_JK_build, contract_metric_AO_core_symm, ...) and compiles them unmodified against a stand-in for
the little of psi4's runtime they touch (oracle/ref_shim.h).  With one thread and the whole auxiliary index in one block
the two must agree BIT FOR BIT -- same loops, same BLAS calls in the same order; that is what "restatement" claims.

Two places where only a tolerance is meaningful, both properties of the reference itself:
  * several OpenMP threads: compute_J_symm / compute_J accumulate d_Q per thread rank under schedule(guided) and reduce
    in rank order (dfhelper.cc:3193-3199), so J's rounding depends on the run; K has no such reduction and stays exact;
  * Q blocking: Qshell_blocks_for_JK_build resets total_AO_buffer to 0 after its first block even in the in-core case
    (dfhelper.cc:866), so every later block is sized as if the tensor cost nothing -- the reference's block list is
    [first shell group], [everything else], not uniform slices; sums over Q are merely re-associated.
Skipped when neither /root/reference nor a prebuilt oracle/_ref/libref_dfjk.so is present."""
import numpy as np
import pytest

from psi4_b200 import DFHelper


@pytest.fixture(scope="module")
def ref(oracle):
    if oracle.ref_lib() is None:
        pytest.skip("oracle/_ref/libref_dfjk.so not available")
    return oracle


def system(rng, n, a, density):
    r = rng.random((n, n))
    keep = (r + r.T) * 0.5 < density
    np.fill_diagonal(keep, True)
    d = DFHelper(n, a)
    d.prepare_sparsity(keep=keep)
    t = lambda: (lambda b: b + b.transpose(0, 2, 1))(rng.standard_normal((a, n, n)) * 0.2)  # noqa: E731
    return keep, d, d.pack(t()), d.pack(t()), d.pack(t())


def flat(res):
    return [m for group in res[:3] if group is not None for m in group]


@pytest.mark.parametrize("lr", [True, False])
@pytest.mark.parametrize("density", [1.0, 0.6, 0.15])
def test_restatement_is_bit_identical_to_reference_object_code(ref, lr, density):
    rng = np.random.default_rng(int(density * 100) + lr)
    n, a = 53, 41
    keep, d, P, P1, PW = system(rng, n, a, density)
    sp = ref.Sparsity(keep.astype(np.uint8), a)
    Cl = [rng.standard_normal((n, o)) for o in (7, 0, 1, 30)]
    Cr = None if lr else [rng.standard_normal(c.shape) for c in Cl]
    kw = dict(do_wK=True, m1Ppq=P1, wPpq=PW, nthreads=1)
    port = ref.build_JK(sp, P, Cl, Cr, **kw)
    real = ref.build_JK(sp, P, Cl, Cr, impl="ref", **kw)
    assert len(flat(port)) == 12
    for x, y in zip(flat(port), flat(real)):
        assert np.array_equal(x, y)
    assert not real[1][1].any()  # nocc == 0: K untouched (dfhelper.cc:3354-3357)
    # task toggles reach the same branches
    for dj, dk in ((True, False), (False, True)):
        p2 = ref.build_JK(sp, P, Cl, Cr, do_J=dj, do_K=dk, nthreads=1)
        r2 = ref.build_JK(sp, P, Cl, Cr, do_J=dj, do_K=dk, nthreads=1, impl="ref")
        assert (p2[0] is None) == (not dj) and (p2[1] is None) == (not dk)
        for x, y in zip(flat(p2), flat(r2)):
            assert np.array_equal(x, y)


@pytest.mark.parametrize("lr", [True, False])
def test_threads_and_q_blocking_only_reassociate(ref, lr):
    rng = np.random.default_rng(9 + lr)
    n, a = 47, 38
    keep, d, P, _, _ = system(rng, n, a, 0.5)
    sp = ref.Sparsity(keep.astype(np.uint8), a)
    Cl = [rng.standard_normal((n, 11)), rng.standard_normal((n, 4))]
    Cr = None if lr else [rng.standard_normal(c.shape) for c in Cl]
    base = ref.build_JK(sp, P, Cl, Cr, nthreads=1, impl="ref")
    scale = max(np.abs(m).max() for m in flat(base))
    threaded = ref.build_JK(sp, P, Cl, Cr, nthreads=4, impl="ref")
    for x, y in zip(base[1], threaded[1]):
        assert np.array_equal(x, y)  # K: no cross-thread reduction
    for x, y in zip(base[0], threaded[0]):
        assert np.abs(x - y).max() < 1e-12 * scale
    for qb in (1, 9, 37):
        blocked = ref.build_JK(sp, P, Cl, Cr, nthreads=1, q_block=qb, impl="ref")
        port = ref.build_JK(sp, P, Cl, Cr, nthreads=1, q_block=qb)
        for x, y, z in zip(flat(base), flat(blocked), flat(port)):
            assert np.abs(x - y).max() < 1e-12 * scale and np.abs(x - z).max() < 1e-12 * scale


@pytest.mark.parametrize("density", [1.0, 0.4])
def test_fitting_restatement_is_bit_identical(ref, density):
    rng = np.random.default_rng(3)
    n, a = 31, 26
    keep, d, _, _, _ = system(rng, n, a, density)
    sp = ref.Sparsity(keep.astype(np.uint8), a)
    U = rng.standard_normal((a, n, n))
    U = U + U.transpose(0, 2, 1)
    g = rng.standard_normal((a, a))
    met = g @ g.T / a + np.eye(a)
    whole = ref.contract_metric_AO_core_symm(sp, d.pack_symm(U), met, nthreads=2, impl="ref")
    assert np.array_equal(whole, ref.contract_metric_AO_core_symm(sp, d.pack_symm(U), met, nthreads=2))
    blocks = np.zeros_like(whole)
    for m0 in range(0, n, 8):  # the p-blocked loop of prepare_AO_core (:566-585)
        m1 = min(n, m0 + 8)
        ref.contract_metric_AO_core_symm(sp, d.pack_symm(U, m0, m1), met, blocks, begin=m0, end=m1 - 1, impl="ref")
    assert np.array_equal(whole, blocks)
    assert np.abs(whole - d.pack(np.einsum("QR,Rmn->Qmn", met, U))).max() < 1e-11


def test_reference_object_code_reproduces_the_published_energies(ref):
    """End to end with no restatement in the loop: integrals from the host front end, tensors packed as DFHelper packs
    them, J/K from the reference's own compiled build_JK -> tests/tu1-h2o-energy (-76.0266327341067125) and the scf5
    triplet UHF / ROHF energies (two and docc+socc densities per call) to 1e-8 Eh."""
    import json
    import os

    from oracle_jk import OracleJK
    from psi4_b200 import scf
    from psi4_b200.integrals import BasisSet, Molecule

    anch = json.load(open(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "reference_anchors.json")))
    factory = lambda dfh, Ppq: OracleJK(dfh, Ppq, impl="ref")  # noqa: E731
    a = anch["tu1_h2o_ccpvdz"]
    mol = Molecule.from_zmat_h2o(a["zmat"]["r_oh_angstrom"], a["zmat"]["angle_deg"])
    P, A = BasisSet.build(mol, a["basis"]), BasisSet.build(mol, a["aux"])
    jk = scf.build_jk(mol, P, A, jk_factory=factory)
    jk.initialize()
    assert abs(scf.RHF(mol, P, jk).compute_energy() - a["scf_total_energy"]) < 1e-8
    a = anch["scf5_o2_ccpvtz"]
    mol = Molecule.from_angstrom(["O", "O"], [[0, 0, 0], [0, 0, a["r_oo_angstrom"]]])
    P, A = BasisSet.build(mol, a["basis"]), BasisSet.build(mol, a["aux"])
    jk = scf.build_jk(mol, P, A, jk_factory=factory)
    jk.initialize()
    assert abs(scf.UHF(mol, P, jk, multiplicity=3).compute_energy() - a["triplet_uhf_df"]) < 1e-8
    assert abs(scf.ROHF(mol, P, jk, multiplicity=3).compute_energy() - a["triplet_rohf_df"]) < 1e-8


def test_sparsity_tables_match_reference_object_code(ref):
    """prepare_sparsity's table half (dfhelper.cc:370-420), compiled from the reference, on real Schwarz maxima (five Ar
    atoms / cc-pVDZ, the test_jkmemory system: 75.8 % of the pairs screened) and on random ones: mask rule, every index
    table and the shell mask equal the host mirror (psi4_b200.DFHelper, what b200jk_set_layout is fed) and the
    restatement's tables element for element."""
    from psi4_b200.integrals import BasisSet, MintsHelper, Molecule

    mol = Molecule.from_angstrom(["Ar"] * 5, [[0, 0, z] for z in (0.0, 5.0, 15.0, 25.0, 35.0)])
    P = BasisSet.build(mol, "cc-pvdz")
    cases = [(MintsHelper(mol, P).schwarz_function_maxima(), [P.shell_nfunction(s) for s in range(P.nshell())], 560, 1e-12)]
    rng = np.random.default_rng(0)
    f = 10.0 ** rng.uniform(-30, 0, (40, 40))
    f = np.maximum(f, f.T)
    np.fill_diagonal(f, 1.0)
    cases.append((f, [1, 3, 5, 1, 3, 6, 10, 1, 3, 7], 33, 1e-9))
    for fmax, psh, naux, cutoff in cases:
        n = fmax.shape[0]
        d = DFHelper(n, naux)
        d.set_schwarz_cutoff(cutoff)
        d.prepare_blocking(psh, [naux])
        d.prepare_sparsity(fmax)
        t = ref.ref_sparsity_tables(fmax, naux, cutoff, d.pshell_aggs_)
        assert 0.0 < d.ao_sparsity() < 1.0
        assert np.array_equal(t["schwarz_fun_index_"].reshape(n, n), d.schwarz_fun_index_)
        for k in ("small_skips_", "big_skips_", "symm_small_skips_", "symm_ignored_columns_", "symm_big_skips_"):
            assert np.array_equal(t[k], getattr(d, k)), k
        assert np.array_equal(t["schwarz_shell_mask_"].reshape(len(psh), len(psh)).astype(bool), d.schwarz_shell_mask_)
        sp = ref.Sparsity(d.keep_.astype(np.uint8), naux)
        assert np.array_equal(sp.fun_index.reshape(n, n), d.schwarz_fun_index_) and np.array_equal(sp.big_skips, t["big_skips_"])


def test_metric_power_matches_reference_object_code(ref):
    """Matrix::power (libmints/matrix.cc:2370-2424) compiled from the reference vs the host driver's matrix_power
    (scf.py): same eigenvalue cut (|lambda| < cutoff * |lambda|max dropped only for negative powers), same result on the
    real fitting metric (A|B) of water / cc-pVDZ-jkfit and on a spectrum with one direction below the cut."""
    from psi4_b200 import scf
    from psi4_b200.integrals import BasisSet, MintsHelper, Molecule

    mol = Molecule.from_zmat_h2o(0.96, 104.5)
    P, A = BasisSet.build(mol, "cc-pvdz"), BasisSet.build(mol, "cc-pvdz-jkfit")
    metric = MintsHelper(mol, P).metric(A)
    for alpha in (-0.5, -1.0):  # mpower_ and wmpower_, dfhelper.h:356-357
        R, kept = ref.ref_matrix_power(metric, alpha, 1e-10)
        assert kept == A.nbf()
        M = scf.matrix_power(metric, alpha, 1e-10)
        assert np.abs(R - M).max() < 1e-9 * np.abs(R).max()
    rng = np.random.default_rng(0)
    V = np.linalg.qr(rng.standard_normal((24, 24)))[0]
    w = 10.0 ** rng.uniform(-3, 0, 24)
    w[5] = 1e-13 * w.max()
    S = (V * w) @ V.T
    S = 0.5 * (S + S.T)
    R, kept = ref.ref_matrix_power(S, -0.5, 1e-10)
    assert kept == 23  # the 1e-13 direction is dropped ...
    assert np.abs(R - scf.matrix_power(S, -0.5, 1e-10)).max() < 1e-9 * np.abs(R).max()
    R, kept = ref.ref_matrix_power(S, 0.5, 1e-10)
    assert kept == 24  # ... but never for a positive power (matrix.cc:2403)
    assert np.abs(R - scf.matrix_power(S, 0.5, 1e-10)).max() < 1e-9 * np.abs(R).max()


@pytest.mark.parametrize("puream", [True, False], ids=["spherical", "cartesian"])
def test_dfjk_compare_with_real_integrals(ref, puream):
    """tests/pytests/test_dfjk.py as it is written -- water (R = 1.00, 103.1 deg), cc-pVDZ / cc-pVDZ-jkfit, spherical
    and cartesian, five random spaces, seven (C_left, C_right) pairs, C_right always added -- with REAL integrals from
    the host front end.  The reference's counterpart to MemDFJK there is DiskDFJK, a different algorithm for the same
    definition; here that role is played by the dense definition J/K = f(B), B = J^-1/2 (A|mn), and the MemDFJK side by
    the reference's own object code and by the restatement.  Gate: the reference's 9 decimals."""
    from oracle_jk import OracleJK
    from psi4_b200.integrals import BasisSet, Molecule
    from psi4_b200.jk import JK

    mol = Molecule.from_zmat_h2o(1.00, 103.1)
    primary = BasisSet.build(mol, "cc-pvdz", puream=puream)
    aux = BasisSet.build(mol, "cc-pvdz-jkfit", puream=puream)
    assert (primary.nbf(), aux.nbf()) == ((24, 116) if puream else (25, 131))
    rng = np.random.default_rng(7)
    sizes = [16, 16, 20, 20, 30]
    spaces = [rng.random((primary.nbf(), s)) for s in sizes]
    pairs = [[0, 0], [0, 1], [1, 1], [2, 2], [3, 2], [3, 3], [4, 4]]
    results = {}
    for impl in ("ref", "port"):
        jk = JK.build_JK(primary, aux, jk_factory=lambda dfh, Ppq, impl=impl: OracleJK(dfh, Ppq, impl=impl))
        jk.initialize()
        for left, right in pairs:
            jk.C_left_add(spaces[left])
            jk.C_right_add(spaces[right])
        jk.compute()
        results[impl] = (jk.J(), jk.K())
        dfh, Ppq = jk.dfh_, jk.Ppq
    B = dfh.unpack(Ppq)  # (naux, nbf, nbf): nothing is screened in water, the packed tensor is the dense definition
    assert dfh.ao_sparsity() == 0.0
    for i, (left, right) in enumerate(pairs):
        D = spaces[left] @ spaces[right].T
        Jd = np.einsum("Qmn,Q->mn", B, np.einsum("Qls,ls->Q", B, D))
        Kd = np.einsum("Qmi,Qni->mn", np.einsum("Qmn,ni->Qmi", B, spaces[left]), np.einsum("Qmn,ni->Qmi", B, spaces[right]))
        for impl in ("ref", "port"):
            assert np.abs(results[impl][0][i] - Jd).max() < 1e-9, f"J{i} {impl}"
            assert np.abs(results[impl][1][i] - Kd).max() < 1e-9, f"K{i} {impl}"
