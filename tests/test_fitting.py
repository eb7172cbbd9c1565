"""On-device fitting: B = metric . (A|mn) + mirror copy, DFHelper::contract_metric_AO_core_symm
(dfhelper.cc:1653-1678).  CPU: the oracle restatement vs a dense einsum.  GPU: b200jk_set_metric / b200jk_fit_rows vs
the oracle, in several p-blocks, with screening, with and without a metric; then the reference's tu1 energy with the
tensor fitted on the device."""
import json
import os

import numpy as np
import pytest

from psi4_b200 import DFHelper

ANCH = json.load(open(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "reference_anchors.json")))


def case(rng, n, a, density):
    r = rng.random((n, n))
    keep = (r + r.T) * 0.5 < density
    np.fill_diagonal(keep, True)
    d = DFHelper(n, a)
    d.prepare_sparsity(keep=keep)
    U = rng.standard_normal((a, n, n))
    U = U + U.transpose(0, 2, 1)  # unfitted (A|mn), symmetric in mn
    g = rng.standard_normal((a, a))
    met = g @ g.T / a + np.eye(a)  # any symmetric matrix works for the contraction
    return keep, d, U, met


@pytest.mark.parametrize("density", [1.0, 0.5])
def test_oracle_fitting_matches_dense(oracle, density):
    rng = np.random.default_rng(4)
    n, a = 23, 17
    keep, d, U, met = case(rng, n, a, density)
    sp = oracle.Sparsity(keep.astype(np.uint8), a)
    assert np.array_equal(d.symm_big_skips_, sp.symm_big_skips)
    P = np.zeros(sp.packed_size)
    for m0 in range(0, n, 7):  # blocks, as prepare_AO_core feeds them
        m1 = min(n, m0 + 7)
        oracle.contract_metric_AO_core_symm(sp, d.pack_symm(U, m0, m1), met, P, begin=m0, end=m1 - 1)
    ref = d.pack(np.einsum("QR,Rmn->Qmn", met, U))
    assert np.abs(P - ref).max() < 1e-11


@pytest.mark.gpu
@pytest.mark.parametrize("density,block", [(1.0, 1000), (0.6, 13), (0.15, 40)])
def test_gpu_fit_rows_matches_oracle(oracle, density, block):
    from psi4_b200 import Engine

    rng = np.random.default_rng(int(density * 10))
    n, a = 150, 203
    keep, d, U, met = case(rng, n, a, density)
    sp = oracle.Sparsity(keep.astype(np.uint8), a)
    ref = oracle.contract_metric_AO_core_symm(sp, d.pack_symm(U), met)
    e = Engine(1)
    e.set_layout(n, a, d.small_skips_, d.big_skips_, d.schwarz_fun_index_)
    e.set_metric(met)
    for m0 in range(0, n, block):
        m1 = min(n, m0 + block)
        e.fit_rows(0, m0, m1, d.pack_symm(U, m0, m1))
    scale = np.abs(ref).max()
    for m in range(n):
        got = e.download_rows(0, m, 0, a).ravel()
        want = ref[int(d.big_skips_[m]):int(d.big_skips_[m + 1])]
        assert np.abs(got - want).max() < 1e-12 * max(1.0, scale), f"row-block {m}"
    st = e.fit_stats()
    assert st["flops"] > 0 and st["ms_gemm"] > 0
    # the fitted tensor drives a JK build like an uploaded one
    C = rng.standard_normal((n, 9))
    J, K, _ = e.compute([C], None, [C @ C.T])
    Jo, Ko, _, _ = oracle.build_JK(sp, ref, [C])
    assert np.abs(J[0] - Jo[0]).max() < 1e-10 * max(1.0, np.abs(Jo[0]).max())
    assert np.abs(K[0] - Ko[0]).max() < 1e-10 * max(1.0, np.abs(Ko[0]).max())
    # no metric: plain scatter + mirror (the wPpq_ tensor)
    e.set_metric(None)
    e.fit_rows(2, 0, n, d.pack_symm(U))
    plain = d.pack(U * keep[None])
    for m in (0, n // 2, n - 1):
        assert np.array_equal(e.download_rows(2, m, 0, a).ravel(), plain[int(d.big_skips_[m]):int(d.big_skips_[m + 1])])
    e.close()


@pytest.mark.gpu
def test_gpu_tu1_energy_with_device_fitting():
    from psi4_b200 import scf
    from psi4_b200.integrals import BasisSet, Molecule

    a = ANCH["tu1_h2o_ccpvdz"]
    mol = Molecule.from_zmat_h2o(a["zmat"]["r_oh_angstrom"], a["zmat"]["angle_deg"])
    P, A = BasisSet.build(mol, a["basis"]), BasisSet.build(mol, a["aux"])
    jk = scf.build_jk(mol, P, A, fit_on_device=True, fit_block=5)
    jk.initialize()
    E = scf.RHF(mol, P, jk).compute_energy()
    assert abs(E - a["scf_total_energy"]) < 1e-8


@pytest.mark.gpu
@pytest.mark.parametrize("registered", [False, True])
def test_gpu_producers_pipeline_with_reused_block_buffer(oracle, monkeypatch, registered):
    """psi4 fills ONE block buffer over and over (Mp / Ppq rows, dfhelper.cc:553-585).  b200jk_fit_rows and
    b200jk_upload_rows return before the GPU is done with a block, so the buffer may be overwritten right away: a
    pageable buffer is copied into the page-locked ring first, a registered one is held until its DMA has finished.
    A tiny group size drives many groups through both ring slots and both device buffers."""
    from psi4_b200 import Engine

    monkeypatch.setenv("B200JK_STAGE_BYTES", str(48 * 1024))
    rng = np.random.default_rng(77)
    n, a, block = 120, 97, 17
    keep, d, U, met = case(rng, n, a, 0.7)
    sp = oracle.Sparsity(keep.astype(np.uint8), a)
    ref = oracle.contract_metric_AO_core_symm(sp, d.pack_symm(U), met)
    e = Engine(1)
    e.set_layout(n, a, d.small_skips_, d.big_skips_, d.schwarz_fun_index_)
    e.set_metric(met)
    biggest = max(int(d.symm_big_skips_[min(n, m0 + block)] - d.symm_big_skips_[m0]) for m0 in range(0, n, block))
    buf = np.zeros(biggest)
    if registered:
        e.register_host(buf)
    for m0 in range(0, n, block):
        m1 = min(n, m0 + block)
        blk = d.pack_symm(U, m0, m1)
        buf[:blk.size] = blk
        e.fit_rows(0, m0, m1, buf)
        buf[:] = np.nan  # the caller reuses its buffer immediately
    for m in range(n):
        got = e.download_rows(0, m, 0, a).ravel()
        want = ref[int(d.big_skips_[m]):int(d.big_skips_[m + 1])]
        assert np.abs(got - want).max() < 1e-12 * max(1.0, np.abs(ref).max()), f"row-block {m}"
    # the same for the fitted-rows upload (streaming b200jk_upload_rows into tensor slot 1)
    biggest = max(int(d.big_skips_[min(n, m0 + block)] - d.big_skips_[m0]) for m0 in range(0, n, block))
    buf2 = np.zeros(biggest)
    if registered:
        e.register_host(buf2)
    for m0 in range(0, n, block):
        m1 = min(n, m0 + block)
        rows = ref[int(d.big_skips_[m0]):int(d.big_skips_[m1])]
        buf2[:rows.size] = rows
        e.upload_rows(1, m0, m1, buf2[:rows.size])
        buf2[:] = np.nan
    for m in range(n):
        assert np.array_equal(e.download_rows(1, m, 0, a).ravel(), ref[int(d.big_skips_[m]):int(d.big_skips_[m + 1])])
    if registered:
        e.unregister_host(buf)
        e.unregister_host(buf2)
    e.close()


@pytest.mark.gpu
def test_gpu_matrix_power_follows_the_reference_drop_rule():
    """b200jk_matrix_power = Matrix::power (libmints/matrix.cc:2370-2424) on the device: same result as the host
    restatement (itself checked against the reference's compiled Matrix::power in test_reference_slice.py) on the real
    fitting metric of water, the same eigenvalue cut on a spectrum with one direction below it, no cut for positive
    powers, and the count of eigenvalues kept."""
    from psi4_b200 import Engine, scf
    from psi4_b200.integrals import BasisSet, MintsHelper, Molecule

    e = Engine(1)
    mol = Molecule.from_zmat_h2o(0.96, 104.5)
    P, A = BasisSet.build(mol, "cc-pvdz"), BasisSet.build(mol, "cc-pvdz-jkfit")
    metric = MintsHelper(mol, P).metric(A)
    for alpha in (-0.5, -1.0):  # mpower_ and wmpower_, dfhelper.h:356-357
        R, kept, ms = e.matrix_power(metric, alpha, 1e-10, with_info=True)
        assert kept == A.nbf() and ms > 0
        M = scf.matrix_power(metric, alpha, 1e-10)
        assert np.abs(R - M).max() < 1e-9 * np.abs(M).max()
        assert np.abs(R - R.T).max() < 1e-12 * np.abs(M).max()
    rng = np.random.default_rng(0)
    n = 300
    V = np.linalg.qr(rng.standard_normal((n, n)))[0]
    w = 10.0 ** rng.uniform(-3, 0, n)
    w[5] = 1e-13 * w.max()
    S = (V * w) @ V.T
    S = 0.5 * (S + S.T)
    R, kept, _ = e.matrix_power(S, -0.5, 1e-10, with_info=True)
    assert kept == n - 1  # the 1e-13 direction is dropped ...
    assert np.abs(R - scf.matrix_power(S, -0.5, 1e-10)).max() < 1e-9 * np.abs(R).max()
    R, kept, _ = e.matrix_power(S, 0.5, 1e-10, with_info=True)
    assert kept == n  # ... but never for a positive power (matrix.cc:2403)
    assert np.abs(R - scf.matrix_power(S, 0.5, 1e-10)).max() < 1e-9 * np.abs(R).max()
    e.close()
