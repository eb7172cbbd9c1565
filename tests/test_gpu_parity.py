"""GPU parity tests (-m gpu): the CUDA engine, through the C ABI (ctypes) and the MemDFJK mirror,
against the CPU oracle (restatement of dfhelper.cc) and the dense-einsum oracle on identical
seeded inputs.  Gate: max-abs 1e-10 on J/K/wK elements (BASELINE.json north_star)."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu

TOL = 1e-10


def random_mask(rng, n, density):
    r = rng.random((n, n))
    keep = (r + r.T) * 0.5 < density
    np.fill_diagonal(keep, True)
    return keep


def banded_mask(n, half):
    i = np.arange(n)
    return np.abs(i[:, None] - i[None, :]) <= half


def sym_tensor(rng, a, n, scale=1.0):
    b = rng.standard_normal((a, n, n)) * scale
    return b + b.transpose(0, 2, 1)


def make(oracle, rng, n, a, keep, scale=1.0):
    from psi4_b200 import DFHelper

    sp = oracle.Sparsity(keep, a)
    d = DFHelper(n, a)
    d.prepare_sparsity(keep=keep)
    B = sym_tensor(rng, a, n, scale)
    return sp, d, B, d.pack(B)


def check(got, ref, tol=TOL, what=""):
    for i, (g, r) in enumerate(zip(got, ref)):
        err = np.abs(g - r).max() if g.size else 0.0
        assert err < tol, f"{what}[{i}] max-abs {err:.3e}"


def test_dfjk_compare_like_reference(oracle):
    """Structure of tests/pytests/test_dfjk.py:12-74: H2O cc-pVDZ / cc-pVDZ-jkfit sizes (24 bf, 116 aux),
    five random spaces (16,16,20,20,30 columns), seven (C_left, C_right) pairs incl. non-symmetric ones,
    C_right always added => lr_symmetric false for every pair."""
    from psi4_b200 import MemDFJK

    rng = np.random.default_rng(2024)
    n, a = 24, 116
    keep = np.ones((n, n), bool)
    sp, d, B, P = make(oracle, rng, n, a, keep, 0.1)
    sizes = [16, 16, 20, 20, 30]
    spaces = [rng.random((n, s)) for s in sizes]
    pairs = [[0, 0], [0, 1], [1, 1], [2, 2], [3, 2], [3, 3], [4, 4]]
    jk = MemDFJK(d, P)
    jk.initialize()
    jk.print_header()
    for l, r in pairs:
        jk.C_left_add(spaces[l])
        jk.C_right_add(spaces[r])
    jk.compute()
    Cl = [spaces[l] for l, _ in pairs]
    Cr = [spaces[r] for _, r in pairs]
    J, K, _, _ = oracle.build_JK(sp, P, Cl, Cr)
    check(jk.J(), J, what="J")
    check(jk.K(), K, what="K")
    Jd, Kd, _ = oracle.dense_JK(B, keep, Cl, Cr)
    check(jk.J(), Jd, what="J(dense)")
    check(jk.K(), Kd, what="K(dense)")
    check(jk.D(), [x @ y.T for x, y in zip(Cl, Cr)], what="D")
    jk.finalize()


@pytest.mark.parametrize("lr", [True, False])
@pytest.mark.parametrize("density", [1.0, 0.55, 0.12])
def test_random_masks_ragged_nocc(oracle, lr, density):
    """Symmetric fast path (compute_J_symm, T2=T1) and general path; ragged nocc incl. 0 and odd."""
    from psi4_b200 import Engine

    rng = np.random.default_rng(int(density * 100) + lr)
    n, a = 61, 83
    keep = random_mask(rng, n, density)
    sp, d, B, P = make(oracle, rng, n, a, keep, 0.2)
    noccs = [7, 0, 1, 33]
    Cl = [rng.standard_normal((n, o)) for o in noccs]
    Cr = None if lr else [rng.standard_normal((n, o)) for o in noccs]
    D = [x @ (x if lr else y).T for x, y in zip(Cl, Cl if lr else Cr)]
    e = Engine(1)
    e.set_layout(n, a, d.small_skips_, d.big_skips_, d.schwarz_fun_index_)
    e.upload(0, P)
    J, K, _ = e.compute(Cl, Cr, D)
    Jo, Ko, _, _ = oracle.build_JK(sp, P, Cl, Cr, D=D)
    check(J, Jo, what="J")
    check(K, Ko, what="K")
    assert not K[1].any()
    if lr:
        for k in K:
            assert np.array_equal(k, k.T)  # mirrored triangle: exactly symmetric
    # toggles: J only / K only (scf_iterator.py:121 do_K=false for pure DFT)
    J2, K2, _ = e.compute(Cl, Cr, D, do_J=True, do_K=False)
    assert K2 is None
    check(J2, Jo, what="J-only")
    J3, K3, _ = e.compute(Cl, Cr, D, do_J=False, do_K=True)
    assert J3 is None
    check(K3, Ko, what="K-only")
    st = e.stats()
    assert st["launches"] > 0 and st["ms_total"] > 0
    e.close()


@pytest.mark.parametrize("lr", [True, False])
def test_medium_banded_and_q_chunking(oracle, lr):
    """Benzene-dimer-like size with a banded (chain-molecule) mask; then force the K build to loop over
    Q chunks with a small work budget (stands in for Qshell_blocks_for_JK_build, dfhelper.cc:814-869)."""
    from psi4_b200 import Engine

    rng = np.random.default_rng(77 + lr)
    n, a, o = 228, 300, 42
    keep = banded_mask(n, 70)
    sp, d, B, P = make(oracle, rng, n, a, keep, 0.05)
    Cl = [np.linalg.qr(rng.standard_normal((n, o)))[0]]
    Cr = None if lr else [np.linalg.qr(rng.standard_normal((n, o)))[0]]
    D = [Cl[0] @ (Cl[0] if lr else Cr[0]).T]
    Jo, Ko, _, _ = oracle.build_JK(sp, P, Cl, Cr, D=D)
    e = Engine(1)
    e.set_layout(n, a, d.small_skips_, d.big_skips_, d.schwarz_fun_index_)
    e.upload(0, P)
    J, K, _ = e.compute(Cl, Cr, D)
    check(J, Jo, what="J")
    check(K, Ko, what="K")
    e.set_work_budget(2 * 128 * n * o * 8 + 1024)  # room for one 128-row chunk of T1 and T2
    J, K, _ = e.compute(Cl, Cr, D)
    check(J, Jo, what="J(chunked)")
    check(K, Ko, what="K(chunked)")
    e.close()


@pytest.mark.parametrize("lr", [True, False])
def test_wK(oracle, lr):
    from psi4_b200 import MemDFJK

    rng = np.random.default_rng(5 + lr)
    n, a = 40, 57
    keep = random_mask(rng, n, 0.6)
    sp, d, B, P = make(oracle, rng, n, a, keep, 0.2)
    M1, W = sym_tensor(rng, a, n, 0.2), sym_tensor(rng, a, n, 0.2)
    P1, PW = d.pack(M1), d.pack(W)
    C = [rng.standard_normal((n, 6)), rng.standard_normal((n, 11))]
    C2 = [rng.standard_normal((n, 6)), rng.standard_normal((n, 11))]
    jk = MemDFJK(d, P, P1, PW)
    jk.set_do_wK(True)
    jk.set_omega(0.3)
    jk.initialize()
    for i, c in enumerate(C):
        jk.C_left_add(c)
        if not lr:
            jk.C_right_add(C2[i])
    jk.compute()
    Jo, Ko, wKo, _ = oracle.build_JK(sp, P, C, None if lr else C2, do_wK=True, m1Ppq=P1, wPpq=PW)
    check(jk.J(), Jo, what="J")
    check(jk.K(), Ko, what="K")
    check(jk.wK(), wKo, what="wK")
    if lr:
        assert np.array_equal(jk.wK()[0], jk.wK()[0].T)


def test_upload_rows_streaming_and_download(oracle):
    from psi4_b200 import Engine

    rng = np.random.default_rng(9)
    n, a = 33, 21
    keep = random_mask(rng, n, 0.5)
    sp, d, B, P = make(oracle, rng, n, a, keep)
    e = Engine(1)
    e.set_layout(n, a, d.small_skips_, d.big_skips_, d.schwarz_fun_index_)
    for m0 in range(0, n, 10):  # p-blocked construction loop of prepare_AO_core
        m1 = min(n, m0 + 10)
        e.upload_rows(0, m0, m1, P[int(d.big_skips_[m0]):int(d.big_skips_[m1])])
    for m in (0, 7, n - 1):
        got = e.download_rows(0, m, 3, 17)
        ref = P[int(d.big_skips_[m]):int(d.big_skips_[m + 1])].reshape(a, -1)[3:17]
        assert np.array_equal(got, ref)
    C = [rng.standard_normal((n, 4))]
    J, K, _ = e.compute(C, None, [C[0] @ C[0].T])
    Jo, Ko, _, _ = oracle.build_JK(sp, P, C)
    check(J, Jo)
    check(K, Ko)
    e.close()


def test_device_synthetic_fill_bit_exact(oracle):
    """b200jk_fill_synthetic must reproduce oracle_synth_fill bit for bit (integer hash + one multiply)."""
    from psi4_b200 import Engine, workloads

    n, a = 50, 37
    keep = workloads.banded_mask(n, 0.5, block=5)
    sp = oracle.Sparsity(keep, a)
    amp = workloads.amplitude(n)
    ref = oracle.synth_fill(sp, 0, a, 4242, amp)
    e = Engine(1)
    e.set_layout(n, a, sp.small_skips, sp.big_skips, sp.fun_index)
    e.fill_synthetic(0, 4242, amp)
    for m in range(n):
        got = e.download_rows(0, m, 0, a)
        assert np.array_equal(got.ravel(), ref[int(sp.big_skips[m]):int(sp.big_skips[m + 1])])
    e.close()


def test_device_operands_match_host_operands(oracle):
    from psi4_b200 import Engine

    rng = np.random.default_rng(21)
    n, a, o = 70, 45, 9
    keep = random_mask(rng, n, 0.8)
    sp, d, B, P = make(oracle, rng, n, a, keep, 0.3)
    C = rng.standard_normal((n, o))
    D = C @ C.T
    e = Engine(1)
    e.set_layout(n, a, d.small_skips_, d.big_skips_, d.schwarz_fun_index_)
    e.upload(0, P)
    Jh, Kh, _ = e.compute([C], None, [D])
    dC, dD = e.dev_put(C), e.dev_put(D)
    dJ, dK = e.dev_alloc(n * n * 8), e.dev_alloc(n * n * 8)
    e.compute_device([dC], None, [o], [dD], [dJ], [dK], None)
    assert np.array_equal(e.dev_get(dJ, (n, n)), Jh[0])
    assert np.array_equal(e.dev_get(dK, (n, n)), Kh[0])  # deterministic reductions: bitwise repeatable
    for p in (dC, dD, dJ, dK):
        e.dev_free(p)
    e.close()


def test_error_paths():
    from psi4_b200 import B200JKError, DFHelper, Engine, MemDFJK, PsiException

    e = Engine(1)
    with pytest.raises(B200JKError):
        e.compute([np.zeros((4, 2))], None, [np.zeros((4, 4))])  # before set_layout
    n, a = 6, 4
    keep = np.ones((n, n), bool)
    d = DFHelper(n, a)
    d.prepare_sparsity(keep=keep)
    bad = d.small_skips_.copy()
    bad[2] -= 1
    with pytest.raises(B200JKError):
        e.set_layout(n, a, bad, d.big_skips_, d.schwarz_fun_index_)
    e.set_layout(n, a, d.small_skips_, d.big_skips_, d.schwarz_fun_index_)
    with pytest.raises(B200JKError):
        e.compute([np.zeros((n, 2))], None, [np.zeros((n, n))])  # tensor not uploaded
    e.upload(0, np.zeros(int(d.big_skips_[n])))
    with pytest.raises(B200JKError):
        e.compute([np.zeros((n, 2))], None, [np.zeros((n, n))], do_wK=True)  # wK tensors missing
    e.close()
    jk = MemDFJK(d, np.zeros(int(d.big_skips_[n])))
    jk.initialize()
    jk.C_left_add(np.zeros((n, 2)))
    jk.C_right_add(np.zeros((n, 3)))
    with pytest.raises(PsiException):
        jk.compute()  # jk.cc:612-615 zip mismatch
    with pytest.raises(PsiException):
        jk.set_wcombine(True)


@pytest.mark.parametrize("nocc", [1, 2, 7, 8, 9, 63, 64, 65, 127, 128, 129, 200])
def test_nocc_sweep_tile_edges(oracle, nocc):
    """Every orbital-tile shape of the half transform (1..8 DMMA column blocks, 1 and 2 i-tiles, ragged last block)
    and K-GEMM k lengths that are not multiples of the stage depth."""
    from psi4_b200 import Engine

    rng = np.random.default_rng(1000 + nocc)
    n, a = 150, 37 + (nocc % 5)
    keep = random_mask(rng, n, 0.7) if nocc % 2 else np.ones((n, n), bool)
    sp, d, B, P = make(oracle, rng, n, a, keep, 0.1)
    C = [rng.standard_normal((n, nocc))]
    D = [C[0] @ C[0].T]
    e = Engine(1)
    e.set_layout(n, a, d.small_skips_, d.big_skips_, d.schwarz_fun_index_)
    e.upload(0, P)
    J, K, _ = e.compute(C, None, D)
    Jo, Ko, _, _ = oracle.build_JK(sp, P, C, D=D)
    scale = max(1.0, np.abs(Ko[0]).max())
    assert np.abs(J[0] - Jo[0]).max() < TOL * max(1.0, np.abs(Jo[0]).max())
    assert np.abs(K[0] - Ko[0]).max() < TOL * scale
    e.close()


@pytest.mark.parametrize("seed", range(8))
def test_random_shapes(oracle, seed):
    """Randomised shapes: nbf around the 128-row tile edges, naux below / across one q tile, random screening,
    symmetric and general paths, 1-3 densities with ragged nocc."""
    from psi4_b200 import Engine

    rng = np.random.default_rng(seed)
    n = int(rng.choice([5, 31, 127, 128, 129, 200, 257]))
    a = int(rng.choice([1, 3, 64, 127, 128, 129, 300]))
    dens = float(rng.choice([1.0, 0.8, 0.3, 0.05]))
    lr = bool(rng.integers(2))
    keep = random_mask(rng, n, dens)
    sp, d, B, P = make(oracle, rng, n, a, keep, 0.2)
    nmat = int(rng.integers(1, 4))
    noccs = [int(rng.integers(0, min(n, 40) + 1)) for _ in range(nmat)]
    Cl = [rng.standard_normal((n, o)) for o in noccs]
    Cr = None if lr else [rng.standard_normal((n, o)) for o in noccs]
    D = [x @ (x if lr else y).T for x, y in zip(Cl, Cl if lr else Cr)]
    e = Engine(1)
    e.set_layout(n, a, d.small_skips_, d.big_skips_, d.schwarz_fun_index_)
    e.upload(0, P)
    J, K, _ = e.compute(Cl, Cr, D)
    Jo, Ko, _, _ = oracle.build_JK(sp, P, Cl, Cr, D=D)
    for i in range(nmat):
        assert np.abs(J[i] - Jo[i]).max() < TOL * max(1.0, np.abs(Jo[i]).max()), (n, a, dens, lr, noccs)
        assert np.abs(K[i] - Ko[i]).max() < TOL * max(1.0, np.abs(Ko[i]).max()), (n, a, dens, lr, noccs)
    e.close()


def test_oom_is_an_error_not_a_fallback():
    """In-core only: a tensor that does not fit in HBM fails loudly (mirrors SCF_SUBTYPE=INCORE throwing,
    dfhelper.cc:259-262); nothing spills to the host."""
    from psi4_b200 import B200JKError, DFHelper, Engine

    n, a = 2048, 20000  # 2048^2 * 20000 * 8 B = 671 GB
    d = DFHelper(n, a)
    d.prepare_sparsity(keep=np.ones((n, n), bool))
    e = Engine(1)
    e.set_layout(n, a, d.small_skips_, d.big_skips_, d.schwarz_fun_index_)
    with pytest.raises(B200JKError) as ei:
        e.fill_synthetic(0, 1, np.ones((n, n)))
    assert ei.value.code == 3  # B200JK_ERR_OOM
    e.close()


def test_registered_host_buffers_same_result(oracle):
    """b200jk_register_host: D/J/K in page-locked caller memory are DMA'd in place; results identical to staging."""
    from psi4_b200 import Engine

    rng = np.random.default_rng(31)
    n, a, o = 80, 50, 11
    keep = random_mask(rng, n, 0.7)
    sp, d, B, P = make(oracle, rng, n, a, keep, 0.2)
    C = rng.standard_normal((n, o))
    D = C @ C.T
    e = Engine(1)
    e.set_layout(n, a, d.small_skips_, d.big_skips_, d.schwarz_fun_index_)
    e.upload(0, P)
    J0, K0, _ = e.compute([C], None, [D])
    e.register_host(D)
    J1, K1, _ = e.compute([C], None, [D], reuse_outputs=True)  # persistent, registered outputs
    assert np.array_equal(J0[0], J1[0]) and np.array_equal(K0[0], K1[0])
    J2, K2, _ = e.compute([2 * C], None, [4 * D], reuse_outputs=True)  # 4*D is a fresh unregistered array -> staged
    assert J2[0] is J1[0] and np.array_equal(J2[0], 4 * J0[0]) and np.array_equal(K2[0], 4 * K0[0])
    e.unregister_host(D)
    e.close()


@pytest.mark.parametrize("do_wK", [False, True])
def test_response_batch_shared_c_left(oracle, monkeypatch, do_wK):
    """The reference's response callers (twoel_Hx: `Cl.push_back(Co)` for every trial vector, libscf_solver/rhf.cc:466-484)
    hand ONE occupied block as every C_left and a different C_right per vector.  The engine keeps the first half
    transform across such matrices; the result must be the oracle's and bit-identical to the build that recomputes it,
    and the skipped transforms must not be credited as work."""
    from psi4_b200 import Engine

    rng = np.random.default_rng(909 + do_wK)
    n, a, o = 70, 64, 9
    keep = random_mask(rng, n, 0.6)
    sp, d, B, P = make(oracle, rng, n, a, keep, 0.2)
    Co = rng.standard_normal((n, o))
    # [shared object, shared object, equal copy (USO2AO-style), a different left block, shared with THAT one, other nocc]
    other = rng.standard_normal((n, o))
    Cl = [Co, Co, Co.copy(), other, other, rng.standard_normal((n, o + 2))]
    Cr = [rng.standard_normal(c.shape) for c in Cl]
    D = [x @ y.T for x, y in zip(Cl, Cr)]
    e = Engine(1)
    e.set_layout(n, a, d.small_skips_, d.big_skips_, d.schwarz_fun_index_)
    e.upload(0, P)
    kw = {}
    if do_wK:
        M1, W = sym_tensor(rng, a, n, 0.2), sym_tensor(rng, a, n, 0.2)
        P1, PW = d.pack(M1), d.pack(W)
        e.upload(1, P1)
        e.upload(2, PW)
        kw = dict(do_wK=True, m1Ppq=P1, wPpq=PW)
    J, K, wK = e.compute(Cl, Cr, D, do_wK=do_wK)
    st = e.stats()
    Jo, Ko, wKo, _ = oracle.build_JK(sp, P, Cl, Cr, D=D, **kw)
    check(J, Jo, what="J")
    check(K, Ko, what="K")
    if do_wK:
        check(wK, wKo, what="wK")
    monkeypatch.setenv("B200JK_NO_T1_REUSE", "1")
    J2, K2, wK2 = e.compute(Cl, Cr, D, do_wK=do_wK)
    st2 = e.stats()
    for x, y in zip(J + K + (wK if do_wK else []), J2 + K2 + (wK2 if do_wK else [])):
        assert np.array_equal(x, y)
    # 12 transforms (24 with wK) without reuse; matrices 1, 2 and 4 skip their first one
    ntr, skipped = (24, 6) if do_wK else (12, 3)
    assert st2["half_flops"] > 0
    # nocc differs for the last matrix, so compare through the per-transform unit 2*A*P*o
    unit = 2.0 * a * int(d.small_skips_[n])
    per_pass = 2 * unit * (5 * o + (o + 2))
    assert st2["half_flops"] == pytest.approx(per_pass * (2 if do_wK else 1), rel=1e-12)
    assert st["half_flops"] == pytest.approx(st2["half_flops"] - skipped * unit * o, rel=1e-12)
    assert ntr * 2 > skipped
    e.close()
