"""GPU parity tests (-m gpu) of the INT8-tensor-core arm of the K GEMM (psi4_b200/csrc/i8_kgemm.cuh): the same
C_DGEMM('N','T') of DFHelper::compute_K / compute_wK (lib3index/dfhelper.cc:3374, :3433) computed exactly modulo 13
coprime numbers on tcgen05.mma.kind::i8 and rebuilt by the Chinese remainder theorem.  Gate: the north_star's 1e-10
max-abs on K / wK against the CPU oracle and the reference's golden vectors, as for the DMMA arm."""
import os

import numpy as np
import pytest

from test_gpu_parity import banded_mask, check, make, random_mask

pytestmark = pytest.mark.gpu

TOL = 1e-10


def engine_for(d, P, n, a, arm, moduli=0, tensors=None):
    from psi4_b200 import Engine

    e = Engine(1)
    e.set_layout(n, a, d.small_skips_, d.big_skips_, d.schwarz_fun_index_)
    e.upload(0, P)
    for which, t in (tensors or {}).items():
        e.upload(which, t)
    e.set_kgemm(arm, moduli)
    return e


@pytest.mark.parametrize("lr", [True, False])
@pytest.mark.parametrize("density", [1.0, 0.4])
def test_i8_arm_matches_oracle_and_dmma_arm(oracle, lr, density):
    """Symmetric (T2 = T1, mirrored triangle) and general product; ragged nocc incl. 0 and odd; 300 basis functions so
    that the last 128-row block is ragged (44 rows) and both remainder column widths occur."""
    rng = np.random.default_rng(11 + lr + int(10 * density))
    n, a = 300, 150
    keep = random_mask(rng, n, density)
    sp, d, B, P = make(oracle, rng, n, a, keep, 0.1)
    noccs = [23, 0, 1, 64]
    Cl = [rng.standard_normal((n, o)) / np.sqrt(n) for o in noccs]  # K elements of order 1-100, as in an SCF
    Cr = None if lr else [rng.standard_normal((n, o)) / np.sqrt(n) for o in noccs]
    D = [x @ (x if lr else y).T for x, y in zip(Cl, Cl if lr else Cr)]
    Jo, Ko, _, _ = oracle.build_JK(sp, P, Cl, Cr, D=D)
    e = engine_for(d, P, n, a, "i8")
    J, K, _ = e.compute(Cl, Cr, D)
    st = e.stats()
    assert st["kgemm_kind"] == 1 and st["kgemm_moduli"] == 13, st
    check(J, Jo, what="J")
    check(K, Ko, what="K(i8)")
    assert not K[1].any()
    if lr:
        for k in K:
            assert np.array_equal(k, k.T)
    # run to run: integer arithmetic, bit-identical
    J2, K2, _ = e.compute(Cl, Cr, D)
    for x, y in zip(K, K2):
        assert np.array_equal(x, y)
    # the other arm on the same handle
    e.set_kgemm("dmma")
    _, Kd, _ = e.compute(Cl, Cr, D)
    assert e.stats()["kgemm_kind"] == 0
    check(K, Kd, tol=1e-11, what="K(i8) vs K(dmma)")
    e.close()


def test_i8_error_follows_the_number_of_moduli(oracle):
    """The only rounding of the residue arm is the scaling of the rows of T to integers: 2^-bits of the row norm, bits =
    (log2 of the product of the moduli - 1) / 2.  Fewer moduli, proportionally larger error; 13 moduli sit at the level
    of the DMMA arm's own rounding."""
    rng = np.random.default_rng(5)
    n, a, o = 256, 128, 40
    keep = np.ones((n, n), bool)
    sp, d, B, P = make(oracle, rng, n, a, keep, 0.1)
    Cl = [np.linalg.qr(rng.standard_normal((n, o)))[0]]
    _, Ko, _, _ = oracle.build_JK(sp, P, Cl, None)
    scale = np.sqrt(np.outer(np.diag(Ko[0]), np.diag(Ko[0]))).max()  # |T[m,:]| |T[n,:]| = sqrt(K_mm K_nn)
    e = engine_for(d, P, n, a, "i8")
    moduli = [256, 255, 253, 251, 247, 241, 239, 233, 229, 227, 223, 217, 211]
    errs = {}
    for nm in (8, 10, 12, 13):
        e.set_kgemm("i8", nm)
        _, K, _ = e.compute(Cl, None, None, do_J=False)
        assert e.stats()["kgemm_moduli"] == nm
        bits = (np.sum(np.log2(moduli[:nm])) - 1) / 2
        errs[nm] = np.abs(K[0] - Ko[0]).max()
        assert errs[nm] <= 8 * 2.0 ** -min(bits, 51) * scale + 1e-13, (nm, bits, errs[nm], scale)
    assert errs[8] > errs[10] > errs[12]
    assert errs[13] < 1e-12
    e.close()


def test_i8_wk_and_q_chunks(oracle):
    """wK (full square product of two different transforms, then hermitivitized) and the K build looping over Q chunks
    (beta = 1 accumulation across chunks, each chunk with its own row scales)."""
    rng = np.random.default_rng(99)
    n, a, o = 228, 300, 42
    keep = banded_mask(n, 90)
    sp, d, B, P = make(oracle, rng, n, a, keep, 0.05)
    B1 = rng.standard_normal((a, n, n)) * 0.05
    B1 = B1 + B1.transpose(0, 2, 1)
    B2 = rng.standard_normal((a, n, n)) * 0.05
    B2 = B2 + B2.transpose(0, 2, 1)
    m1, w = d.pack(B1), d.pack(B2)
    Cl = [np.linalg.qr(rng.standard_normal((n, o)))[0]]
    D = [Cl[0] @ Cl[0].T]
    Jo, Ko, wKo, _ = oracle.build_JK(sp, P, Cl, None, D=D, do_wK=True, m1Ppq=m1, wPpq=w)
    e = engine_for(d, P, n, a, "i8", tensors={1: m1, 2: w})
    J, K, wK = e.compute(Cl, None, D, do_wK=True)
    check(J, Jo, what="J")
    check(K, Ko, what="K")
    check(wK, wKo, what="wK")
    e.set_work_budget(2 * 128 * n * o * 8 + 1024)  # one 128-row chunk of T1 and T2 at a time
    J, K, wK = e.compute(Cl, None, D, do_wK=True)
    check(K, Ko, what="K chunked")
    check(wK, wKo, what="wK chunked")
    assert e.stats()["kgemm_kind"] == 1
    e.close()


def test_i8_rows_of_very_different_size(oracle):
    """Every row of T carries its own power-of-two scale: rows eight orders of magnitude apart keep their relative
    accuracy (a common scale would wipe out the small ones)."""
    rng = np.random.default_rng(3)
    n, a, o = 192, 100, 30
    keep = np.ones((n, n), bool)
    sp, d, B, P = make(oracle, rng, n, a, keep, 0.1)
    w = 10.0 ** rng.uniform(-4, 4, n)  # basis-function "sizes"
    Bs = B * w[None, :, None] * w[None, None, :]
    Ps = d.pack(Bs)
    Cl = [np.linalg.qr(rng.standard_normal((n, o)))[0] / w[:, None]]
    _, Ko, _, _ = oracle.build_JK(sp, Ps, Cl, None)
    e = engine_for(d, Ps, n, a, "i8")
    _, K, _ = e.compute(Cl, None, None, do_J=False)
    rel = np.abs(K[0] - Ko[0]) / np.sqrt(np.outer(np.diag(Ko[0]), np.diag(Ko[0])))
    assert rel.max() < 1e-13, rel.max()
    e.close()


def test_i8_golden_vectors():
    """The reference's own object code froze these J / K / wK (tests/golden/reference_jk_vectors.npz, made by
    tools/make_golden_jk.py); the residue arm is held to the same 1e-10 as the DMMA arm (tests/test_golden_vectors.py)."""
    from psi4_b200 import DFHelper, Engine

    g = np.load(os.path.join(os.path.dirname(__file__), "golden", "reference_jk_vectors.npz"))
    keep = g["keep"]
    n, a = keep.shape[0], int(g["naux"])
    d = DFHelper(n, a)
    d.prepare_sparsity(keep=keep)
    e = Engine(1)
    e.set_layout(n, a, d.small_skips_, d.big_skips_, d.schwarz_fun_index_)
    e.upload(0, g["Ppq"])
    e.upload(1, g["m1Ppq"])
    e.upload(2, g["wPpq"])
    e.set_kgemm("i8")
    nmat = len(g["noccs"])
    for tag, lr in (("sym", True), ("gen", False)):
        Cl = [g[f"Cl{i}"] for i in range(nmat)]
        Cr = None if lr else [g[f"Cr{i}"] for i in range(nmat)]
        D = [x @ (x if lr else y).T for x, y in zip(Cl, Cl if lr else Cr)]
        J, K, wK = e.compute(Cl, Cr, D, do_wK=True)
        assert e.stats()["kgemm_kind"] == 1
        for i in range(nmat):
            assert np.abs(K[i] - g[f"K_{tag}{i}"]).max() < TOL
            assert np.abs(wK[i] - g[f"wK_{tag}{i}"]).max() < TOL
            assert np.abs(J[i] - g[f"J_{tag}{i}"]).max() < TOL
    e.close()


# ---- the half transform on the INT8 tensor cores (psi4_b200/csrc/i8_half.cuh) -------------------------------------------


@pytest.mark.parametrize("lr", [True, False])
@pytest.mark.parametrize("density", [1.0, 0.4])
@pytest.mark.parametrize("kgemm", ["dmma", "i8"])
def test_i8_half_transform_matches_oracle(oracle, lr, density, kgemm):
    """DFHelper::first_transform_pQq (lib3index/dfhelper.cc:2162-2186) by residues: screened and dense masks, ragged nocc
    incl. 0 / 1 / odd, two different C matrices, with either arm of the K GEMM behind it.  The first J sweep is a column of
    the residue GEMM where the tiling has a free one (nocc 23, 1), its own kernel otherwise (nocc 64)."""
    rng = np.random.default_rng(21 + lr + int(10 * density))
    n, a = 300, 150
    keep = random_mask(rng, n, density)
    sp, d, B, P = make(oracle, rng, n, a, keep, 0.1)
    noccs = [23, 0, 1, 64]
    Cl = [rng.standard_normal((n, o)) / np.sqrt(n) for o in noccs]
    Cr = None if lr else [rng.standard_normal((n, o)) / np.sqrt(n) for o in noccs]
    D = [x @ (x if lr else y).T for x, y in zip(Cl, Cl if lr else Cr)]
    Jo, Ko, _, _ = oracle.build_JK(sp, P, Cl, Cr, D=D)
    e = engine_for(d, P, n, a, kgemm)
    e.set_half("i8")
    J, K, _ = e.compute(Cl, Cr, D)
    st = e.stats()
    assert st["half_kind"] == 1, st
    check(J, Jo, what="J")
    check(K, Ko, what="K(half i8)")
    assert not K[1].any()
    J2, K2, _ = e.compute(Cl, Cr, D)
    for x, y in zip(K, K2):
        assert np.array_equal(x, y)  # run to run
    e.set_half("dmma")
    _, Kd, _ = e.compute(Cl, Cr, D)
    assert e.stats()["half_kind"] == 0
    check(K, Kd, tol=1e-11, what="half i8 vs half dmma")
    e.close()


def test_i8_half_wk_q_chunks_wide_nocc_and_tensor_change(oracle):
    """wK (the other two tensors carry their own row scales), the K build in Q chunks, more than 256 occupied columns (two
    orbital tiles), and a tensor that is uploaded again with other values (the row scales must follow it)."""
    rng = np.random.default_rng(98)
    n, a, o = 300, 260, 270
    keep = banded_mask(n, 120)
    sp, d, B, P = make(oracle, rng, n, a, keep, 0.05)
    B1 = rng.standard_normal((a, n, n)) * 0.05
    B1 = B1 + B1.transpose(0, 2, 1)
    B2 = rng.standard_normal((a, n, n)) * 0.05
    B2 = B2 + B2.transpose(0, 2, 1)
    m1, w = d.pack(B1), d.pack(B2)
    Cl = [np.linalg.qr(rng.standard_normal((n, o)))[0], np.linalg.qr(rng.standard_normal((n, 37)))[0]]
    D = [c @ c.T for c in Cl]
    Jo, Ko, wKo, _ = oracle.build_JK(sp, P, Cl, None, D=D, do_wK=True, m1Ppq=m1, wPpq=w)
    e = engine_for(d, P, n, a, "i8", tensors={1: m1, 2: w})
    e.set_half("i8")
    J, K, wK = e.compute(Cl, None, D, do_wK=True)
    assert e.stats()["half_kind"] == 1
    check(J, Jo, what="J")
    check(K, Ko, what="K")
    check(wK, wKo, what="wK")
    e.set_work_budget(2 * 128 * n * o * 8 + 1024)  # one 128-row chunk of T1 and T2 at a time
    J, K, wK = e.compute(Cl, None, D, do_wK=True)
    check(K, Ko, what="K chunked")
    check(wK, wKo, what="wK chunked")
    e.set_work_budget(0)
    e.upload(0, 1024.0 * P)  # same handle, tensor 1024 times larger: K grows by 2^20 exactly if the scales were recomputed
    _, K3, _ = e.compute(Cl, None, None, do_J=False)
    for x, y in zip(K3, Ko):
        assert np.abs(x / 1048576.0 - y).max() < TOL
    e.close()


def test_i8_half_golden_vectors():
    """Both residue arms together against the vectors frozen from the reference's own object code."""
    from psi4_b200 import DFHelper, Engine

    g = np.load(os.path.join(os.path.dirname(__file__), "golden", "reference_jk_vectors.npz"))
    keep = g["keep"]
    n, a = keep.shape[0], int(g["naux"])
    d = DFHelper(n, a)
    d.prepare_sparsity(keep=keep)
    e = Engine(1)
    e.set_layout(n, a, d.small_skips_, d.big_skips_, d.schwarz_fun_index_)
    e.upload(0, g["Ppq"])
    e.upload(1, g["m1Ppq"])
    e.upload(2, g["wPpq"])
    e.set_kgemm("i8")
    e.set_half("i8", 13)
    nmat = len(g["noccs"])
    for tag, lr in (("sym", True), ("gen", False)):
        Cl = [g[f"Cl{i}"] for i in range(nmat)]
        Cr = None if lr else [g[f"Cr{i}"] for i in range(nmat)]
        D = [x @ (x if lr else y).T for x, y in zip(Cl, Cl if lr else Cr)]
        J, K, wK = e.compute(Cl, Cr, D, do_wK=True)
        st = e.stats()
        assert st["kgemm_kind"] == 1 and st["half_kind"] == 1
        for i in range(nmat):
            assert np.abs(K[i] - g[f"K_{tag}{i}"]).max() < TOL
            assert np.abs(wK[i] - g[f"wK_{tag}{i}"]).max() < TOL
            assert np.abs(J[i] - g[f"J_{tag}{i}"]).max() < TOL
    e.close()


def test_i8_first_j_sweep_rides_on_the_gemm(oracle):
    """When the residue planes of the shard stay resident, the first J sweep d_Q = B_Q . D (DGEMV 'N', dfhelper.cc:3193 /
    :3258) is one more column of the residue GEMM (I8HalfFuseJ::gemm_col): exact in the integers, so the build that
    converts the planes and the builds that find them cached give the same bits, for RHF-like and for general densities,
    and the tensor is read once less (j_bytes counts one pass, the second sweep)."""
    rng = np.random.default_rng(77)
    n, a = 300, 200
    keep = random_mask(rng, n, 0.5)
    sp, d, B, P = make(oracle, rng, n, a, keep, 0.1)
    noccs = [23, 40]
    Cl = [np.linalg.qr(rng.standard_normal((n, o)))[0] for o in noccs]
    D = [2.0 * c @ c.T for c in Cl]
    Jo, Ko, _, _ = oracle.build_JK(sp, P, Cl, None, D=D)
    e = engine_for(d, P, n, a, "i8")
    e.set_half("i8")
    J1, K1, _ = e.compute(Cl, None, D)
    st = e.stats()
    assert st["half_kind"] == 1 and st["half_i8_cached"] == len(noccs) - 1, st  # the second density finds the planes of the first
    check(J1, Jo, what="J (first build)")
    check(K1, Ko, what="K")
    ptri = sum(int(np.count_nonzero(keep[m, m:])) for m in range(n))
    assert st["j_bytes"] == 8.0 * a * ptri, (st["j_bytes"], 8.0 * a * ptri)  # one pass: the batched second sweep
    J2, K2, _ = e.compute(Cl, None, D)
    assert e.stats()["half_i8_cached"] == len(noccs)
    for x, y in zip(J1 + K1, J2 + K2):
        assert np.array_equal(x, y)
    # a general (non-symmetric) density pair through the same column
    Cr = [np.linalg.qr(rng.standard_normal((n, o)))[0] for o in noccs]
    Dg = [x @ y.T for x, y in zip(Cl, Cr)]
    Jo, Ko, _, _ = oracle.build_JK(sp, P, Cl, Cr, D=Dg)
    J, K, _ = e.compute(Cl, Cr, Dg)
    check(J, Jo, what="J general")
    check(K, Ko, what="K general")
    e.close()


RESIDENT_SCRIPT = r"""
import os, sys
import numpy as np
sys.path.insert(0, os.environ["B2_ROOT"]); sys.path.insert(0, os.path.join(os.environ["B2_ROOT"], "oracle"))
import dfjk_oracle as oracle
from psi4_b200 import DFHelper, Engine
rng = np.random.default_rng(31)
n, a = 300, 200
r = rng.random((n, n)); keep = (r + r.T) < 1.2; np.fill_diagonal(keep, True)
d = DFHelper(n, a); d.prepare_sparsity(keep=keep)
B = rng.standard_normal((a, n, n)) * 0.1; B = B + B.transpose(0, 2, 1)
P = d.pack(B)
Cl = [np.linalg.qr(rng.standard_normal((n, o)))[0] for o in (40, 33)]
D = [2.0 * c @ c.T for c in Cl]
Jo, Ko, _, _ = oracle.build_JK(oracle.Sparsity(keep, a), P, Cl, D=D)
e = Engine(1); e.set_layout(n, a, d.small_skips_, d.big_skips_, d.schwarz_fun_index_); e.upload(0, P)
e.set_half("i8"); e.set_kgemm("i8")
res = []
for it in range(3):
    J, K, _ = e.compute(Cl, None, D)
    st = e.stats()
    res.append(([x.copy() for x in J], [x.copy() for x in K], st))
    err = max(max(np.abs(x - y).max() for x, y in zip(J, Jo)), max(np.abs(x - y).max() for x, y in zip(K, Ko)))
    print(f"build {it}: err {err:.3e} resident {st['half_i8_resident_rows']} chunks {st['half_i8_chunks']} "
          f"convert_bytes {st['half_i8_convert_bytes']:.3e} cached {st['half_i8_cached']}")
    assert err < 1e-10 and st["half_kind"] == 1
want = os.environ["EXPECT"]
rows = res[0][2]["half_i8_resident_rows"]
if want == "partial":
    assert 0 < rows < n, rows
    assert res[0][2]["half_i8_chunks"] > 2 * 2          # several chunks per transform
    # the second build converts only what is not resident; the first one converted everything (twice: the resident part is
    # valid for the second density already)
    assert 0 < res[1][2]["half_i8_convert_bytes"] < res[0][2]["half_i8_convert_bytes"]
elif want == "none":
    assert rows == 0 and res[1][2]["half_i8_convert_bytes"] == res[0][2]["half_i8_convert_bytes"] > 0
else:
    assert rows == n and res[1][2]["half_i8_convert_bytes"] == 0
for it in (1, 2):  # resident or not, converted in this build or in an earlier one: the same bits
    for x, y in zip(res[0][0] + res[0][1], res[it][0] + res[it][1]):
        assert np.array_equal(x, y)
# a new tensor invalidates the resident planes
P2 = d.pack(B * 1.5)
e.upload(0, P2)
J, K, _ = e.compute(Cl, None, D)
Jo2, Ko2, _, _ = oracle.build_JK(oracle.Sparsity(keep, a), P2, Cl, D=D)
err = max(max(np.abs(x - y).max() for x, y in zip(J, Jo2)), max(np.abs(x - y).max() for x, y in zip(K, Ko2)))
print(f"new tensor: err {err:.3e}")
assert err < 1e-10
e.close()
"""


@pytest.mark.parametrize("env,expect", [({"B200JK_I8_ARENA_MB": "150"}, "partial"),
                                        ({"B200JK_I8_ARENA_MB": "150", "B200JK_I8_RESIDENT": "0"}, "none"),
                                        ({"B200JK_I8_ARENA_MB": "60", "B200JK_I8_JCOL": "0"}, "partial"),
                                        ({}, "all")])
def test_i8_half_partly_resident_planes(tmp_path, env, expect):
    """The scratch arena of the residue half transform holds the planes of the first row-blocks for good and converts the
    rest in every build (i8_half_run: resident region + work region).  Forced here by a small arena on a small system:
    J / K against the oracle over three builds and two densities, bit-identical across builds whatever was resident,
    with the A/B switches (nothing resident; the first J sweep as its own kernel), and after the tensor is replaced."""
    import subprocess
    import sys

    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    script = tmp_path / "resident.py"
    script.write_text(RESIDENT_SCRIPT)
    r = subprocess.run([sys.executable, str(script)], env=dict(os.environ, B2_ROOT=root, EXPECT=expect, **env), capture_output=True,
                       text=True, timeout=600)
    assert r.returncode == 0, r.stdout[-2500:] + r.stderr[-2500:]
