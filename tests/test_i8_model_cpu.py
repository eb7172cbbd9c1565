"""CPU checks of the arithmetic of the INT8-tensor-core arms on its numpy / Python-integer restatement
oracle/i8_residue_model.py: the reconstruction is exact, the only rounding is the scaling of the operands with the stated
number of bits, the floating-point CRT agrees with the exact one far below that, and the reference's golden K vectors
(produced by its own object code, tools/make_golden_jk.py) are reproduced to the north_star's 1e-10."""
import math
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "oracle"))
sys.path.insert(0, ROOT)

import i8_residue_model as model  # noqa: E402


def test_moduli_are_pairwise_coprime_and_bits_are_as_documented():
    for i, p in enumerate(model.MODULI):
        for q in model.MODULI[:i]:
            assert math.gcd(p, q) == 1
    # b200jk.h / DESIGN.md: 50.7 bits at 13 moduli (capped by the 2^51 of the double residues), 46.9 at 12, 43.0 at 11
    assert abs(model.bits(13, 1800) - 50.7) < 0.35
    assert abs(model.bits(12, 1800) - 46.9) < 0.1
    assert abs(model.bits(11, 1800) - 43.0) < 0.1
    assert model.row_bound(13, 1) <= 2.0 ** 51


@pytest.mark.parametrize("nmod", [6, 12, 13])
def test_crt_exact_and_fast_on_random_integers(nmod):
    rng = np.random.default_rng(nmod)
    M = model.modulus_product(nmod)
    xs = [int(rng.integers(-2 ** 62, 2 ** 62)) * int(rng.integers(1, 2 ** 30)) % M for _ in range(400)]
    xs = [x - M if x > M // 2 else x for x in xs] + [0, 1, -1, M // 2, -(M // 2) + 1]
    res = np.array([[x % p for x in xs] for p in model.MODULI[:nmod]], dtype=np.int64)
    exact = model.crt_exact(res)
    assert [int(v) for v in exact] == xs
    fast = model.crt_fast(res)
    # one double rounding of the value (2^-53 relative) on top of a floor of ~2^-67 M from the low parts of the fraction
    for f, x in zip(fast, xs):
        if abs(x) < M // 2 - 2 ** 40:  # away from the wrap-around at +-M/2
            assert abs(float(f) - x) <= 2.0 ** -51 * abs(x) + 2.0 ** -62 * M, (f, x, nmod)


@pytest.mark.parametrize("nmod", [12, 13])
def test_product_is_exact_in_the_integers_and_rounding_is_only_the_scaling(nmod):
    rng = np.random.default_rng(100 + nmod)
    m, n, k = 24, 20, 300
    A = rng.standard_normal((m, k)) * 10.0 ** rng.uniform(-3, 3, (m, 1))  # rows of very different size
    B = rng.standard_normal((n, k)) * 10.0 ** rng.uniform(-3, 3, (n, 1))
    Rb = model.row_bound(nmod, k)
    ea, eb = model.row_exponents(A, Rb), model.row_exponents(B, Rb)
    Ya, Yb = model.quantize(A, ea), model.quantize(B, eb)
    assert np.sqrt((Ya.astype(float) ** 2).sum(1)).max() <= Rb + 0.5 * math.sqrt(k) + 1
    # exact integer product of the quantised operands, in Python integers
    want = np.array([[sum(int(a) * int(b) for a, b in zip(Ya[i], Yb[j])) for j in range(n)] for i in range(m)], dtype=object)
    res = model.modular_products(model.residues(Ya, nmod), model.residues(Yb, nmod))
    got = model.crt_exact(res)
    assert (got == want).all()
    # against the real product: norm-wise 2^-bits
    C = model.matmul_nt(A, B, nmod)
    ref = np.array([[math.fsum(A[i] * B[j]) for j in range(n)] for i in range(m)])
    scale = np.outer(np.sqrt((A ** 2).sum(1)), np.sqrt((B ** 2).sum(1)))
    rel = np.abs(C - ref) / scale
    assert rel.max() <= 4 * 2.0 ** -model.bits(nmod, k), (rel.max(), nmod)
    Cf = model.matmul_nt(A, B, nmod, fast_crt=True)
    assert (np.abs(Cf - C) / scale).max() <= 2.0 ** -52  # the floating-point CRT: one more double rounding of the value


def test_residue_arms_reproduce_the_reference_golden_vectors():
    """K of the reference's object code (tests/golden/reference_jk_vectors.npz) through the model of BOTH arms -- half
    transform by 12 moduli with the floating-point CRT, K GEMM by 13 with the exact one -- symmetric and general pairs,
    nocc 6 / 0 / 3 / 17 on a 40 %-screened mask: the same 1e-10 the GPU tests hold the kernels to."""
    from psi4_b200 import DFHelper

    g = np.load(os.path.join(ROOT, "tests", "golden", "reference_jk_vectors.npz"))
    keep = g["keep"].astype(bool)
    n, a = keep.shape[0], int(g["naux"])
    d = DFHelper(n, a)
    d.prepare_sparsity(keep=keep)
    B = d.unpack(g["Ppq"])
    worst = 0.0
    for i in range(len(g["noccs"])):
        Cl, Cr = g[f"Cl{i}"], g[f"Cr{i}"]
        K = model.build_K(B, keep, Cl)
        worst = max(worst, float(np.abs(K - g[f"K_sym{i}"]).max()))
        K = model.build_K(B, keep, Cl, Cr)
        worst = max(worst, float(np.abs(K - g[f"K_gen{i}"]).max()))
    assert worst < 1e-10, worst


def test_first_j_sweep_through_the_residue_arm_reproduces_the_golden_j():
    """The first J sweep as a column of the K3 residue GEMM (12 moduli, floating-point CRT, the density row scaled by its
    full norm) followed by the FP64 second sweep: J of the reference's object code to 1e-10, symmetric and general."""
    from psi4_b200 import DFHelper

    g = np.load(os.path.join(ROOT, "tests", "golden", "reference_jk_vectors.npz"))
    keep = g["keep"].astype(bool)
    n, a = keep.shape[0], int(g["naux"])
    d = DFHelper(n, a)
    d.prepare_sparsity(keep=keep)
    B = d.unpack(g["Ppq"])
    worst = 0.0
    for i in range(len(g["noccs"])):
        Cl, Cr = g[f"Cl{i}"], g[f"Cr{i}"]
        worst = max(worst, float(np.abs(model.build_J(B, keep, Cl @ Cl.T, True) - g[f"J_sym{i}"]).max()))
        worst = max(worst, float(np.abs(model.build_J(B, keep, Cl @ Cr.T, False) - g[f"J_gen{i}"]).max()))
    assert worst < 1e-10, worst
