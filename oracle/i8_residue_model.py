"""TEST INFRASTRUCTURE ONLY -- a numpy / Python-integer restatement of the ARITHMETIC of the engine's INT8-tensor-core arms
(psi4_b200/csrc/i8_kgemm.cuh, i8_half.cuh), used by tests/test_i8_model_cpu.py to check, on the CPU,

  * the row bound R and the power-of-two row scales (i8_row_bound, i8_rowscale_kernel / i8h_rowscale_kernel),
  * the residues (i8_convert_kernel: symmetric representatives, low byte for p = 256),
  * the per-modulus integer products reduced mod p (the epilogue of the tcgen05 GEMMs: i8_mod_acc),
  * both reconstructions -- Garner / exact integers (i8_crt_value) and the floating-point one (i8_fill_crt_fast,
    i8_crt_value_fast) --

against exact integer arithmetic, the error bounds DESIGN.md section 3c states, and the golden J / K vectors the
reference's own object code produced (tests/golden/reference_jk_vectors.npz).  It mirrors what the CUDA kernels compute
for the two contractions of DFHelper::compute_K (lib3index/dfhelper.cc:3374, C_DGEMM('N','T')) and
DFHelper::first_transform_pQq (:2162-2186); it is not a fast path and nothing in psi4_b200/ imports it.
"""
from __future__ import annotations

import math

import numpy as np

MODULI = [256, 255, 253, 251, 247, 241, 239, 233, 229, 227, 223, 217, 211]  # kI8Moduli, pairwise coprime


def modulus_product(nmod: int) -> int:
    return math.prod(MODULI[:nmod])


def row_bound(nmod: int, kdim: int) -> float:
    """i8_row_bound: rows are scaled to 2-norm <= R with R^2 < M / 2 (Cauchy-Schwarz keeps every dot product below M / 2),
    R <= 2^51 (the residues are taken in double arithmetic), minus the growth of the norm under rounding to integers."""
    M = modulus_product(nmod)
    R = math.sqrt(M // 2) * (1.0 - 1e-9)
    R = min(R, 2.0 ** 51)
    return R - (0.5 * math.sqrt(kdim) + 1.0)


def bits(nmod: int, kdim: int = 1) -> float:
    return math.log2(row_bound(nmod, kdim))


def row_exponents(X: np.ndarray, Rb: float) -> np.ndarray:
    """i8_rowscale_kernel: the largest e with |X[m,:]| 2^e <= Rb (0 for a zero row), clamped to +-1000."""
    out = np.zeros(X.shape[0], dtype=np.int64)
    for m in range(X.shape[0]):
        nrm = math.sqrt(float(np.dot(X[m], X[m]))) * (1.0 + 1e-12)
        if nrm > 0 and math.isfinite(nrm):
            _, t = math.frexp(Rb / nrm)
            ex = t - 1
            if math.ldexp(nrm, ex) > Rb:
                ex -= 1
            out[m] = max(-1000, min(1000, ex))
    return out


def quantize(X: np.ndarray, e: np.ndarray) -> np.ndarray:
    """rint(X * 2^e) row by row: exact integers below 2^51 in magnitude, kept in int64."""
    return np.rint(np.ldexp(X, e[:, None].astype(np.int64))).astype(np.int64)


def residues(Y: np.ndarray, nmod: int) -> np.ndarray:
    """i8_convert_kernel: planes [nmod][rows][k] of int8 residues -- the low byte for p = 256, y - p rint(y / p) otherwise."""
    planes = np.empty((nmod,) + Y.shape, dtype=np.int8)
    planes[0] = (Y & 0xFF).astype(np.uint8).view(np.int8)
    for j in range(1, nmod):
        p = MODULI[j]
        r = Y - p * np.rint(Y.astype(np.float64) * (1.0 / p)).astype(np.int64)
        assert np.abs(r).max(initial=0) <= 128
        planes[j] = (r & 0xFF).astype(np.uint8).view(np.int8)  # (y - p q) mod 256, as the kernel packs it
    return planes


def modular_products(PA: np.ndarray, PB: np.ndarray) -> np.ndarray:
    """Per modulus the int8 x int8 -> int32 GEMM  A B^T  (exact), reduced to [0, p): what the epilogues store as bytes."""
    nmod = PA.shape[0]
    out = np.empty((nmod, PA.shape[1], PB.shape[1]), dtype=np.int64)
    for j in range(nmod):
        acc = PA[j].astype(np.int64) @ PB[j].astype(np.int64).T
        assert np.abs(acc).max(initial=0) < 2 ** 31  # int32 accumulators: k-ranges of <= 65536
        out[j] = np.mod(acc, MODULI[j])
    return out


def crt_exact(res: np.ndarray) -> np.ndarray:
    """The integer in (-M/2, M/2] with the given residues (object array of Python ints): Garner's algorithm."""
    nmod = res.shape[0]
    M = modulus_product(nmod)
    coef = []
    for j in range(nmod):
        Mj = M // MODULI[j]
        coef.append(Mj * pow(Mj % MODULI[j], -1, MODULI[j]))
    flat = res.reshape(nmod, -1)
    out = np.empty(flat.shape[1], dtype=object)
    for i in range(flat.shape[1]):
        x = sum(int(flat[j, i]) * coef[j] for j in range(nmod)) % M
        out[i] = x - M if x > M // 2 else x
    return out.reshape(res.shape[1:])


def crt_fast_constants(nmod: int):
    """i8_fill_crt_fast: q_j = (M / p_j)^-1 mod p_j (symmetric), 1 / p_j = ih_j + il_j with ih_j a multiple of 2^-32."""
    M = modulus_product(nmod)
    q, ih, il = [], [], []
    for j in range(nmod):
        p = MODULI[j]
        qq = pow((M // p) % p, -1, p)
        if qq > p // 2:
            qq -= p
        q.append(qq)
        h = round(2 ** 32 / p) / 2 ** 32  # rintl(inv * 2^32) / 2^32, exact in double
        ih.append(h)
        # inv - ih in extended precision, rounded to double: via exact rationals
        from fractions import Fraction

        il.append(float(Fraction(1, p) - Fraction(h)))
    return np.array(q, dtype=np.int64), np.array(ih), np.array(il), float(M)


def crt_fast(res: np.ndarray) -> np.ndarray:
    """i8_crt_value_fast: value = M * frac(sum_j r_j q_j / p_j); the high parts are multiples of 2^-32 below 2^7 and sum
    exactly in double, the low parts carry the rest (the kernel uses FMAs for them: at most an ulp of a 2^-14 quantity apart)."""
    nmod = res.shape[0]
    q, ih, il, Md = crt_fast_constants(nmod)
    fh = np.zeros(res.shape[1:])
    fl = np.zeros(res.shape[1:])
    for j in range(nmod):
        sd = (res[j] * q[j]).astype(np.float64)
        fh = fh + sd * ih[j]
        fl = fl + sd * il[j]
    fh = fh - np.rint(fh)
    f = fh + fl
    f = f - np.rint(f)
    return f * Md


def matmul_nt(A: np.ndarray, B: np.ndarray, nmod: int, fast_crt: bool = False) -> np.ndarray:
    """A B^T by residues: scale the rows of both operands, multiply modulo nmod moduli, rebuild, scale back."""
    kdim = A.shape[1]
    Rb = row_bound(nmod, kdim)
    ea, eb = row_exponents(A, Rb), row_exponents(B, Rb)
    res = modular_products(residues(quantize(A, ea), nmod), residues(quantize(B, eb), nmod))
    if fast_crt:
        val = crt_fast(res)
    else:
        val = crt_exact(res).astype(np.float64)  # one rounding of the exact integer, as the 128-bit -> double conversion
    return np.ldexp(val, -(ea[:, None] + eb[None, :]).astype(np.int64))


def build_K(B_dense: np.ndarray, keep: np.ndarray, Cl: np.ndarray, Cr: np.ndarray | None = None, nmod_half: int = 12,
            nmod_k: int = 13) -> np.ndarray:
    """K of DFHelper::compute_K through both residue arms: T[m, q, i] = sum_{n kept} B[q, m, n] C[n, i] per row-block m with the
    floating-point CRT (K3), then K[m, n] = sum_{q i} T1[m, qi] T2[n, qi] with the exact one (K4)."""
    naux, nbf, _ = B_dense.shape

    def half(C):
        o = C.shape[1]
        T = np.zeros((nbf, naux, o))
        if o == 0:
            return T
        for m in range(nbf):
            cols = np.flatnonzero(keep[m])
            A = np.ascontiguousarray(B_dense[:, m, cols])      # rows (m, q), columns the kept partners
            Ct = np.ascontiguousarray(C[cols, :].T)
            # the engine scales the columns of C by their FULL norm (a bound for every kept-partner subset)
            Rb = row_bound(nmod_half, nbf)
            ea = row_exponents(A, Rb)
            ec = row_exponents(np.ascontiguousarray(C.T), Rb)
            res = modular_products(residues(quantize(A, ea), nmod_half), residues(quantize(Ct, ec), nmod_half))
            T[m] = np.ldexp(crt_fast(res), -(ea[:, None] + ec[None, :]).astype(np.int64))
        return T

    T1 = half(Cl)
    T2 = T1 if Cr is None else half(Cr)
    if Cl.shape[1] == 0:
        return np.zeros((nbf, nbf))
    return matmul_nt(T1.reshape(nbf, -1), T2.reshape(nbf, -1), nmod_k)


def build_J(B_dense: np.ndarray, keep: np.ndarray, D: np.ndarray, symmetric: bool, nmod: int = 12) -> np.ndarray:
    """J of DFHelper::compute_J_symm / compute_J (lib3index/dfhelper.cc:3162-3283) with the FIRST sweep taken through the
    residue arm, as the engine does when the sweep rides on the K3 GEMM (I8HalfFuseJ::gemm_col): per row-block m
        d_part[m][q] = sum_{n kept} B[q, m, n] D'[m, n],   D' = the density row as j_prep_dm_kernel prepares it
    (symmetric: 2 D above the diagonal, D on it, 0 below; general: D), d_Q = sum_m d_part[m][q] in row-block order, and the
    second sweep J[m, n] = sum_Q B[Q, m, n] d_Q in double (j_mn_kernel), mirrored when symmetric."""
    naux, nbf, _ = B_dense.shape
    Dp = np.array(D, dtype=np.float64)
    if symmetric:
        Dp = np.triu(2.0 * Dp, 1) + np.diag(np.diag(D))
    Rb = row_bound(nmod, nbf)
    eD = row_exponents(Dp, Rb)  # full row norm, as the engine takes it
    d = np.zeros(naux)
    for m in range(nbf):
        cols = np.flatnonzero(keep[m])
        A = np.ascontiguousarray(B_dense[:, m, cols])
        ea = row_exponents(A, Rb)
        drow = Dp[m:m + 1, cols]
        res = modular_products(residues(quantize(A, ea), nmod), residues(quantize(drow, eD[m:m + 1]), nmod))
        d += np.ldexp(crt_fast(res)[:, 0], -(ea + eD[m]).astype(np.int64))
    J = np.zeros((nbf, nbf))
    for m in range(nbf):
        for n in np.flatnonzero(keep[m]):
            if symmetric and n < m:
                continue
            J[m, n] = float(np.dot(B_dense[:, m, n], d))
            if symmetric:
                J[n, m] = J[m, n]
    return J
