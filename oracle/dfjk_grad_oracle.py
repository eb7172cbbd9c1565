"""CPU restatement of the density-fitted SCF gradient contractions of the reference (DFJKGrad, psi4/src/psi4/scfgrad/
jk_grad.cc) -- TEST INFRASTRUCTURE: the checker of b200jk_grad_* (include/b200jk.h), never the product path.

It follows the reference literally, i.e. from the UNFITTED integrals (A|mn) and the full inverse metric
(FittingMetric::form_full_eig_inverse = Matrix::power(-1.0, condition), lib3index/fittingmetric.cc:432-439), not from
the fitted tensor the engine holds -- so agreement with the engine also checks the algebra that lets the engine avoid
recomputing integrals (d = J^-1/2 (B . Dt), (A|ij)_fitted = J^-1/2 C^T B C).

Plain numpy at the sizes the tests use (tens of functions); every function cites the statement it restates.
Pinned by tests/test_gradient.py: the total DF-RHF gradient of tests/fd-gradient (H2O / STO-3G, output.ref:330-335)
assembled from these intermediates reproduces the reference's analytic gradient."""
from __future__ import annotations

import numpy as np


def matrix_power(A: np.ndarray, alpha: float, cutoff: float) -> np.ndarray:
    """Matrix::power (libmints/matrix.cc:2370-2424)."""
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


def build_Amn_terms(Amn: np.ndarray, Dt: np.ndarray, Ca: np.ndarray, Cb: np.ndarray | None = None):
    """jk_grad.cc:294-475: c_A = (A|mn) Dt_mn (:431), (A|mi) = (A|mn) C_ni (:437), (A|ij) = C^T (A|mi) (:441-444).
    Cb None = restricted (the reference's Ca_ == Cb_)."""
    naux, nso, _ = Amn.shape
    c = Amn.reshape(naux, nso * nso) @ Dt.reshape(-1)
    out = []
    for C in ([Ca] if Cb is None else [Ca, Cb]):
        if C.shape[1] == 0:
            out.append(np.zeros((naux, 0, 0)))
            continue
        Ami = Amn @ C                                   # (naux, nso, na)
        out.append(np.einsum("mi,Amj->Aij", C, Ami))    # (naux, na, na)
    return c, out


def build_AB_inv_terms(metric: np.ndarray, condition: float, c: np.ndarray, Aij: list[np.ndarray]):
    """jk_grad.cc:634-721: d = J^-1 c (:659), (A|ij) <- J^-1 (B|ij) (:705)."""
    Jinv = matrix_power(metric, -1.0, condition)
    d = Jinv @ c
    fitted = [np.tensordot(Jinv, a, axes=([1], [0])) for a in Aij]
    return d, fitted


def build_UV_terms(Aij_fitted: list[np.ndarray]):
    """jk_grad.cc:722-812: V_AB = sum_spin sum_ij (A|ij)(B|ij), scaled by 2 when restricted."""
    naux = Aij_fitted[0].shape[0]
    V = np.zeros((naux, naux))
    for a in Aij_fitted:
        flat = a.reshape(naux, -1)
        V += flat @ flat.T
    if len(Aij_fitted) == 1:
        V *= 2.0
    return V


def Kmn_rows(Aij_fitted: list[np.ndarray], C: list[np.ndarray], a0: int, a1: int) -> np.ndarray:
    """jk_grad.cc:1010-1023: Kmn[A] = factor sum_spin C (A|ij) C^T for one block of auxiliary rows."""
    factor = 2.0 if len(Aij_fitted) == 1 else 1.0
    nso = C[0].shape[0]
    out = np.zeros((a1 - a0, nso, nso))
    for a, c in zip(Aij_fitted, C):
        if c.shape[1] == 0:
            continue
        out += factor * np.einsum("mi,Aij,nj->Amn", c, a[a0:a1], c, optimize=True)
    return out


def jk_gradient(d, V, Kmn, Dt, dAB, dAmn):
    """The dot products with the derivative integrals:
       metric part   MintsHelper::metric_grad (libmints/mintshelper.cc:2388-2490):  -0.5 (A|B)^x d_A d_B , -0.5 (A|B)^x V_AB
       3-index part  build_Amn_x_terms (jk_grad.cc:1075-1110):                      (A|mn)^x d_A Dt_mn , (A|mn)^x Kmn[A,mn]
    dAB[atom, xyz, A, B] and dAmn[atom, xyz, A, m, n] are total derivatives of the integrals with respect to the nuclear
    coordinate (all centres that sit on the atom move).  Returns (Coulomb, Exchange) gradient matrices (natom, 3); the
    caller scales Exchange by -alpha as scf_grad.cc:266 does."""
    J = -0.5 * np.einsum("axAB,A,B->ax", dAB, d, d) + np.einsum("axAmn,A,mn->ax", dAmn, d, Dt, optimize=True)
    K = -0.5 * np.einsum("axAB,AB->ax", dAB, V) + np.einsum("axAmn,Amn->ax", dAmn, Kmn, optimize=True)
    return J, K
