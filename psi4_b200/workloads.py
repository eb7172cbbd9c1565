"""Synthetic stand-ins for the BASELINE.json configurations (no psi4 / no integrals needed):
sizes from SURVEY.md section 8, a symmetric pair mask, an amplitude envelope for the synthetic
tensor (b200jk_fill_synthetic / oracle_synth_fill) and seeded orthonormal orbitals."""
from __future__ import annotations

import numpy as np

# name -> (nbf, naux, nocc, nmat, sparsity model)
CONFIGS = {
    "h2o_dz": dict(nbf=24, naux=116, nocc=5, nmat=1, band=None),
    "bz2_adz": dict(nbf=384, naux=1416, nocc=42, nmat=1, band=None),
    "c20h42_tz": dict(nbf=1188, naux=2840, nocc=81, nmat=1, band=0.55),
    "c60_tz": dict(nbf=1800, naux=4740, nocc=180, nmat=1, band=None),
    "h2o40_tz": dict(nbf=2320, naux=5560, nocc=200, nmat=2, band=0.6),
    # one eighth of C60's auxiliary index (what each GPU holds at 8 GPUs): short enough for ncu --set full
    "c60_tz_q8": dict(nbf=1800, naux=592, nocc=180, nmat=1, band=None),
}
SEED = 20251017


def pair_mask(nbf: int, band: float | None, block: int = 30) -> np.ndarray:
    """Symmetric boolean mask, diagonal kept.  band=None -> all pairs kept (the C60/cc-pVTZ case
    at cutoff 1e-12, and SCREENING=NONE).  Otherwise a block-banded mask standing in for a
    chain-like molecule: function blocks (atoms) keep partners within band*nbf functions."""
    if band is None:
        return np.ones((nbf, nbf), dtype=bool)
    blk = np.arange(nbf) // block
    reach = max(1, int(round(band * nbf / block / 2)))
    keep = np.abs(blk[:, None] - blk[None, :]) <= reach
    np.fill_diagonal(keep, True)
    return keep


def amplitude(nbf: int, scale: float = 2.0e-2) -> np.ndarray:
    """Symmetric envelope amp[m,n] = scale * exp(-3|m-n|/nbf): O(1e-2) magnitudes like a fitted tensor."""
    i = np.arange(nbf)
    return scale * np.exp(-3.0 * np.abs(i[:, None] - i[None, :]) / nbf)


def orbitals(nbf: int, nocc: int, seed: int = SEED) -> np.ndarray:
    """First nocc columns of the Q factor of a seeded normal matrix (orthonormal)."""
    rng = np.random.default_rng(seed)
    q, _ = np.linalg.qr(rng.standard_normal((nbf, max(nocc, 1))))
    return np.ascontiguousarray(q[:, :nocc])
