"""Stand-ins for the BASELINE.json configurations (no psi4 needed): sizes from SURVEY.md section 8, the pair mask,
an amplitude envelope for the synthetic tensor (b200jk_fill_synthetic / oracle_synth_fill) and seeded orthonormal
orbitals.

Pair masks: for the three large configurations the mask is REAL -- DFHelper::prepare_sparsity at cutoff 1e-12 on
Schwarz integrals (mn|mn) computed by the host integral front end for the actual geometry and cc-pVTZ basis
(tools/real_masks.py; stored bit-packed in share/masks/).  Measured mask sparsities: C60 28.50 %, n-C20H42 60.82 %,
(H2O)40 68.29 %.  (The tensor VALUES stay synthetic: the full (A|mn) of C60 is 88 GB.)  The small configurations are
unscreened at 1e-12, as the reference prints for water and the benzene dimer.
"""
from __future__ import annotations

import json
import os

import numpy as np

_MASKS = os.path.join(os.path.dirname(os.path.abspath(__file__)), "share", "masks")

# name -> sizes; mask = file stem under share/masks (real Schwarz mask) or None (all pairs kept)
CONFIGS = {
    "h2o_dz": dict(nbf=24, naux=116, nocc=5, nmat=1, mask=None),
    "bz2_adz": dict(nbf=384, naux=1416, nocc=42, nmat=1, mask=None),
    "c20h42_tz": dict(nbf=1188, naux=2840, nocc=81, nmat=1, mask="c20h42_cc-pvtz_1e-12"),
    "c60_tz": dict(nbf=1800, naux=4740, nocc=180, nmat=1, mask="c60_cc-pvtz_1e-12"),
    "h2o40_tz": dict(nbf=2320, naux=5560, nocc=200, nmat=2, mask="h2o40_cc-pvtz_1e-12"),
    # one eighth of C60's auxiliary index (what each GPU holds at 8 GPUs): short enough for ncu --set full
    "c60_tz_q8": dict(nbf=1800, naux=592, nocc=180, nmat=1, mask="c60_cc-pvtz_1e-12"),
    # one eighth of (H2O)40's auxiliary index: the per-GPU shard of the 8-GPU UHF run (two densities, gather path)
    "h2o40_tz_q8": dict(nbf=2320, naux=695, nocc=200, nmat=2, mask="h2o40_cc-pvtz_1e-12"),
    # SCREENING=NONE analogue (jk.cc:60-61): every pair kept, the dense upper bound of SURVEY.md section 8
    "c60_tz_dense": dict(nbf=1800, naux=4740, nocc=180, nmat=1, mask=None),
}
SEED = 20251017


def pair_mask(nbf: int, mask: str | None) -> np.ndarray:
    """Symmetric boolean pair mask, diagonal kept."""
    if mask is None:
        return np.ones((nbf, nbf), dtype=bool)
    path = os.path.join(_MASKS, mask + ".npy")
    if not os.path.exists(path):
        raise FileNotFoundError(f"{path} missing: regenerate with tools/real_masks.py")
    keep = np.unpackbits(np.load(path))[: nbf * nbf].reshape(nbf, nbf).astype(bool)
    assert np.array_equal(keep, keep.T) and keep.diagonal().all()
    return keep


def mask_info(mask: str | None) -> str:
    if mask is None:
        return "all pairs kept"
    try:
        d = json.load(open(os.path.join(_MASKS, mask + ".json")))
        return (f"real Schwarz mask ({d['system']} / {d['basis']}, cutoff 1e-12, host integral front end): "
                f"{d['mask_sparsity_percent']:.2f} % sparse")
    except Exception:
        return f"mask {mask}"


def banded_mask(nbf: int, band: float, block: int = 30) -> np.ndarray:
    """Synthetic block-banded mask (tests): function blocks keep partners within band*nbf functions."""
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
