"""Element-wise golden vectors made by the reference's own object code (tools/make_golden_jk.py ->
tests/golden/reference_jk_vectors.npz; /root/reference is not needed to READ them): J, K and wK for four (C_left,
C_right) pairs with nocc 6 / 0 / 3 / 17 on a 40 %-screened mask, symmetric and general.  The CPU restatement must
reproduce them bit for bit (one thread), the CUDA engine through the C ABI to 1e-10 (north_star)."""
import os

import numpy as np
import pytest

from psi4_b200 import DFHelper

G = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "reference_jk_vectors.npz"))
NMAT = len(G["noccs"])


def inputs(tag):
    keep = G["keep"]
    n, a = keep.shape[0], int(G["naux"])
    d = DFHelper(n, a)
    d.prepare_sparsity(keep=keep)
    Cl = [G[f"Cl{i}"] for i in range(NMAT)]
    Cr = None if tag == "sym" else [G[f"Cr{i}"] for i in range(NMAT)]
    assert [c.shape[1] for c in Cl] == list(G["noccs"])
    want = [[G[f"{p}_{tag}{i}"] for i in range(NMAT)] for p in ("J", "K", "wK")]
    return d, Cl, Cr, want


@pytest.mark.parametrize("tag", ["sym", "gen"])
def test_restatement_reproduces_reference_vectors_bit_for_bit(oracle, tag):
    d, Cl, Cr, want = inputs(tag)
    sp = oracle.Sparsity(d.keep_.astype(np.uint8), d.naux_)
    J, K, wK, _ = oracle.build_JK(sp, G["Ppq"], Cl, Cr, do_wK=True, m1Ppq=G["m1Ppq"], wPpq=G["wPpq"], nthreads=1)
    for got, ref in zip((J, K, wK), want):
        for g, r in zip(got, ref):
            assert np.array_equal(g, r)
    assert not want[1][1].any() and np.abs(want[1][0]).max() > 1.0  # nocc 0 -> zero K; the others are not trivial


@pytest.mark.gpu
@pytest.mark.parametrize("tag", ["sym", "gen"])
def test_gpu_engine_reproduces_reference_vectors(tag):
    from psi4_b200 import Engine

    d, Cl, Cr, want = inputs(tag)
    e = Engine(1)
    e.set_layout(d.nbf_, d.naux_, d.small_skips_, d.big_skips_, d.schwarz_fun_index_)
    e.upload(0, G["Ppq"])
    e.upload(1, G["m1Ppq"])
    e.upload(2, G["wPpq"])
    D = [x @ (x if Cr is None else y).T for x, y in zip(Cl, Cl if Cr is None else Cr)]
    J, K, wK = e.compute(Cl, Cr, D, do_wK=True)
    for name, got, ref in zip("J K wK".split(), (J, K, wK), want):
        for i, (g, r) in enumerate(zip(got, ref)):
            assert np.abs(g - r).max() < 1e-10, f"{name}[{i}] {tag}"
    e.close()
