"""Property tests (hypothesis) of the host mirror that feeds b200jk_set_layout / b200jk_upload / b200jk_fit_rows:
whatever symmetric mask with a kept diagonal comes in, the tables obey the invariants the reference's loops rely on
(dfhelper.cc:377-416, :1274-1276, :1666-1677) and packing round-trips."""
import numpy as np
from hypothesis import given, settings
from hypothesis import strategies as st

from psi4_b200 import DFHelper


@st.composite
def masks(draw):
    n = draw(st.integers(1, 14))
    bits = draw(st.lists(st.booleans(), min_size=n * n, max_size=n * n))
    keep = np.array(bits, dtype=bool).reshape(n, n)
    keep = keep | keep.T
    np.fill_diagonal(keep, True)
    return keep, draw(st.integers(1, 6))


@settings(max_examples=60, deadline=None)
@given(masks())
def test_table_invariants(case):
    keep, naux = case
    n = keep.shape[0]
    d = DFHelper(n, naux)
    d.prepare_sparsity(keep=keep)
    f = d.schwarz_fun_index_
    for m in range(n):
        ranks = f[m][keep[m]]
        assert np.array_equal(ranks, np.arange(1, keep[m].sum() + 1))  # 1-based rank among kept partners (:377-387)
        assert not f[m][~keep[m]].any()
        assert d.small_skips_[m] == keep[m].sum()
        assert d.big_skips_[m + 1] - d.big_skips_[m] == naux * keep[m].sum()  # :390-397
        assert d.symm_ignored_columns_[m] == keep[m, :m].sum()  # kept partners below the diagonal (:401-411)
        assert d.symm_small_skips_[m] + d.symm_ignored_columns_[m] == d.small_skips_[m]
        assert d.symm_big_skips_[m + 1] - d.symm_big_skips_[m] == naux * d.symm_small_skips_[m]
    assert d.small_skips_[n] == keep.sum()
    assert 0.0 <= d.ao_sparsity() < 1.0


@settings(max_examples=40, deadline=None)
@given(masks(), st.integers(0, 2 ** 31 - 1))
def test_pack_roundtrip_and_symmetric_block_feed(case, seed):
    keep, naux = case
    n = keep.shape[0]
    d = DFHelper(n, naux)
    d.prepare_sparsity(keep=keep)
    rng = np.random.default_rng(seed)
    B = rng.standard_normal((naux, n, n))
    B = B + B.transpose(0, 2, 1)
    P = d.pack(B)
    assert P.size == d.big_skips_[n]
    assert np.array_equal(d.unpack(P), B * keep[None])  # element (Q,m,n) at big_skips[m] + Q*sp(m) + f(m,n) - 1
    for m in range(n):
        for q in (0, naux - 1):
            for k in np.nonzero(keep[m])[0][:3]:
                assert P[int(d.big_skips_[m]) + q * int(d.small_skips_[m]) + int(d.schwarz_fun_index_[m, k]) - 1] == B[q, m, k]
    # the symmetric-packed feed of b200jk_fit_rows: blocks concatenate to the whole, each block is [naux][mi(m)] over n >= m
    whole = d.pack_symm(B)
    cut = n // 2
    assert np.array_equal(whole, np.concatenate([d.pack_symm(B, 0, cut), d.pack_symm(B, cut, n)]))
    for m in range(n):
        blk = whole[int(d.symm_big_skips_[m]):int(d.symm_big_skips_[m + 1])].reshape(naux, -1)
        cols = [k for k in np.nonzero(keep[m])[0] if k >= m]
        assert np.array_equal(blk, B[:, m, cols])
