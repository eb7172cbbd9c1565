"""Host-side mirror of the part of psi4's DFHelper that the MEM_DF J/K path needs:
the Schwarz pair mask, the packed-index tables and the "pQq" packing of the fitted
three-index tensor.  Pure numpy host logic (cheap, once per SCF); the arithmetic of the
J/K build itself lives in libb200jk.so.

Reference: psi4/src/psi4/lib3index/dfhelper.cc
  prepare_sparsity :299-420   layout :1274-1276 / :1666-1677   get_core_size :216-236
"""
from __future__ import annotations

import numpy as np


class DFHelper:
    """Sparsity tables + packing.  Attribute names follow dfhelper.h:440-452."""

    def __init__(self, nbf: int, naux: int):
        self.nbf_ = int(nbf)
        self.naux_ = int(naux)
        self.cutoff_ = 1e-12  # dfhelper.h: schwarz cutoff default; JK passes INTS_TOLERANCE (jk.cc:58-68)
        self.sparsity_prepared_ = False
        self.do_wK_ = False
        self.Qshell_max_ = 0  # largest auxiliary shell, prepare_blocking dfhelper.cc:84-103

    # ---- knobs (dfhelper.h:100-175) ----
    def set_schwarz_cutoff(self, cutoff: float):
        self.cutoff_ = float(cutoff)

    def get_schwarz_cutoff(self) -> float:
        return self.cutoff_

    def set_do_wK(self, do_wK: bool):
        self.do_wK_ = bool(do_wK)

    def prepare_blocking(self, pshell_nfunction, Qshell_nfunction):
        """dfhelper.cc:84-103: function offsets of the primary / auxiliary shells (pshell_aggs_, Qshell_aggs_) and the
        largest auxiliary shell (Qshell_max_), from the per-shell function counts of the two basis sets."""
        p = np.asarray(pshell_nfunction, dtype=np.int64)
        q = np.asarray(Qshell_nfunction, dtype=np.int64)
        if p.sum() != self.nbf_ or q.sum() != self.naux_:
            raise ValueError("DFHelper: shell sizes do not add up to nbf / naux")
        self.pshells_, self.Qshells_ = len(p), len(q)
        self.pshell_aggs_ = np.concatenate([[0], np.cumsum(p)]).astype(np.uintp)
        self.Qshell_aggs_ = np.concatenate([[0], np.cumsum(q)]).astype(np.uintp)
        self.Qshell_max_ = int(q.max())

    def pshell_blocks_for_AO_build(self, mem: int, symm: int = 1, hold_met: bool = False):
        """dfhelper.cc:700-762 for the in-core (symm) case: consecutive primary shells are grouped while
        block + full tensor + metric (or a second block buffer) fits in `mem` doubles.  Returns (steps, largest buffer,
        largest block) with steps = [(first shell, last shell), ...] -- the p-blocks prepare_AO_core feeds to
        compute_sparse_pQq_blocking_p_symm / contract_metric_AO_core_symm (:566-585), i.e. to b200jk_fit_rows."""
        if not symm:
            raise NotImplementedError("the on-disk blocking is out of scope (no DISK_DF behind the engine)")
        full = int(self.big_skips_[self.nbf_])
        steps, largest, block_size = [], 0, 0
        i, count, total, tmpbs = 0, 0, 0, 0
        scale = 3 if self.do_wK_ else 1
        while i < self.pshells_:
            count += 1
            begin, end = int(self.pshell_aggs_[i]), int(self.pshell_aggs_[i + 1]) - 1
            tmpbs += end - begin + 1
            current = scale * int(self.symm_big_skips_[end + 1] - self.symm_big_skips_[begin])
            total += current
            constraint = total + full + (self.naux_ * self.naux_ if hold_met else total)
            last = i == self.pshells_ - 1
            if constraint > mem or last:
                if count == 1 and not last:
                    raise MemoryError("DFHelper: not enough memory for (p shell) AO blocking! required memory: "
                                      f"{constraint * 8 / 1024.0 ** 3} [GiB].")
                if constraint > mem:
                    total -= current
                    tmpbs -= end - begin + 1
                    steps.append((i - count + 1, i - 1))
                    i -= 1
                elif last:
                    steps.append((i - count + 1, i))
                if largest < total:
                    largest, block_size = total, tmpbs
                count = total = tmpbs = 0
            i += 1
        return steps, largest, block_size

    def set_Qshell_max(self, qshell_max: int):
        """Qshell_max_ of prepare_blocking (dfhelper.cc:84-103): functions in the largest auxiliary shell."""
        self.Qshell_max_ = int(qshell_max)

    # ---- dfhelper.cc:371-416 ----
    def prepare_sparsity(self, fun_max_vals: np.ndarray | None = None, keep: np.ndarray | None = None):
        """Build the mask from per-pair Schwarz maxima |(mn|mn)| (tolerance = cutoff^2 / max, :371)
        or take a ready boolean mask, then the index tables."""
        n = self.nbf_
        if keep is None:
            f = np.asarray(fun_max_vals, dtype=np.float64).reshape(n, n)
            max_val = f.max()
            tolerance = self.cutoff_ * self.cutoff_ / max_val
            keep = f >= tolerance
        keep = np.asarray(keep, dtype=bool).reshape(n, n)
        if not np.all(np.diag(keep)):
            raise ValueError("DFHelper: a diagonal pair is screened out (dfhelper.cc:3213 assumes it exists)")
        self.keep_ = keep
        if getattr(self, "pshell_aggs_", None) is not None:
            # schwarz_shell_mask_ (:374): a shell pair is significant iff any of its function pairs is
            a = self.pshell_aggs_.astype(np.int64)
            self.schwarz_shell_mask_ = np.add.reduceat(np.add.reduceat(keep.astype(np.int64), a[:-1], axis=0), a[:-1], axis=1) > 0
        count = np.cumsum(keep, axis=1)
        self.schwarz_fun_index_ = np.where(keep, count, 0).astype(np.uintp)       # :377-387
        sp = keep.sum(axis=1).astype(np.uintp)
        self.small_skips_ = np.concatenate([sp, [sp.sum()]]).astype(np.uintp)      # :386,:398
        self.big_skips_ = np.concatenate([[0], np.cumsum(sp * np.uintp(self.naux_))]).astype(np.uintp)  # :390-397
        lower = np.tril(keep, -1).sum(axis=1).astype(np.uintp)
        self.symm_ignored_columns_ = lower                                         # :401-411
        self.symm_small_skips_ = (sp - lower).astype(np.uintp)
        self.symm_big_skips_ = np.concatenate(
            [[0], np.cumsum(self.symm_small_skips_ * np.uintp(self.naux_))]).astype(np.uintp)  # :413-416
        self.sparsity_prepared_ = True

    def ao_sparsity(self) -> float:
        """dfhelper.h:101  fraction of screened pairs."""
        return 1.0 - float(self.small_skips_[self.nbf_]) / float(self.nbf_ * self.nbf_)

    def get_core_size(self, nthreads: int = 1, qshell_max: int | None = None) -> int:
        """dfhelper.cc:216-236 required_core_size_ in doubles (host model; Qshell_max needs the aux shells)."""
        if qshell_max is None:
            qshell_max = self.Qshell_max_
        big = int(self.big_skips_[self.nbf_])
        req = 3 * big if self.do_wK_ else big
        req += self.naux_ * self.naux_
        req += nthreads * self.nbf_ * self.nbf_
        req += 3 * self.nbf_ * self.nbf_ * qshell_max
        return req

    # ---- packing ----
    def kept_columns(self, m: int) -> np.ndarray:
        return np.nonzero(self.keep_[m])[0]

    def pack(self, dense_Qmn: np.ndarray) -> np.ndarray:
        """(naux, nbf, nbf) -> pQq packed vector: B(Q,m,n) at big_skips[m] + Q*sp(m) + f(m,n)-1."""
        B = np.asarray(dense_Qmn, dtype=np.float64)
        assert B.shape == (self.naux_, self.nbf_, self.nbf_)
        out = np.empty(int(self.big_skips_[self.nbf_]), dtype=np.float64)
        for m in range(self.nbf_):
            cols = self.kept_columns(m)
            a, b = int(self.big_skips_[m]), int(self.big_skips_[m + 1])
            out[a:b] = B[:, m, cols].ravel()
        return out

    def pack_symm(self, dense_Amn: np.ndarray, m0: int = 0, m1: int | None = None) -> np.ndarray:
        """(naux, nbf, nbf) -> the symmetric-packed buffer of compute_sparse_pQq_blocking_p_symm (dfhelper.cc:1338-1340)
        for rows [m0, m1): block m is [naux][mi(m)] over the kept partners n >= m."""
        m1 = self.nbf_ if m1 is None else m1
        B = np.asarray(dense_Amn, dtype=np.float64)
        base = int(self.symm_big_skips_[m0])
        out = np.empty(int(self.symm_big_skips_[m1]) - base, dtype=np.float64)
        for m in range(m0, m1):
            cols = self.kept_columns(m)
            cols = cols[cols >= m]
            a, b = int(self.symm_big_skips_[m]) - base, int(self.symm_big_skips_[m + 1]) - base
            out[a:b] = B[:, m, cols].ravel()
        return out

    def unpack(self, packed: np.ndarray) -> np.ndarray:
        out = np.zeros((self.naux_, self.nbf_, self.nbf_))
        for m in range(self.nbf_):
            cols = self.kept_columns(m)
            a, b = int(self.big_skips_[m]), int(self.big_skips_[m + 1])
            out[:, m, cols] = packed[a:b].reshape(self.naux_, len(cols))
        return out
