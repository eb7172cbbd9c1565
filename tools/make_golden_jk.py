"""Generate tests/golden/reference_jk_vectors.npz: element-wise J / K / wK produced by the REFERENCE'S OWN object code
(oracle/_ref/libref_dfjk.so, built by oracle/ref_build.py from /root/reference/psi4/src/psi4/lib3index/dfhelper.cc) on
small seeded inputs, inputs included.  /root/reference does not exist on the GPU box, so the `-m gpu` parity tests read
these vectors instead (tests/test_golden_vectors.py).  One OpenMP thread: the reference's own J is then reproducible."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "oracle"))
import dfjk_oracle as oracle  # noqa: E402

from psi4_b200 import DFHelper  # noqa: E402

OUT = os.path.join(ROOT, "tests", "golden", "reference_jk_vectors.npz")


def main():
    if oracle.ref_lib() is None:
        raise SystemExit("needs /root/reference (oracle/ref_build.py)")
    rng = np.random.default_rng(20261017)
    n, a = 44, 37
    r = rng.random((n, n))
    keep = (r + r.T) * 0.5 < 0.55
    np.fill_diagonal(keep, True)
    d = DFHelper(n, a)
    d.prepare_sparsity(keep=keep)
    sym = lambda: (lambda b: b + b.transpose(0, 2, 1))(rng.standard_normal((a, n, n)) * 0.2)  # noqa: E731
    P, P1, PW = d.pack(sym()), d.pack(sym()), d.pack(sym())
    noccs = [6, 0, 3, 17]
    Cl = [rng.standard_normal((n, o)) for o in noccs]
    Cr = [rng.standard_normal((n, o)) for o in noccs]
    sp = oracle.Sparsity(keep.astype(np.uint8), a)
    out = {"keep": keep, "naux": np.array(a), "Ppq": P, "m1Ppq": P1, "wPpq": PW, "noccs": np.array(noccs)}
    for i, (x, y) in enumerate(zip(Cl, Cr)):
        out[f"Cl{i}"], out[f"Cr{i}"] = x, y
    for tag, right in (("sym", None), ("gen", Cr)):
        J, K, wK, _ = oracle.build_JK(sp, P, Cl, right, do_wK=True, m1Ppq=P1, wPpq=PW, nthreads=1, impl="ref")
        for i in range(len(noccs)):
            out[f"J_{tag}{i}"], out[f"K_{tag}{i}"], out[f"wK_{tag}{i}"] = J[i], K[i], wK[i]
    np.savez_compressed(OUT, **out)
    print(OUT, os.path.getsize(OUT), "bytes; sparsity", d.ao_sparsity())


if __name__ == "__main__":
    main()
