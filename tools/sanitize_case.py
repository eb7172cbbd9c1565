"""One small JK build (screened mask, ragged nocc, fused J, symmetric and general, wK) through the C ABI -- the workload
tools/sanitize.sh runs under compute-sanitizer (memcheck / racecheck / initcheck).  Sizes are tiny on purpose: racecheck
slows kernels down by two to three orders of magnitude."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from psi4_b200 import DFHelper, Engine  # noqa: E402

rng = np.random.default_rng(11)
n, a = int(os.environ.get("SAN_NBF", "150")), int(os.environ.get("SAN_NAUX", "140"))
r = rng.random((n, n))
keep = (r + r.T) < 1.3
np.fill_diagonal(keep, True)
d = DFHelper(n, a)
d.prepare_sparsity(keep=keep)


def packed():
    B = rng.standard_normal((a, n, n)) * 0.1
    return d.pack(B + B.transpose(0, 2, 1))


e = Engine(1)
e.set_layout(n, a, d.small_skips_, d.big_skips_, d.schwarz_fun_index_)
for w in range(3):
    e.upload(w, packed())
Cl = [rng.standard_normal((n, 21)), rng.standard_normal((n, 8))]
Cr = [rng.standard_normal((n, 21)), rng.standard_normal((n, 8))]
J, K, wK = e.compute(Cl, None, [c @ c.T for c in Cl], do_wK=True)             # symmetric, fused first J sweep
J2, K2, _ = e.compute(Cl, Cr, [x @ y.T for x, y in zip(Cl, Cr)])              # general
J3, K3, _ = e.compute(Cl, None, [c @ c.T for c in Cl], do_K=False)            # J only (separate first sweep)
print("sanitize_case ok", float(np.abs(K[0]).max()), float(np.abs(J2[1]).max()), float(np.abs(J3[0]).max()))
e.close()
