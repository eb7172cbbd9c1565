"""Time b200jk_upload of a fitted packed tensor from pageable and from registered host memory (GB/s over PCIe)."""
import argparse
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np  # noqa: E402

from psi4_b200 import DFHelper, Engine  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--nbf", type=int, default=400)
ap.add_argument("--naux", type=int, default=4740)
ap.add_argument("--gpus", type=int, default=1)
args = ap.parse_args()
n, a = args.nbf, args.naux
d = DFHelper(n, a)
d.prepare_sparsity(keep=np.ones((n, n), bool))
P = np.ones(int(d.big_skips_[n]))
out = {"nbf": n, "naux": a, "gpus": args.gpus, "tensor_gb": P.nbytes / 1e9}
for mode in ("pageable", "pageable_again", "registered"):
    e = Engine(args.gpus)
    e.set_layout(n, a, d.small_skips_, d.big_skips_, d.schwarz_fun_index_)
    if mode == "registered":
        t0 = time.perf_counter()
        e.register_host(P)
        out["register_s"] = time.perf_counter() - t0
    t0 = time.perf_counter()
    e.upload(0, P)
    dt = time.perf_counter() - t0
    out[mode + "_s"] = dt
    out[mode + "_GBs"] = P.nbytes / dt / 1e9
    got = e.download_rows(0, n // 2, 0, a)
    assert np.array_equal(got.ravel(), np.ones(got.size))
    if mode == "registered":
        e.unregister_host(P)
    e.close()
print(json.dumps(out))
