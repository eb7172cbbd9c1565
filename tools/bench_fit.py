"""Time the on-device fitting step (b200jk_fit_rows: transpose + metric GEMM + mirror) on a synthetic unfitted tensor.
Reports the DMMA rate of the contraction (device events) and the wall time incl. the pageable H2D of the raw blocks,
next to the CPU restatement of contract_metric_AO_core_symm on a few row-blocks (OpenMP + OpenBLAS)."""
import argparse
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "oracle"))
import numpy as np  # noqa: E402

from psi4_b200 import DFHelper, Engine  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--nbf", type=int, default=600)
ap.add_argument("--naux", type=int, default=4740)
ap.add_argument("--block", type=int, default=40)
ap.add_argument("--pageable", action="store_true", help="do not page-lock the block buffer")
ap.add_argument("--gpus", type=int, default=1, help="Q shards driven by this one process")
args = ap.parse_args()
n, a = args.nbf, args.naux
d = DFHelper(n, a)
d.prepare_sparsity(keep=np.ones((n, n), bool))
rng = np.random.default_rng(0)
g = rng.standard_normal((a, a))
met = g @ g.T / a + np.eye(a)
e = Engine(args.gpus)
e.set_layout(n, a, d.small_skips_, d.big_skips_, d.schwarz_fun_index_)
e.set_metric(met)
wall = 0.0
calls = []
first_block = None
# psi4 reuses ONE block buffer (Qpq / Mp, dfhelper.cc:553) for every p-block: page-lock it once (b200jk_register_host)
# so the raw integrals go up by DMA at PCIe speed; --pageable shows the unregistered path
sizes = [int(d.symm_big_skips_[min(n, m0 + args.block)] - d.symm_big_skips_[m0]) for m0 in range(0, n, args.block)]
buf = np.zeros(max(sizes))
if not args.pageable:
    e.register_host(buf)
for m0 in range(0, n, args.block):
    m1 = min(n, m0 + args.block)
    size = int(d.symm_big_skips_[m1] - d.symm_big_skips_[m0])
    blk = buf[:size]
    blk[:] = rng.standard_normal(size)
    if first_block is None:
        first_block = (m0, m1, blk.copy())
    t0 = time.perf_counter()
    e.fit_rows(0, m0, m1, blk)
    calls.append(time.perf_counter() - t0)
    wall += calls[-1]
st = e.fit_stats()
out = {"nbf": n, "naux": a, "pair_columns": int(d.symm_big_skips_[n] // a), "raw_gb": float(d.symm_big_skips_[n]) * 8 / 1e9,
       "gpu_gemm_ms": st["ms_gemm"], "gpu_gemm_tflops": st["tflops"], "gpu_wall_s_incl_h2d": wall, "gpus": args.gpus,
       # fit_rows returns once the caller's block has been read; only the last call waits for the GPU to finish
       "host_blocked_s_before_last_block": sum(calls[:-1]), "last_call_s": calls[-1],
       "block_buffer": "pageable" if args.pageable else "registered (page-locked once)"}
try:
    import dfjk_oracle as oracle

    sp = oracle.Sparsity(np.ones((n, n), np.uint8), a)
    m0, m1, blk = first_block
    P = np.zeros(sp.packed_size)
    t0 = time.perf_counter()
    oracle.contract_metric_AO_core_symm(sp, blk, met, P, begin=m0, end=m1 - 1)
    dt = time.perf_counter() - t0
    fl = 2.0 * a * a * float(d.symm_big_skips_[m1] - d.symm_big_skips_[m0]) / a
    out.update({"cpu_cores": os.cpu_count(), "cpu_block_s": dt, "cpu_gflops": fl / dt / 1e9,
                "cpu_extrapolated_s": dt * float(d.symm_big_skips_[n]) / float(d.symm_big_skips_[m1] - d.symm_big_skips_[m0])})
    # parity of that block
    got = e.download_rows(0, m0, 0, a).ravel()
    want = P[int(d.big_skips_[m0]):int(d.big_skips_[m0 + 1])]
    out["block_max_abs_diff"] = float(np.abs(got - want).max())
except Exception as ex:  # oracle is optional here
    out["cpu_error"] = repr(ex)
print(json.dumps(out))
