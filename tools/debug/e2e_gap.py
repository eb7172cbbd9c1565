"""Per-build device stats and wall time of the host-operand entry point for one workload (debugging aid)."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import numpy as np
from psi4_b200 import DFHelper, Engine, workloads
import bench

class A: pass
args = A(); args.nonsymmetric = False; args.response = 0
name = sys.argv[1] if len(sys.argv) > 1 else "c20h42_tz"
cfg, keep, amp, Cl, Crl, D = bench.make_inputs(args, name)
nbf, naux = cfg["nbf"], cfg["naux"]
d = DFHelper(nbf, naux); d.prepare_sparsity(keep=keep)
e = Engine(1); e.set_layout(nbf, naux, d.small_skips_, d.big_skips_, d.schwarz_fun_index_)
e.set_half(sys.argv[2] if len(sys.argv) > 2 else "auto"); e.set_kgemm(sys.argv[3] if len(sys.argv) > 3 else "auto")
e.fill_synthetic(0, workloads.SEED, amp)
n2b = nbf * nbf * 8
dC = [e.dev_put(x) for x in Cl]; dD = [e.dev_put(x) for x in D]
dJ = [e.dev_alloc(n2b) for _ in Cl]; dK = [e.dev_alloc(n2b) for _ in Cl]
for it in range(4):
    t0 = time.perf_counter()
    e.compute_device(dC, None, [cfg["nocc"]] * len(Cl), dD, dJ, dK, None)
    w = (time.perf_counter() - t0) * 1e3
    st = e.stats()
    print(f"build dev{it}: wall {w:8.2f} ms  total {st['ms_total']:7.2f} half {st['ms_half']:7.2f} kgemm {st['ms_kgemm']:7.2f} launches {st['launches']}")
for x in D: e.register_host(x)
for it in range(6):
    t0 = time.perf_counter()
    J, K, _ = e.compute(Cl, Crl, D, reuse_outputs=True)
    w = (time.perf_counter() - t0) * 1e3
    st = e.stats()
    print(f"build {it}: wall {w:8.2f} ms  total {st['ms_total']:7.2f} half {st['ms_half']:7.2f} kgemm {st['ms_kgemm']:7.2f} j {st['ms_j']:6.2f} "
          f"h2d {st['ms_h2d']:5.2f} d2h {st['ms_d2h']:5.2f} launches {st['launches']} half_kind {st['half_kind']} kgemm_kind {st['kgemm_kind']} work_gb {st['hbm_work_bytes']/1e9:.1f}")
e.close()
