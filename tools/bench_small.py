"""Small-system latency (VERDICT r1 missing #7 / next #9): what a tiny JK handle costs end to end.

  cfg1    H2O / cc-pVDZ shape (24 bf, 116 aux, 5 occupied): create + set_layout + upload + first build + steady builds
  sad10   a SAD-style sequence (libscf_solver/sad.cc:706-763 builds one private MemDFJK per unique atom): ten handles
          of atom size (14 bf, 70 aux), each created, fed, used for a few builds and destroyed
  bz2     benzene dimer / aug-cc-pVDZ shape (384 bf, 1416 aux, 42 occupied): the launch-bound mid-size case

Prints one JSON line; wall-clock (host) times, since latency IS the metric here."""
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from psi4_b200 import DFHelper, Engine  # noqa: E402


def case(n, a, o, seed=1):
    rng = np.random.default_rng(seed)
    d = DFHelper(n, a)
    d.prepare_sparsity(keep=np.ones((n, n), dtype=bool))
    B = rng.standard_normal((a, n, n)) * 0.1
    P = d.pack(B + B.transpose(0, 2, 1))
    C = np.linalg.qr(rng.standard_normal((n, o)))[0]
    return d, P, C


def life_cycle(n, a, o, builds):
    d, P, C = case(n, a, o)
    D = C @ C.T
    t = {}
    t0 = time.perf_counter()
    e = Engine(1)
    t["create_ms"] = (time.perf_counter() - t0) * 1e3
    t0 = time.perf_counter()
    e.set_layout(n, a, d.small_skips_, d.big_skips_, d.schwarz_fun_index_)
    t["set_layout_ms"] = (time.perf_counter() - t0) * 1e3
    t0 = time.perf_counter()
    e.upload(0, P)
    t["upload_ms"] = (time.perf_counter() - t0) * 1e3
    t0 = time.perf_counter()
    e.compute([C], None, [D], reuse_outputs=True)
    t["first_build_ms"] = (time.perf_counter() - t0) * 1e3
    ts, dev, launches = [], [], 0
    for _ in range(builds):
        t0 = time.perf_counter()
        e.compute([C], None, [D], reuse_outputs=True)
        ts.append((time.perf_counter() - t0) * 1e3)
        st = e.stats()
        dev.append(st["ms_total"])
        launches = st["launches"]
    t["steady_build_ms"] = float(np.median(ts))
    t["steady_build_device_ms"] = float(np.median(dev))
    t["launches_per_build"] = int(launches)
    t0 = time.perf_counter()
    e.close()
    t["destroy_ms"] = (time.perf_counter() - t0) * 1e3
    return t


def main():
    life_cycle(24, 116, 5, 2)  # context creation, module load: not a per-handle cost
    out = {"cfg1_h2o_dz": life_cycle(24, 116, 5, 50), "bz2_adz": life_cycle(384, 1416, 42, 20)}
    t0 = time.perf_counter()
    per = [life_cycle(14, 70, 4, 6) for _ in range(10)]
    out["sad10"] = {"total_ms": (time.perf_counter() - t0) * 1e3,
                    "per_handle_ms": float(np.median([sum(v for k, v in p.items() if k.endswith("_ms") and "steady" not in k)
                                                      + 6 * p["steady_build_ms"] for p in per])),
                    "median": {k: float(np.median([p[k] for p in per])) for k in per[0]}}
    print(json.dumps(out))


if __name__ == "__main__":
    main()
