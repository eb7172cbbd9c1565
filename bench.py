#!/usr/bin/env python
"""bench.py -- DF-JK ms per SCF iteration (one MemDFJK::compute_JK build) on the BASELINE.json
workload: C60 / cc-pVTZ DF-RHF (nbf 1800, naux 4740, nocc 180), synthetic tensor of that shape.

  python bench.py [--gpus N] [--steps K] [--warmup W] [--workload c60_tz] [--impl reference]

One process per GPU (torchrun for N>1): the auxiliary index Q is sharded over the ranks and the
partial J/K are summed by one NCCL all-reduce inside the engine ("strong" scaling: total work fixed).

  value        ms per build with C/D/J/K resident in HBM (b200jk_compute_device), device time
               from CUDA events on the engine's stream, max over ranks
  e2e          ms per build through the host-pointer C ABI call b200jk_compute (what psi4's
               MemDFJK::compute_JK would call): pinned staging + H2D of C and D + D2H of J and K inside
  roofline     dominant kernel = K3 half-transform (DMMA); denominators: FP64 DMMA ceiling measured
               live by a register-resident m8n8k4 loop (MEASURED_PEAKS.json has no FP64 figure)
  cpu_baseline the oracle restatement of the reference's OpenMP+BLAS loops on this box's host cores,
               on a Q-slice of the same workload, extrapolated linearly in naux (every hot loop is
               linear in the Q extent: dfhelper.cc:3193, :3208, :2183, :3374)
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

import numpy as np  # noqa: E402

METRIC = "DF-JK ms/SCF-iter at C60/cc-pVTZ"


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--workload", default="c60_tz")
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--cpu-slice", type=int, default=0, help="Q rows in the CPU baseline sample (0 = auto)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--skip-probes", action="store_true", help="no FP64 ceiling probes (ncu launch lists); roofline.peak = last recorded")
    ap.add_argument("--nonsymmetric", action="store_true", help="C_right != C_left (general path)")
    ap.add_argument("--response", type=int, default=0, metavar="R",
                    help="R right-hand sides sharing ONE C_left with R different C_right, the shape of the reference's "
                         "response builds (twoel_Hx, libscf_solver/rhf.cc:466-484); implies --nonsymmetric")
    ap.add_argument("--single-process", action="store_true",
                    help="with --gpus N and no torchrun: one process drives N GPUs (psi4's deployment mode)")
    return ap.parse_args()


# --------------------------------------------------------------------------------------------
# clocks sampler (B200_PROFILING.md "clocks DURING the timed region")
# --------------------------------------------------------------------------------------------
class Clocks:
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, index=0):
        self.rows, self.stop_flag, self.index = [], False, index
        self.t = threading.Thread(target=self.run, daemon=True)

    def run(self):
        while not self.stop_flag:
            try:
                out = subprocess.run(["nvidia-smi", "-i", str(self.index), f"--query-gpu={self.Q}",
                                      "--format=csv,noheader,nounits"], capture_output=True, text=True, timeout=5).stdout
                parts = [x.strip() for x in out.strip().split(",")]
                if len(parts) >= 7:
                    self.rows.append(parts)
            except Exception:
                pass
            time.sleep(0.1)

    def start(self):
        self.t.start()

    def stop(self):
        self.stop_flag = True
        self.t.join(timeout=6)
        sm = [float(r[0]) for r in self.rows if r[0].replace(".", "").isdigit()]
        mx = [float(r[1]) for r in self.rows if r[1].replace(".", "").isdigit()]
        pw = [float(r[2]) for r in self.rows if r[2].replace(".", "").isdigit()]
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = [n for i, n in enumerate(names) if any(r[3 + i].lower().startswith("active") for r in self.rows)]
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "power_w_max": max(pw) if pw else None, "samples": len(self.rows), "reasons": reasons}


# --------------------------------------------------------------------------------------------
# CPU baseline: the oracle on a Q-slice of the same workload
# --------------------------------------------------------------------------------------------
def cpu_baseline(cfg, keep, amp, C, Crl, slice_rows, steps=1, warmup=0):
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import dfjk_oracle as oracle  # bench.py's cpu_baseline / --impl reference legs may use oracle/

    from psi4_b200 import workloads

    nbf, naux = cfg["nbf"], cfg["naux"]
    cores = os.cpu_count() or 1
    oracle.lib().oracle_set_blas_threads(cores)
    if not slice_rows:
        # ~15 s of CPU work at an assumed 25 GFLOP/s/core: K flops per Q row = 2*(P*o) + 2*N^2*o (+ second transform)
        per_row = (2.0 * keep.sum() * cfg["nocc"] * (1 if Crl is None else 2) + 2.0 * nbf * nbf * cfg["nocc"]) * cfg["nmat"]
        slice_rows = int(max(8, min(naux, 15.0 * cores * 25e9 / per_row)))
        slice_rows = min(slice_rows, int(4e9 / (8.0 * keep.sum())))  # <= 4 GB host slice
    sp = oracle.Sparsity(keep.astype(np.uint8), slice_rows)
    P = oracle.synth_fill(sp, 0, slice_rows, workloads.SEED, amp)
    Cl = [C] * cfg["nmat"]
    D = [C @ (C if Crl is None else Crl[i]).T for i in range(cfg["nmat"])]
    # oracle/_ref/libref_dfjk.so = the reference's own DFHelper::build_JK and callees, compiled from its source by
    # oracle/ref_build.py (prebuilt file on the GPU box); the C restatement is the fallback
    try:
        impl = "ref" if oracle.ref_lib() is not None else "port"
    except Exception:  # a prebuilt library that does not load here: time the restatement instead
        impl = "port"
    times, parts = [], None
    for it in range(warmup + steps):
        t0 = time.perf_counter()
        _, _, _, tm = oracle.build_JK(sp, P, Cl, Crl, D=D, nthreads=cores, impl=impl)
        dt = time.perf_counter() - t0
        if it >= warmup:
            times.append(dt)
            parts = tm
    scale = naux / slice_rows
    ms = float(np.mean(times)) * 1e3 * scale
    what = ("reference object code (DFHelper::build_JK and callees sliced from lib3index/dfhelper.cc, compiled unmodified; "
            if impl == "ref" else "oracle (C restatement of dfhelper.cc JK loops; ")
    out = {"value": ms, "unit": "ms", "cores": cores, "kind": "reference" if impl == "ref" else "port",
           "sample": f"{what}OpenMP+OpenBLAS {cores} threads) on Q rows [0,{slice_rows}) of naux={naux}, measured "
                     f"{np.mean(times) * 1e3:.1f} ms x {scale:.2f} (linear in Q)",
           "blas": oracle.lib().oracle_blas_config().decode()}
    if impl == "port":
        out.update({"J_ms": parts["J"] * 1e3 * scale, "K_ms": parts["K"] * 1e3 * scale})
    return out, times


_OUT = None


def emit(line):
    """The ONE JSON line of the contract, on the process's original stdout."""
    print(json.dumps(line), file=_OUT or sys.stdout, flush=True)


def main():
    global _OUT
    args = parse()
    # stdout carries exactly one JSON line: anything a library prints there (NCCL's version banner under NCCL_DEBUG,
    # OpenBLAS notices) is sent to stderr by pointing fd 1 at fd 2 and keeping a private copy of the real stdout
    sys.stdout.flush()
    _OUT = os.fdopen(os.dup(1), "w")
    os.dup2(2, 1)
    from psi4_b200 import workloads

    cfg = dict(workloads.CONFIGS[args.workload])
    if args.response:
        cfg["nmat"] = args.response
    nbf, naux, nocc, nmat = cfg["nbf"], cfg["naux"], cfg["nocc"], cfg["nmat"]
    keep = workloads.pair_mask(nbf, cfg["mask"])
    amp = workloads.amplitude(nbf)
    C = workloads.orbitals(nbf, nocc)
    Cr = workloads.orbitals(nbf, nocc, workloads.SEED + 1) if (args.nonsymmetric or args.response) else None
    # one (C_left, C_right, D) triple per matrix.  Default: the same pair nmat times, each uploaded and transformed on
    # its own.  --response: ONE C_left object and a different C_right per right-hand side.
    Cl = [C] * nmat
    if args.response:
        Crl = [workloads.orbitals(nbf, nocc, workloads.SEED + 1 + i) for i in range(nmat)]
    else:
        Crl = None if Cr is None else [Cr] * nmat
    D = [C @ C.T] * nmat if Crl is None else ([C @ Cr.T] * nmat if not args.response else [C @ x.T for x in Crl])
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    config = {"workload": f"{args.workload}: nbf={nbf} naux={naux} nocc={nocc} nmat={nmat} "
                          f"kept_pairs={int(keep.sum())} lr_symmetric={Cr is None} do_J=1 do_K=1"
                          + (f" response: {nmat} right-hand sides share one C_left" if args.response else ""),
              "pair_mask": workloads.mask_info(cfg["mask"]),
              "q_sharding": f"Q split over {world} rank(s)" + ("; every rank uploads C/D, rank 0 reads J/K" if world > 1 else ""), "l2": "inputs_exceed_l2 (tensor shard >> 126 MB)"}

    if args.impl == "reference":
        if rank != 0:
            return
        cb, times = cpu_baseline(cfg, keep, amp, C, Crl, args.cpu_slice, steps=args.steps, warmup=args.warmup)
        line = {"metric": METRIC, "value": cb["value"], "unit": "ms", "n_gpus": args.gpus, "steps": args.steps,
                "warmup": args.warmup, "ms_per_step": cb["value"], "higher_is_better": False, "scaling": "strong",
                "vs_baseline": None, "dtype": "f64", "data": "synthetic", "config": config, "impl": "reference",
                "cpu_baseline": cb,
                "e2e": {"value": cb["value"], "unit": "ms", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
                "gpu_launches": 0}
        emit(line)
        return

    import torch
    import torch.distributed as dist

    from psi4_b200 import DFHelper, Engine

    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device; the B200 JK engine has no CPU fallback")
    torch.cuda.set_device(local_rank)
    nccl_id = None
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
        idt = torch.zeros(128, dtype=torch.uint8, device="cuda")
        if rank == 0:
            idt = torch.tensor(list(Engine.nccl_unique_id()), dtype=torch.uint8, device="cuda")
        dist.broadcast(idt, 0)
        nccl_id = bytes(idt.cpu().tolist())

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def max_over_ranks(x):
        if world == 1:
            return x
        t = torch.tensor([x], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    d = DFHelper(nbf, naux)
    d.prepare_sparsity(keep=keep)
    if args.single_process and world == 1 and args.gpus > 1:
        # psi4's situation: ONE process drives all GPUs (b200jk_create(ngpu)); only the host-pointer call exists here
        eng = Engine(args.gpus)
        eng.set_layout(nbf, naux, d.small_skips_, d.big_skips_, d.schwarz_fun_index_)
        eng.fill_synthetic(0, workloads.SEED, amp)
        for x in {id(x): x for x in D}.values():
            eng.register_host(x)
        for _ in range(args.warmup):
            eng.compute(Cl, Crl, D, reuse_outputs=True)
        dev_ms, launches, parts = 0.0, 0, {"ms_j": 0.0, "ms_half": 0.0, "ms_kgemm": 0.0, "ms_allreduce": 0.0}
        w0 = time.perf_counter()
        for _ in range(args.steps):
            eng.compute(Cl, Crl, D, reuse_outputs=True)
            st = eng.stats()
            dev_ms += st["ms_total"] / args.steps
            launches += st["launches"]
            for k in parts:
                parts[k] += st[k] / args.steps
        e2e_ms = (time.perf_counter() - w0) * 1e3 / args.steps
        n2b = nbf * nbf * 8
        emit({
            "metric": METRIC, "value": dev_ms, "unit": "ms", "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": dev_ms, "higher_is_better": False, "scaling": "strong", "vs_baseline": None, "dtype": "f64",
            "data": "synthetic", "config": dict(config, q_sharding=f"Q split over {args.gpus} GPUs driven by ONE process"),
            "e2e": {"value": e2e_ms, "unit": "ms", "h2d_bytes_per_step": int(args.gpus * nmat * (C.nbytes + n2b)),
                    "d2h_bytes_per_step": int(nmat * 2 * n2b)},
            "gpu_launches": int(launches), "kernels_ms_max_over_gpus": parts, "mode": "single_process"})
        return
    eng = Engine(rank=rank, world=world, device=local_rank, nccl_id=nccl_id)
    eng.set_layout(nbf, naux, d.small_skips_, d.big_skips_, d.schwarz_fun_index_)
    t0 = time.perf_counter()
    eng.fill_synthetic(0, workloads.SEED, amp)
    fill_s = time.perf_counter() - t0

    n2b = nbf * nbf * 8
    # device operands mirror the host lists: a matrix passed twice is one buffer passed twice only under --response
    # (the engine recognises a repeated C_left by identity there); otherwise every matrix gets its own copy
    dC = [eng.dev_put(C)] * nmat if args.response else [eng.dev_put(C) for _ in range(nmat)]
    dCr = None if Crl is None else [eng.dev_put(x) for x in Crl]
    dD = [eng.dev_put(x) for x in D]
    dJ = [eng.dev_alloc(n2b) for _ in range(nmat)]
    dK = [eng.dev_alloc(n2b) for _ in range(nmat)]
    noccs = [nocc] * nmat

    pk_dmma = eng.fp64_peak(0, 0.0 if args.skip_probes else 1.5) if rank == 0 else None
    pk_dfma = eng.fp64_peak(1, 0.0 if args.skip_probes else 0.5) if rank == 0 else None
    # kernels timed inside a long step -> sustained ceiling (B200_PROFILING.md); the burst figure is reported too
    peak_dmma = pk_dmma["sustained_tflops"] if pk_dmma else 0.0

    # ---- kernel-only arm: operands resident in HBM ----
    for _ in range(args.warmup):
        eng.compute_device(dC, dCr, noccs, dD, dJ, dK, None)
    clocks = Clocks(local_rank)
    if rank == 0:
        clocks.start()
    barrier()
    w0 = time.perf_counter()
    dev_ms, parts = 0.0, {"ms_j": 0.0, "ms_half": 0.0, "ms_kgemm": 0.0, "ms_allreduce": 0.0}
    launches = 0
    for _ in range(args.steps):
        eng.compute_device(dC, dCr, noccs, dD, dJ, dK, None)
        st = eng.stats()
        dev_ms += st["ms_total"]
        for k in parts:
            parts[k] += st[k]
        launches += st["launches"]
    barrier()
    wall_ms = (time.perf_counter() - w0) * 1e3 / args.steps
    value = max_over_ranks(dev_ms / args.steps)
    wall_ms = max_over_ranks(wall_ms)
    for k in parts:
        parts[k] = max_over_ranks(parts[k] / args.steps)
    st_dev = eng.stats()

    # ---- end-to-end arm: host pointers through b200jk_compute ----
    # D, J, K live in persistent caller matrices, as psi4's D_ao_/J_ao_/K_ao_ do (allocated once, jk.cc:355-446); the
    # glue page-locks them once (b200jk_register_host).  C is a fresh pageable array every iteration and is staged.
    for x in {id(x): x for x in D}.values():
        eng.register_host(x)
    # the SCF driver is one process: rank 0 reads the all-reduced result, the other ranks only contribute to it
    fetch = rank == 0
    for _ in range(min(args.warmup, 2)):
        eng.compute(Cl, Crl, D, reuse_outputs=True, fetch=fetch)
    barrier()
    w0 = time.perf_counter()
    e2e_parts = {"ms_h2d": 0.0, "ms_d2h": 0.0}
    for _ in range(args.steps):
        J, K, _ = eng.compute(Cl, Crl, D, reuse_outputs=True, fetch=fetch)  # persistent J/K, as psi4's JK owns them
        st = eng.stats()
        launches += st["launches"]
        for k in e2e_parts:
            e2e_parts[k] += st[k] / args.steps
    barrier()
    e2e_ms = max_over_ranks((time.perf_counter() - w0) * 1e3 / args.steps)
    clk = clocks.stop() if rank == 0 else None

    # sanity: the two arms agree bit for bit (deterministic reductions)
    Jd = eng.dev_get(dJ[0], (nbf, nbf))
    Kd = eng.dev_get(dK[0], (nbf, nbf))
    arms_equal = bool(np.array_equal(Jd, J[0]) and np.array_equal(Kd, K[0])) if fetch else True

    if rank != 0:
        if world > 1:
            dist.barrier()
            dist.destroy_process_group()
        return

    # ---- roofline of the dominant kernel (K3 half transform) + the others for context ----
    hbm_peak, peak_src = 6650.0, "fallback"
    try:
        mp = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
        hbm_peak, peak_src = float(mp["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
    except Exception:
        pass
    half_tf = st_dev["half_flops"] / (parts["ms_half"] * 1e-3) / 1e12 if parts["ms_half"] else 0.0
    kg_tf = st_dev["kgemm_flops"] / (parts["ms_kgemm"] * 1e-3) / 1e12 if parts["ms_kgemm"] else 0.0
    j_gbs = st_dev["j_bytes"] / (parts["ms_j"] * 1e-3) / 1e9 if parts["ms_j"] else 0.0
    traffic = None
    try:
        traffic = json.load(open(os.path.join(ROOT, "profiles", "traffic.json"))).get(args.workload, {}).get("half_transform")
    except Exception:
        pass
    roofline = {"kernel": "half_transform_kernel (K3)", "bound": "tensor", "achieved": half_tf, "peak": peak_dmma,
                "unit": "TFLOP/s", "frac": half_tf / peak_dmma if peak_dmma else None, "traffic": traffic,
                "peak_source": "sustained FP64 DMMA m8n8k4 register-resident loop (1.5 s back to back) measured live "
                               "on this GPU; no FP64 entry in MEASURED_PEAKS.json; datasheet 37-40 TFLOP/s",
                "peak_burst": pk_dmma["burst_tflops"], "frac_of_burst": half_tf / pk_dmma["burst_tflops"],
                "dmma_probe": pk_dmma, "dfma_probe": pk_dfma}
    kernels = {
        "half_transform": {"ms": parts["ms_half"], "tflops": half_tf, "frac_of_dmma_peak": half_tf / peak_dmma if peak_dmma else None,
                           "hbm_read_gbs": st_dev["half_bytes"] / (parts["ms_half"] * 1e-3) / 1e9 if parts["ms_half"] else 0.0},
        "k_gemm": {"ms": parts["ms_kgemm"], "tflops": kg_tf, "frac_of_dmma_peak": kg_tf / peak_dmma if peak_dmma else None},
        "j_sweeps": {"ms": parts["ms_j"], "gbs": j_gbs, "frac_of_hbm_peak": j_gbs / hbm_peak, "hbm_peak_gbs": hbm_peak,
                     "hbm_peak_source": peak_src},
        "allreduce": {"ms": parts["ms_allreduce"]},
    }

    cb = None
    if world == 1 and not args.no_cpu_baseline:
        cb, _ = cpu_baseline(cfg, keep, amp, C, Crl, args.cpu_slice)

    line = {
        "metric": METRIC, "value": value, "unit": "ms", "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": value, "higher_is_better": False, "scaling": "strong", "vs_baseline": None, "dtype": "f64",
        "data": "synthetic", "config": config,
        "e2e": {"value": e2e_ms, "unit": "ms", "h2d_bytes_per_step": int(nmat * (C.nbytes * (1 if Cr is None else 2) + n2b)),
                "d2h_bytes_per_step": int(nmat * 2 * n2b), "ms_h2d": e2e_parts["ms_h2d"], "ms_d2h": e2e_parts["ms_d2h"]},
        "gpu_launches": int(launches), "roofline": roofline, "cpu_baseline": cb, "kernels": kernels, "clocks": clk,
        "wall_ms_per_step": wall_ms, "arms_bit_identical": arms_equal,
        "hbm": {"tensor_gb": st_dev["hbm_tensor_bytes"] / 1e9, "work_gb": st_dev["hbm_work_bytes"] / 1e9, "fill_s": fill_s},
    }
    emit(line)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
