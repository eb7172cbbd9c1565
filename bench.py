#!/usr/bin/env python
"""bench.py -- DF-JK ms per SCF iteration (one MemDFJK::compute_JK build) on the BASELINE.json
workload: C60 / cc-pVTZ DF-RHF (nbf 1800, naux 4740, nocc 180), synthetic tensor of that shape.

  python bench.py [--gpus N] [--steps K] [--warmup W] [--workload c60_tz] [--impl reference]

One process per GPU (torchrun for N>1): the auxiliary index Q is sharded over the ranks and the
partial J/K are summed over NVLink inside the engine ("strong" scaling: total work fixed).

  value        ms per build with C/D/J/K resident in HBM (b200jk_compute_device), device time
               from CUDA events on the engine's stream, max over ranks
  e2e          ms per build through the host-pointer C ABI call b200jk_compute (what psi4's
               MemDFJK::compute_JK would call): pinned staging + H2D of C and D + D2H of J and K inside
  roofline     the dominant kernel of the timed builds with the roof that binds it.  On the default arms of this
               workload (both GEMMs on the INT8 tensor cores by residues) that is the K3 residue GEMM, a stream of
               the residue planes: algorithmic bytes / CUDA-event time against hbm_gbs of MEASURED_PEAKS.json, the
               int8 tensor view (2 x the measured bf16 figure) beside it; kernels.*.int8_arm lists every sub-phase
  roofline_fp64  K3 on the FP64 tensor pipe (what the north_star names), measured in the same run on the FP64
               arms (`fp64_arms`): FP64 DMMA ceiling measured live by a register-resident m8n8k4 loop
               (MEASURED_PEAKS.json has no FP64 figure), with a cuBLAS DGEMM of the K GEMM's shape timed beside
               it as the library comparator
  parity_spot  outside the timed region, at every N: rows of J and a sample of K elements of rank 0's
               summed result recomputed on the host from the counter hash that defines the synthetic
               tensor (the oracle as CHECKER); the run exits non-zero above 1e-10
  workloads    the other BASELINE.json configurations that fit this launch (n-C20H42 at every N,
               (H2O)40 UHF at 8 GPUs), same measurements, C60 stays the headline `value`
  cpu_baseline the reference's own object code (oracle/_ref) or the oracle restatement on this box's host
               cores, on a Q-slice of the same workload, extrapolated linearly in naux (every hot loop is
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
SPOT_TOL = 1e-10


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--workload", default="c60_tz")
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--cpu-slice", type=int, default=0, help="Q rows in the CPU baseline sample (0 = auto)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-extra", action="store_true", help="headline workload only (no `workloads` block, no cuBLAS calibration)")
    ap.add_argument("--no-spot", action="store_true", help="skip the host-side spot parity (ncu / profiling runs)")
    ap.add_argument("--skip-probes", action="store_true", help="no FP64 ceiling probes (ncu launch lists); roofline.peak = last recorded")
    ap.add_argument("--kgemm", default="auto", choices=["auto", "dmma", "i8"],
                    help="arm of the K GEMM: FP64 tensor pipe (dmma) or INT8 tensor cores by residues (i8); auto = the engine's rule")
    ap.add_argument("--half", default="auto", choices=["auto", "dmma", "i8"], help="the same for the half transform")
    ap.add_argument("--no-ab", action="store_true", help="skip the extra timed builds with both arms on the FP64 tensor pipe")
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


def _oracle():
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import dfjk_oracle as oracle  # bench.py may use oracle/ as the CPU baseline and as the checker, never as the product

    return oracle


# --------------------------------------------------------------------------------------------
# CPU baseline: the reference's object code (or the oracle) on a Q-slice of the same workload
# --------------------------------------------------------------------------------------------
def cpu_baseline(cfg, keep, amp, C, Crl, slice_rows, steps=1, warmup=0):
    oracle = _oracle()
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
                     f"{np.mean(times) * 1e3:.1f} ms x {scale:.2f} (linear in Q; the per-m DGEMMs of the slice are skinnier "
                     f"than at {naux} rows, so the extrapolation flatters the GPU)",
           "blas": oracle.lib().oracle_blas_config().decode()}
    if impl == "port":
        out.update({"J_ms": parts["J"] * 1e3 * scale, "K_ms": parts["K"] * 1e3 * scale})
    return out, times


# --------------------------------------------------------------------------------------------
# spot parity: rows of J and a sample of K recomputed on the host from the counter hash (checker only)
# --------------------------------------------------------------------------------------------
def spot_parity(keep, naux, amp, Cl, Crl, D, J, K, seed):
    """max |J - J_ref| over four whole rows and max |K - K_ref| over the 4 x 4 elements they span, per density,
    scaled by max(1, |ref|) as tests/test_gpu_fullsize.py does.  Cost: one pass of naux * kept-pairs hash
    evaluations per density (seconds on the box's cores)."""
    oracle = _oracle()
    nbf = keep.shape[0]
    keep8 = keep.astype(np.uint8)
    rows = sorted({0, min(7, nbf - 1), min(nbf // 2 + 1, nbf - 1), nbf - 1})
    t0 = time.perf_counter()
    Bm = {m: oracle.synth_rowblock(keep8, naux, seed, amp, m) for m in rows}  # (naux, nbf) each
    out = {"rows": rows, "max_abs_J": 0.0, "max_abs_K": 0.0, "max_scaled_J": 0.0, "max_scaled_K": 0.0, "k_elements": 0}
    dq_cache = {}
    for i in range(len(D)):
        lr = Crl is None
        Di = np.triu(D[i]) + np.triu(D[i], 1).T if lr else D[i]  # the symmetric path reads the upper triangle (:3188)
        key = id(D[i])
        if key not in dq_cache:
            dq_cache[key] = oracle.synth_dq(keep8, naux, seed, amp, Di)
        dq = dq_cache[key]
        Tl = {m: Bm[m] @ Cl[i] for m in rows}
        Tr = Tl if lr else {m: Bm[m] @ Crl[i] for m in rows}
        for m in rows:
            jref = Bm[m].T @ dq
            dj = float(np.abs(J[i][m] - jref).max())
            out["max_abs_J"] = max(out["max_abs_J"], dj)
            out["max_scaled_J"] = max(out["max_scaled_J"], dj / max(1.0, float(np.abs(jref).max())))
            for n in rows:
                kref = float(np.vdot(Tl[m], Tr[n]))
                dk = abs(float(K[i][m, n]) - kref)
                out["max_abs_K"] = max(out["max_abs_K"], dk)
                out["max_scaled_K"] = max(out["max_scaled_K"], dk / max(1.0, abs(kref)))
                out["k_elements"] += 1
    out["seconds"] = time.perf_counter() - t0
    out["tolerance"] = SPOT_TOL
    out["ok"] = bool(out["max_scaled_J"] < SPOT_TOL and out["max_scaled_K"] < SPOT_TOL)
    return out


def cublas_dgemm_calibration(nbf, kdim):
    """cuBLAS DGEMM (through torch.mm on float64) at the K GEMM's own shape -- K = T T^T, nbf x nbf x kdim with both
    operands k-contiguous, the full square as the reference's C_DGEMM('N','T') executes it (dfhelper.cc:3374) -- and
    at 8192^3, on this box in this run.  The library comparator of kernels.k_gemm."""
    import torch

    out = {}
    free, _ = torch.cuda.mem_get_info()
    kd = int(min(kdim, (free * 0.5) // (8 * nbf)))
    gen = torch.Generator(device="cuda").manual_seed(1)
    for tag, (m, k) in {"k_gemm_shape": (nbf, kd), "8192_cubed": (8192, 8192)}.items():
        a = torch.randn((m, k), dtype=torch.float64, device="cuda", generator=gen)
        ev = [torch.cuda.Event(enable_timing=True) for _ in range(2)]
        torch.mm(a, a.T)
        torch.cuda.synchronize()
        best = 1e30
        for _ in range(3):
            ev[0].record()
            torch.mm(a, a.T)
            ev[1].record()
            torch.cuda.synchronize()
            best = min(best, ev[0].elapsed_time(ev[1]))
        out[tag] = {"m": m, "n": m, "k": k, "ms": best, "tflops": 2.0 * m * m * k / (best * 1e-3) / 1e12}
        del a
        torch.cuda.empty_cache()
    return out


_OUT = None


def emit(line):
    """The ONE JSON line of the contract, on the process's original stdout."""
    print(json.dumps(line), file=_OUT or sys.stdout, flush=True)


class Dist:
    """rank / world plumbing (torch.distributed over NCCL when launched by torchrun)."""

    def __init__(self):
        import torch

        self.torch = torch
        self.rank = int(os.environ.get("RANK", "0"))
        self.world = int(os.environ.get("WORLD_SIZE", "1"))
        self.local_rank = int(os.environ.get("LOCAL_RANK", "0"))
        if not torch.cuda.is_available():
            raise SystemExit("bench.py: no CUDA device; the B200 JK engine has no CPU fallback")
        torch.cuda.set_device(self.local_rank)
        self.dist = None
        if self.world > 1:
            import torch.distributed as dist

            dist.init_process_group("nccl", device_id=torch.device("cuda", self.local_rank))
            self.dist = dist

    def nccl_id(self, Engine):
        if self.world == 1:
            return None
        torch = self.torch
        idt = torch.zeros(128, dtype=torch.uint8, device="cuda")
        if self.rank == 0:
            idt = torch.tensor(list(Engine.nccl_unique_id()), dtype=torch.uint8, device="cuda")
        self.dist.broadcast(idt, 0)
        return bytes(idt.cpu().tolist())

    def barrier(self):
        self.torch.cuda.synchronize()
        if self.dist:
            self.dist.barrier()
        self.torch.cuda.synchronize()

    def max(self, x):
        if not self.dist:
            return x
        t = self.torch.tensor([x], dtype=self.torch.float64, device="cuda")
        self.dist.all_reduce(t, op=self.dist.ReduceOp.MAX)
        return float(t.item())

    def all_true(self, flag):
        return self.max(0.0 if flag else 1.0) == 0.0

    def close(self):
        if self.dist:
            self.dist.barrier()
            self.dist.destroy_process_group()


def make_inputs(args, name):
    from psi4_b200 import workloads

    cfg = dict(workloads.CONFIGS[name])
    if args.response:
        cfg["nmat"] = args.response
    nbf, nocc, nmat = cfg["nbf"], cfg["nocc"], cfg["nmat"]
    keep = workloads.pair_mask(nbf, cfg["mask"])
    amp = workloads.amplitude(nbf)
    nonsym = args.nonsymmetric or args.response
    # one (C_left, C_right, D) triple per matrix.  RHF workloads pass the same pair nmat times; the UHF workload
    # ((H2O)40, nmat = 2) gets two different occupied blocks (alpha / beta); --response: ONE C_left object and a
    # different C_right per right-hand side.
    if nmat > 1 and not args.response:
        Cl = [workloads.orbitals(nbf, nocc, workloads.SEED + 10 * i) for i in range(nmat)]
    else:
        Cl = [workloads.orbitals(nbf, nocc)] * nmat
    if args.response:
        Crl = [workloads.orbitals(nbf, nocc, workloads.SEED + 1 + i) for i in range(nmat)]
    elif nonsym:
        Crl = [workloads.orbitals(nbf, nocc, workloads.SEED + 1)] * nmat
    else:
        Crl = None
    D = [Cl[i] @ (Cl[i] if Crl is None else Crl[i]).T for i in range(nmat)]
    return cfg, keep, amp, Cl, Crl, D


def measure(args, ds, name, steps, warmup, headline):
    """Device-resident arm + host-pointer arm + spot parity of one workload at this launch's N."""
    from psi4_b200 import DFHelper, Engine, workloads

    rank, world = ds.rank, ds.world
    cfg, keep, amp, Cl, Crl, D = make_inputs(args, name)
    nbf, naux, nocc, nmat = cfg["nbf"], cfg["naux"], cfg["nocc"], cfg["nmat"]
    n2b = nbf * nbf * 8
    d = DFHelper(nbf, naux)
    d.prepare_sparsity(keep=keep)
    t0 = time.perf_counter()
    eng = Engine(rank=rank, world=world, device=ds.local_rank, nccl_id=ds.nccl_id(Engine))
    eng.set_layout(nbf, naux, d.small_skips_, d.big_skips_, d.schwarz_fun_index_)
    eng.set_kgemm(args.kgemm)
    eng.set_half(args.half)
    layout_s = time.perf_counter() - t0
    t0 = time.perf_counter()
    eng.fill_synthetic(0, workloads.SEED, amp)
    fill_s = time.perf_counter() - t0

    # device operands mirror the host lists: a matrix passed twice is one buffer passed twice only under --response
    # (the engine recognises a repeated C_left by identity there); otherwise every matrix gets its own copy
    dC = [eng.dev_put(Cl[0])] * nmat if args.response else [eng.dev_put(x) for x in Cl]
    dCr = None if Crl is None else [eng.dev_put(x) for x in Crl]
    dD = [eng.dev_put(x) for x in D]
    dJ = [eng.dev_alloc(n2b) for _ in range(nmat)]
    dK = [eng.dev_alloc(n2b) for _ in range(nmat)]
    noccs = [nocc] * nmat

    pk_dmma = pk_dfma = None
    if headline and rank == 0:
        pk_dmma = eng.fp64_peak(0, 0.0 if args.skip_probes else 1.5)
        pk_dfma = eng.fp64_peak(1, 0.0 if args.skip_probes else 0.5)

    # ---- kernel-only arm: operands resident in HBM ----
    import zlib

    def dev_crcs():
        return [zlib.crc32(eng.dev_get(p, (nbf, nbf)).tobytes()) for p in dJ + dK]

    dev_first = None
    for _ in range(warmup):
        eng.compute_device(dC, dCr, noccs, dD, dJ, dK, None)
        if dev_first is None:
            dev_first = dev_crcs()  # the device-operand arm's own run-to-run check (first build vs last)
    clocks = Clocks(ds.local_rank)
    if rank == 0:
        clocks.start()
    ds.barrier()
    w0 = time.perf_counter()
    dev_ms, parts = 0.0, {"ms_j": 0.0, "ms_half": 0.0, "ms_kgemm": 0.0, "ms_allreduce": 0.0}
    launches = 0
    sub = {"ms_half_i8": [0.0] * 4, "ms_kgemm_i8": [0.0] * 3}  # sub-phases of the INT8 arms (zero on the DMMA arms)
    for _ in range(steps):
        eng.compute_device(dC, dCr, noccs, dD, dJ, dK, None)
        st = eng.stats()
        dev_ms += st["ms_total"]
        for k in parts:
            parts[k] += st[k]
        for k in sub:
            sub[k] = [a + b for a, b in zip(sub[k], st[k])]
        launches += st["launches"]
    ds.barrier()
    wall_ms = ds.max((time.perf_counter() - w0) * 1e3 / steps)
    value = ds.max(dev_ms / steps)
    for k in parts:
        parts[k] = ds.max(parts[k] / steps)
    for k in sub:
        sub[k] = [ds.max(x / steps) for x in sub[k]]
    st_dev = eng.stats()

    # ---- end-to-end arm: host pointers through b200jk_compute ----
    # D, J, K live in persistent caller matrices, as psi4's D_ao_/J_ao_/K_ao_ do (allocated once, jk.cc:355-446); the
    # glue page-locks them once (b200jk_register_host).  C is a fresh pageable array every iteration and is staged.
    for x in {id(x): x for x in D}.values():
        eng.register_host(x)
    # the SCF driver is one process: rank 0 reads the summed result, the other ranks only contribute to it
    fetch = rank == 0
    first = None
    for _ in range(min(warmup, 2)):
        J, K, _ = eng.compute(Cl, Crl, D, reuse_outputs=True, fetch=fetch)
        if fetch and first is None:
            first = ([x.copy() for x in J], [x.copy() for x in K])
    ds.barrier()
    w0 = time.perf_counter()
    e2e_parts = {"ms_h2d": 0.0, "ms_d2h": 0.0}
    for _ in range(steps):
        J, K, _ = eng.compute(Cl, Crl, D, reuse_outputs=True, fetch=fetch)  # persistent J/K, as psi4's JK owns them
        st = eng.stats()
        launches += st["launches"]
        for k in e2e_parts:
            e2e_parts[k] += st[k] / steps
    ds.barrier()
    e2e_ms = ds.max((time.perf_counter() - w0) * 1e3 / steps)
    clk = clocks.stop() if rank == 0 else None

    # ---- determinism: the two arms, and the first and the last host-arm build, agree bit for bit ----
    arms_equal = run_equal = True
    arms_diff = {}
    if fetch:
        for i in range(nmat):
            for tag, dev, host in (("J", dJ[i], J[i]), ("K", dK[i], K[i])):
                got = eng.dev_get(dev, (nbf, nbf))
                if not np.array_equal(got, host):
                    arms_equal = False
                    arms_diff[f"{tag}[{i}]"] = {"max_abs": float(np.abs(got - host).max()),
                                                "n_differ": int(np.count_nonzero(got != host))}
            if first is not None:
                run_equal = run_equal and np.array_equal(first[0][i], J[i]) and np.array_equal(first[1][i], K[i])
    # every rank's device-arm result is the same bits (the sum is formed once per element and broadcast)
    dev_last = dev_crcs()
    if dev_first is not None and dev_first != dev_last:
        run_equal = False
        arms_diff["device_arm_run_to_run"] = "first and last device-operand build differ"
    ranks_equal = all(ds.max(float(c)) == -ds.max(-float(c)) for c in dev_last)

    # ---- spot parity against the on-the-fly oracle (rank 0, outside every timed region) ----
    spot = None
    if fetch and not args.no_spot:
        spot = spot_parity(keep, naux, amp, Cl, Crl, D, J, K, workloads.SEED)
    # ---- one-time setup cost that is NOT in the per-iteration metric: Matrix::power of the naux x naux fitting metric
    # (libmints/matrix.cc:2370-2424; tens of seconds of LAPACK on the host cores in the reference) on the device ----
    setup = None
    if headline and fetch and not args.no_extra:
        try:
            rng = np.random.default_rng(7)
            U = rng.standard_normal((naux, 48))
            Ms = np.eye(naux) + (U @ U.T) / 48.0  # well-conditioned SPD stand-in for (A|B)
            t0 = time.perf_counter()
            _, kept, ms_dev = eng.matrix_power(Ms, -0.5, 1e-10, with_info=True)
            setup = {"what": f"b200jk_matrix_power, naux = {naux}, alpha = -1/2 (cuSOLVER dsyevd + engine kernels; includes "
                             f"H2D/D2H of the matrix in the wall time)", "device_ms": ms_dev,
                     "wall_s": time.perf_counter() - t0, "eigenvalues_kept": int(kept)}
        except Exception as ex:  # never let the extra cost the bench line
            setup = {"error": str(ex)[:200]}
    # ---- A/B: the same build with both GEMMs on the FP64 tensor pipe (DMMA), device-resident operands, outside the timed
    # region of the headline: the north_star's "DMMA utilisation against FP64 peak" stays measured when an INT8 arm is the default
    fp64_arms = None
    if headline and not args.no_ab and (st_dev["kgemm_kind"] or st_dev["half_kind"]):
        eng.set_kgemm("dmma")
        eng.set_half("dmma")
        eng.compute_device(dC, dCr, noccs, dD, dJ, dK, None)
        ab = {"ms_total": 0.0, "ms_j": 0.0, "ms_half": 0.0, "ms_kgemm": 0.0}
        nab = 3
        for _ in range(nab):
            eng.compute_device(dC, dCr, noccs, dD, dJ, dK, None)
            st = eng.stats()
            for k in ab:
                ab[k] += st[k] / nab
        for k in ab:
            ab[k] = ds.max(ab[k])
        ab["max_abs_dK_vs_default_arms"] = None
        if fetch:
            ab["max_abs_dK_vs_default_arms"] = float(max(np.abs(eng.dev_get(dK[i], (nbf, nbf)) - K[i]).max() for i in range(nmat)))
        fp64_arms = ab
        eng.set_kgemm(args.kgemm)
        eng.set_half(args.half)
    ds.barrier()
    reduce_kind = {0: "none (one GPU)", 1: "fixed-rank-order peer-memory kernel over NVLink (peer_reduce.cuh)",
                   2: "NCCL all-reduce"}.get(st_dev["reduce_kind"], "?")
    eng.close()
    res = dict(cfg=cfg, keep=keep, amp=amp, Cl=Cl, Crl=Crl, value=value, wall_ms=wall_ms, parts=parts, sub=sub, st_dev=st_dev,
               e2e_ms=e2e_ms, e2e_parts=e2e_parts, launches=launches, clk=clk, arms_equal=bool(arms_equal), arms_diff=arms_diff,
               run_equal=bool(run_equal), ranks_equal=bool(ranks_equal), spot=spot, pk_dmma=pk_dmma, pk_dfma=pk_dfma,
               layout_s=layout_s, fill_s=fill_s, reduce_kind=reduce_kind, setup=setup, fp64_arms=fp64_arms,
               h2d=int(nmat * (Cl[0].nbytes * (1 if Crl is None else 2) + n2b)), d2h=int(nmat * 2 * n2b))
    return res


def int8_peak():
    """INT8 dense tensor ceiling for the residue GEMMs.  MEASURED_PEAKS.json holds no int8 figure; on this part the dense
    int8 rate of tcgen05.mma is twice the bf16 one (4.5 vs 2.25 Pop/s nominal), so the ceiling is 2 x the MEASURED bf16
    numbers: sustained for kernels timed inside a long step (B200_PROFILING.md), burst beside it."""
    try:
        mp = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
        return {"sustained_tops": 2.0 * float(mp.get("bf16_tflops_sustained", mp["bf16_tflops"])), "burst_tops": 2.0 * float(mp["bf16_tflops"]),
                "source": "2 x MEASURED_PEAKS.json bf16 (sustained / burst); no int8 entry there"}
    except Exception:
        return {"sustained_tops": 4500.0, "burst_tops": 4500.0, "source": "nominal 4.5 Pop/s dense int8 (MEASURED_PEAKS.json absent)"}


def i8_table(res, hbm_peak):
    """Sub-phases of the INT8-tensor-core arms of the timed builds (CUDA events between their launches): every one with the
    roofline that bounds it -- the converter and the CRT are streaming kernels (HBM), the two GEMMs tensor work whose
    operands (residue planes) also come from HBM, so both views are given."""
    st, sub = res["st_dev"], res["sub"]
    pk = int8_peak()
    out = {}
    if st.get("half_kind"):
        cv, ga, ge, cr = sub["ms_half_i8"]
        cfg = res["cfg"]
        nq = st["q_end"] - st["q_begin"]
        ntr = cfg["nmat"] * (1 if res["Crl"] is None else 2)  # transforms per build
        t_bytes = 8.0 * cfg["nbf"] * nq * (cfg["nocc"] + cfg["nocc"] % 2) * ntr
        # algorithmic traffic of the residue GEMM: the A planes once, the gathered C^T once, the residue bytes it writes
        # (orbital tiling of i8h_tiling at one CTA per cluster: nocc rounded up to 16, tiles of <= 256 columns)
        opw0 = (cfg["nocc"] + 15) // 16 * 16
        nit = (opw0 + 255) // 256
        opw = nit * (((opw0 + nit - 1) // nit + 15) // 16 * 16)
        nqt = (nq + 127) // 128
        cg_bytes = st["half_i8_plane_bytes"] / (nqt * 128.0) * opw if nqt else 0.0
        ws_bytes = float(st["half_moduli"]) * cfg["nbf"] * nq * opw * ntr
        gemm_bytes = st["half_i8_plane_bytes"] + cg_bytes + ws_bytes
        out["half_transform"] = {
            "moduli": st["half_moduli"], "chunks": st["half_i8_chunks"], "planes_cached": st["half_i8_cached"],
            "resident_row_blocks": st["half_i8_resident_rows"],
            "convert": {"ms": cv, "bound": "hbm", "bytes": st["half_i8_convert_bytes"],
                        "gbs": st["half_i8_convert_bytes"] / (cv * 1e-3) / 1e9 if cv else 0.0,
                        "frac_of_hbm_peak": st["half_i8_convert_bytes"] / (cv * 1e-3) / 1e9 / hbm_peak if cv else None,
                        "what": "f64 rows read + residue planes written (+ the first J sweep riding on it)"},
            "gather": {"ms": ga},
            "gemm": {"ms": ge, "bound": "hbm (the planes are streamed once; the int8 tensor roof is the looser one)",
                     "int8_tops": st["half_i8_ops"] / (ge * 1e-3) / 1e12 if ge else 0.0,
                     "frac_of_int8_peak": st["half_i8_ops"] / (ge * 1e-3) / 1e12 / pk["sustained_tops"] if ge else None,
                     "plane_read_gbs": st["half_i8_plane_bytes"] / (ge * 1e-3) / 1e9 if ge else 0.0,
                     "bytes": gemm_bytes, "gbs": gemm_bytes / (ge * 1e-3) / 1e9 if ge else 0.0,
                     "frac_of_hbm_peak": gemm_bytes / (ge * 1e-3) / 1e9 / hbm_peak if ge else None,
                     "what": "tcgen05.mma.kind::i8 over all moduli, whole padded tiles; bytes = residue planes read once + "
                             "gathered C^T read once + residue bytes written (ncu on the 592-row shard: 26.6 GB of DRAM traffic "
                             "for 26.5 GB counted this way, profiles/r02_ncu_i8_q8_final2.json)"},
            "crt": {"ms": cr, "bound": "hbm", "t_write_gbs": t_bytes / (cr * 1e-3) / 1e9 if cr else 0.0},
        }
    if st.get("kgemm_kind"):
        pl, ge, cr = sub["ms_kgemm_i8"]
        out["k_gemm"] = {
            "moduli": st["kgemm_moduli"], "planes": {"ms": pl, "bound": "hbm"},
            "gemm": {"ms": ge, "bound": "tensor", "int8_tops": st["kgemm_i8_ops"] / (ge * 1e-3) / 1e12 if ge else 0.0,
                     "frac_of_int8_peak": st["kgemm_i8_ops"] / (ge * 1e-3) / 1e12 / pk["sustained_tops"] if ge else None},
            "crt": {"ms": cr},
        }
    if out:
        out["int8_peak"] = pk
    return out


def kernel_table(res, peak_dmma, hbm_peak, peak_src):
    parts, st = res["parts"], res["st_dev"]
    half_tf = st["half_flops"] / (parts["ms_half"] * 1e-3) / 1e12 if parts["ms_half"] else 0.0
    kg_tf = st["kgemm_flops"] / (parts["ms_kgemm"] * 1e-3) / 1e12 if parts["ms_kgemm"] else 0.0
    j_gbs = st["j_bytes"] / (parts["ms_j"] * 1e-3) / 1e9 if parts["ms_j"] else 0.0
    arm = {0: "FP64 tensor pipe (DMMA m8n8k4)", 1: "INT8 tensor cores (tcgen05.mma.kind::i8) by residues + CRT; tflops = FP64-equivalent"}
    kt = {
        "half_transform": {"ms": parts["ms_half"], "tflops": half_tf, "arm": arm[st.get("half_kind", 0)],
                           "frac_of_dmma_peak": half_tf / peak_dmma if peak_dmma else None,
                           "hbm_read_gbs": st["half_bytes"] / (parts["ms_half"] * 1e-3) / 1e9 if parts["ms_half"] else 0.0},
        "k_gemm": {"ms": parts["ms_kgemm"], "tflops": kg_tf, "arm": arm[st.get("kgemm_kind", 0)],
                   "moduli": st.get("kgemm_moduli", 0), "frac_of_dmma_peak": kg_tf / peak_dmma if peak_dmma else None},
        "j_sweeps": {"ms": parts["ms_j"], "gbs": j_gbs, "frac_of_hbm_peak": j_gbs / hbm_peak, "hbm_peak_gbs": hbm_peak,
                     "hbm_peak_source": peak_src},
        "cross_gpu_sum": {"ms": parts["ms_allreduce"], "how": res["reduce_kind"]},
    }
    i8 = i8_table(res, hbm_peak)
    for k in ("half_transform", "k_gemm"):
        if k in i8:
            kt[k]["int8_arm"] = i8[k]
    if i8:
        kt["int8_peak"] = i8["int8_peak"]
    return kt, half_tf


def single_process(args, cfg_name):
    """psi4's situation: ONE process drives all GPUs (b200jk_create(ngpu)); only the host-pointer call exists here."""
    from psi4_b200 import DFHelper, Engine, workloads

    cfg, keep, amp, Cl, Crl, D = make_inputs(args, cfg_name)
    nbf, naux, nocc, nmat = cfg["nbf"], cfg["naux"], cfg["nocc"], cfg["nmat"]
    d = DFHelper(nbf, naux)
    d.prepare_sparsity(keep=keep)
    eng = Engine(args.gpus)
    eng.set_layout(nbf, naux, d.small_skips_, d.big_skips_, d.schwarz_fun_index_)
    eng.fill_synthetic(0, workloads.SEED, amp)
    for x in {id(x): x for x in D}.values():
        eng.register_host(x)
    first = None
    for _ in range(args.warmup):
        J, K, _ = eng.compute(Cl, Crl, D, reuse_outputs=True)
        if first is None:
            first = ([x.copy() for x in J], [x.copy() for x in K])
    dev_ms, launches, parts = 0.0, 0, {"ms_j": 0.0, "ms_half": 0.0, "ms_kgemm": 0.0, "ms_allreduce": 0.0}
    w0 = time.perf_counter()
    for _ in range(args.steps):
        J, K, _ = eng.compute(Cl, Crl, D, reuse_outputs=True)
        st = eng.stats()
        dev_ms += st["ms_total"] / args.steps
        launches += st["launches"]
        for k in parts:
            parts[k] += st[k] / args.steps
    e2e_ms = (time.perf_counter() - w0) * 1e3 / args.steps
    run_equal = all(np.array_equal(a, b) for a, b in zip(first[0] + first[1], J + K))
    spot = None if args.no_spot else spot_parity(keep, naux, amp, Cl, Crl, D, J, K, workloads.SEED)
    n2b = nbf * nbf * 8
    emit({
        "metric": METRIC, "value": dev_ms, "unit": "ms", "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": dev_ms, "higher_is_better": False, "scaling": "strong", "vs_baseline": None, "dtype": "f64",
        "data": "synthetic", "config": {"workload": f"{cfg_name}: nbf={nbf} naux={naux} nocc={nocc} nmat={nmat}",
                                        "q_sharding": f"Q split over {args.gpus} GPUs driven by ONE process"},
        "e2e": {"value": e2e_ms, "unit": "ms", "h2d_bytes_per_step": int(args.gpus * nmat * (Cl[0].nbytes + n2b)),
                "d2h_bytes_per_step": int(nmat * 2 * n2b)},
        "gpu_launches": int(launches), "kernels_ms_max_over_gpus": parts, "mode": "single_process",
        "run_to_run_bit_identical": bool(run_equal), "parity_spot": spot,
        "cross_gpu_sum": {0: "none", 1: "peer-memory kernel", 2: "nccl"}.get(eng.stats()["reduce_kind"])})
    ok = run_equal and (spot is None or spot["ok"])
    eng.close()
    return 0 if ok else 2


def main():
    global _OUT
    args = parse()
    # stdout carries exactly one JSON line: anything a library prints there (NCCL's version banner under NCCL_DEBUG,
    # OpenBLAS notices) is sent to stderr by pointing fd 1 at fd 2 and keeping a private copy of the real stdout
    sys.stdout.flush()
    _OUT = os.fdopen(os.dup(1), "w")
    os.dup2(2, 1)
    from psi4_b200 import workloads

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))

    def config_of(name, cfg, keep, Crl):
        return {"workload": f"{name}: nbf={cfg['nbf']} naux={cfg['naux']} nocc={cfg['nocc']} nmat={cfg['nmat']} "
                            f"kept_pairs={int(keep.sum())} lr_symmetric={Crl is None} do_J=1 do_K=1"
                            + (f" response: {cfg['nmat']} right-hand sides share one C_left" if args.response else ""),
                "pair_mask": workloads.mask_info(cfg["mask"]),
                "q_sharding": f"Q split over {world} rank(s)" + ("; every rank uploads C/D, rank 0 reads J/K" if world > 1 else ""),
                "l2": "inputs_exceed_l2 (tensor shard >> 126 MB)"}

    if args.impl == "reference":
        if rank != 0:
            return 0
        cfg, keep, amp, Cl, Crl, D = make_inputs(args, args.workload)
        cb, times = cpu_baseline(cfg, keep, amp, Cl[0], Crl, args.cpu_slice, steps=args.steps, warmup=args.warmup)
        emit({"metric": METRIC, "value": cb["value"], "unit": "ms", "n_gpus": args.gpus, "steps": args.steps,
              "warmup": args.warmup, "ms_per_step": cb["value"], "higher_is_better": False, "scaling": "strong",
              "vs_baseline": None, "dtype": "f64", "data": "synthetic", "config": config_of(args.workload, cfg, keep, Crl),
              "impl": "reference", "cpu_baseline": cb,
              "e2e": {"value": cb["value"], "unit": "ms", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
              "gpu_launches": 0})
        return 0

    if args.single_process and world == 1 and args.gpus > 1:
        return single_process(args, args.workload)

    ds = Dist()
    res = measure(args, ds, args.workload, args.steps, args.warmup, headline=True)

    # the other BASELINE.json configurations this launch can hold: same measurements, fewer steps
    extra = {}
    if not args.no_extra and args.workload == "c60_tz" and not (args.nonsymmetric or args.response):
        names = ["c20h42_tz"] + (["h2o40_tz"] if world == 8 else [])
        for nm in names:
            extra[nm] = measure(args, ds, nm, max(3, min(args.steps, 5)), 3, headline=False)

    if rank != 0:
        ds.close()
        return 0

    # ---- roofline of the dominant kernel (K3 half transform) + the others for context ----
    hbm_peak, peak_src = 6650.0, "fallback"
    try:
        mp = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
        hbm_peak, peak_src = float(mp["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
    except Exception:
        pass
    pk_dmma, pk_dfma = res["pk_dmma"], res["pk_dfma"]
    # kernels timed inside a long step -> sustained ceiling (B200_PROFILING.md); the burst figure is reported too
    peak_dmma = pk_dmma["sustained_tflops"] if pk_dmma else 0.0
    kernels, half_tf = kernel_table(res, peak_dmma, hbm_peak, peak_src)
    traffic = None
    try:
        traffic = json.load(open(os.path.join(ROOT, "profiles", "traffic.json"))).get(args.workload, {}).get("half_transform")
    except Exception:
        pass
    ab = res["fp64_arms"]
    if ab and res["st_dev"]["half_kind"]:
        # the headline ran the half transform on the INT8 tensor cores: the FP64 roofline of K3 is that of the A/B builds
        half_tf = res["st_dev"]["half_flops"] / (ab["ms_half"] * 1e-3) / 1e12 if ab["ms_half"] else 0.0
    if ab:
        st0 = res["st_dev"]
        ab["half_transform_tflops"] = st0["half_flops"] / (ab["ms_half"] * 1e-3) / 1e12 if ab["ms_half"] else 0.0
        ab["k_gemm_tflops"] = st0["kgemm_flops"] / (ab["ms_kgemm"] * 1e-3) / 1e12 if ab["ms_kgemm"] else 0.0
        ab["half_transform_frac_of_dmma_peak"] = ab["half_transform_tflops"] / peak_dmma if peak_dmma else None
        ab["k_gemm_frac_of_dmma_peak"] = ab["k_gemm_tflops"] / peak_dmma if peak_dmma else None
        ab["what"] = "the same build with both GEMMs on the FP64 tensor pipe (B200JK_KGEMM=dmma B200JK_HALF=dmma), 3 device-resident builds"
    roofline_fp64 = {"kernel": "half_ws_kernel (K3, FP64 tensor pipe)", "bound": "tensor", "achieved": half_tf, "peak": peak_dmma,
                     "unit": "TFLOP/s", "frac": half_tf / peak_dmma if peak_dmma else None, "traffic": traffic,
                     "traffic_source": "profiles/traffic.json (ncu --set full capture of this kernel, recorded, not re-measured in this run)",
                     "peak_source": "sustained FP64 DMMA m8n8k4 register-resident loop (1.5 s back to back) measured live "
                                    "on this GPU; no FP64 entry in MEASURED_PEAKS.json; datasheet 37-40 TFLOP/s",
                     "peak_burst": pk_dmma["burst_tflops"], "frac_of_burst": half_tf / pk_dmma["burst_tflops"] if pk_dmma["burst_tflops"] else None,
                     "dmma_probe": pk_dmma, "dfma_probe": pk_dfma,
                     "measured_in": "the fp64_arms A/B builds of this run" if (ab and res["st_dev"]["half_kind"]) else "the timed region"}
    # `roofline` is the DOMINANT kernel of the timed builds.  On the FP64 arms that is K3 on the DMMA pipe; when the default
    # arms ran on the INT8 tensor cores it is whichever sub-phase of them took longest, with its own bound and ceiling.
    roofline = roofline_fp64
    i8h = kernels["half_transform"].get("int8_arm")
    if i8h:
        cands = [("i8h_convert_kernel (f64 rows -> residue planes, first J sweep fused)", i8h["convert"]["ms"], "hbm"),
                 ("i8h_gemm_kernel (tcgen05.mma.kind::i8, cluster-multicast, TMEM accumulators)", i8h["gemm"]["ms"], "tensor")]
        i8k = kernels["k_gemm"].get("int8_arm")
        if i8k:
            cands.append(("i8_pipeline_kernel (K GEMM, tcgen05.mma.kind::i8)", i8k["gemm"]["ms"], "tensor_k"))
        name, ms, kind = max(cands, key=lambda c: c[1])
        pk8 = kernels["int8_peak"]
        if kind == "hbm":
            roofline = {"kernel": name, "bound": "hbm", "achieved": i8h["convert"]["gbs"], "peak": hbm_peak, "unit": "GB/s",
                        "frac": i8h["convert"]["frac_of_hbm_peak"], "traffic": None, "peak_source": peak_src,
                        "algorithmic_bytes": i8h["convert"]["bytes"], "ms": ms}
        elif kind == "tensor" and (i8h["gemm"]["frac_of_hbm_peak"] or 0) >= (i8h["gemm"]["frac_of_int8_peak"] or 0):
            # the K3 residue GEMM does nocc multiply-adds per plane byte: it is a stream of the planes, nearer its HBM roof
            # than its tensor roof -- report the roof that binds, keep the other beside it
            g = i8h["gemm"]
            roofline = {"kernel": name, "bound": "hbm", "achieved": g["gbs"], "peak": hbm_peak, "unit": "GB/s", "frac": g["frac_of_hbm_peak"],
                        "traffic": None, "peak_source": peak_src, "algorithmic_bytes": g["bytes"], "ms": ms,
                        "tensor_view": {"int8_tops": g["int8_tops"], "peak": pk8["sustained_tops"], "frac": g["frac_of_int8_peak"],
                                        "peak_source": pk8["source"]}}
        else:
            g = i8h["gemm"] if kind == "tensor" else i8k["gemm"]
            roofline = {"kernel": name, "bound": "tensor", "achieved": g["int8_tops"], "peak": pk8["sustained_tops"], "unit": "TOP/s (int8)",
                        "frac": g["frac_of_int8_peak"], "traffic": None, "peak_source": pk8["source"], "ms": ms}
        roofline["share_of_build"] = ms / res["value"] if res["value"] else None
        roofline["note"] = ("dominant kernel of the default (INT8-tensor-core) build; the FP64 tensor-pipe roofline the north_star "
                            "names is `roofline_fp64`")

    cfg = res["cfg"]
    if world == 1 and not args.no_extra:
        try:
            cal = cublas_dgemm_calibration(cfg["nbf"], cfg["naux"] * (cfg["nocc"] + cfg["nocc"] % 2))
            kernels["k_gemm"]["cublas_dgemm_tflops"] = cal["k_gemm_shape"]["tflops"]
            kernels["k_gemm"]["cublas_dgemm"] = cal
            kernels["k_gemm"]["note"] = ("cuBLAS computes the full square (2 N^2 k flop); kgemm_ws_kernel is credited "
                                         "N(N+1) k for the upper-triangle tiles it executes")
        except Exception as ex:  # calibration must never cost the bench line
            kernels["k_gemm"]["cublas_dgemm"] = {"error": str(ex)[:200]}

    cb = None
    if world == 1 and not args.no_cpu_baseline:
        cb, _ = cpu_baseline(cfg, res["keep"], res["amp"], res["Cl"][0], res["Crl"], args.cpu_slice)

    wl = {}
    for nm, r in extra.items():
        kt, _ = kernel_table(r, peak_dmma, hbm_peak, peak_src)
        wl[nm] = {"config": config_of(nm, r["cfg"], r["keep"], r["Crl"]), "value_ms": r["value"], "e2e_ms": r["e2e_ms"],
                  "kernels": kt, "parity_spot": r["spot"], "arms_bit_identical": r["arms_equal"], "arms_diff": r["arms_diff"],
                  "run_to_run_bit_identical": r["run_equal"], "ranks_bit_identical": r["ranks_equal"],
                  "gpu_launches": int(r["launches"]), "hbm_tensor_gb": r["st_dev"]["hbm_tensor_bytes"] / 1e9}

    line = {
        "metric": METRIC, "value": res["value"], "unit": "ms", "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": res["value"], "higher_is_better": False, "scaling": "strong", "vs_baseline": None, "dtype": "f64",
        "dtype_note": ("FP64 operands and FP64 results.  On the default arms the two GEMMs of the K build are computed exactly in "
                       "the integers on the INT8 tensor cores (residues + CRT; the only rounding is the scaling of the operands, "
                       "2^-50.7 / 2^-46.9 of their row norms: the error class of a DGEMM in double, parity_spot and the 1e-10 "
                       "gates hold); `fp64_arms` / `value_fp64_arms_ms` is the same build on the FP64 tensor pipe (DMMA), "
                       "measured in this run") if (res["st_dev"]["kgemm_kind"] or res["st_dev"]["half_kind"]) else
                      "FP64 operands, FP64 tensor pipe (DMMA), FP64 results",
        "value_fp64_arms_ms": ab["ms_total"] if ab else (res["value"] if not (res["st_dev"]["kgemm_kind"] or res["st_dev"]["half_kind"]) else None),
        "data": "synthetic", "config": config_of(args.workload, cfg, res["keep"], res["Crl"]),
        "e2e": {"value": res["e2e_ms"], "unit": "ms", "h2d_bytes_per_step": res["h2d"], "d2h_bytes_per_step": res["d2h"],
                "ms_h2d": res["e2e_parts"]["ms_h2d"], "ms_d2h": res["e2e_parts"]["ms_d2h"]},
        "gpu_launches": int(res["launches"]), "roofline": roofline, "roofline_fp64": roofline_fp64, "cpu_baseline": cb, "kernels": kernels, "clocks": res["clk"],
        "wall_ms_per_step": res["wall_ms"], "arms_bit_identical": res["arms_equal"], "arms_diff": res["arms_diff"],
        "run_to_run_bit_identical": res["run_equal"], "ranks_bit_identical": res["ranks_equal"],
        "parity_spot": res["spot"], "workloads": wl,
        "hbm": {"tensor_gb": res["st_dev"]["hbm_tensor_bytes"] / 1e9, "work_gb": res["st_dev"]["hbm_work_bytes"] / 1e9,
                "fill_s": res["fill_s"], "layout_s": res["layout_s"]},
        "setup": res["setup"], "fp64_arms": res["fp64_arms"],
    }
    emit(line)
    ds.close()
    # A result that misses the oracle is fatal (the run exits non-zero); a determinism violation is
    # recorded in the JSON line (arms_bit_identical / arms_diff / ...) and reported on stderr, the line stays valid.
    bad, warn = [], []
    for nm, r in [(args.workload, res)] + list(extra.items()):
        if r["spot"] is not None and not r["spot"]["ok"]:
            bad.append(f"{nm}: spot parity {r['spot']['max_scaled_J']:.2e} / {r['spot']['max_scaled_K']:.2e} above {SPOT_TOL}")
        if not (r["arms_equal"] and r["run_equal"] and r["ranks_equal"]):
            warn.append(f"{nm}: results not bit-identical (arms {r['arms_equal']} {r['arms_diff']}, run to run {r['run_equal']}, "
                        f"ranks {r['ranks_equal']})")
    if warn:
        print("bench.py: determinism check: " + "; ".join(warn), file=sys.stderr)
    if bad:
        print("bench.py: FAILED checks: " + "; ".join(bad), file=sys.stderr)
        return 2
    return 0


if __name__ == "__main__":
    sys.exit(main())
