"""CPU checks of bench.py's JSON contract: the reference arm (the reference's object code on the host cores; the one place
besides the checker where bench.py executes oracle/) and the kernel / roofline tables of the INT8 arms, fed with the
stats of a recorded run."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def test_reference_arm_prints_the_contract_line():
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "1", "--warmup", "0",
                        "--cpu-slice", "16"], capture_output=True, text=True, timeout=600, cwd=ROOT)
    assert r.returncode == 0, r.stderr[-1500:]
    lines = [x for x in r.stdout.splitlines() if x.strip()]
    assert len(lines) == 1, lines  # stdout carries exactly one JSON line
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["metric"] == "DF-JK ms/SCF-iter at C60/cc-pVTZ" and d["unit"] == "ms"
    assert d["higher_is_better"] is False and d["dtype"] == "f64" and d["gpu_launches"] == 0
    assert d["cpu_baseline"]["kind"] in ("reference", "port") and d["cpu_baseline"]["cores"] >= 1
    assert d["cpu_baseline"]["value"] == d["value"] == d["e2e"]["value"] > 0
    assert d["e2e"]["h2d_bytes_per_step"] == 0 and d["e2e"]["d2h_bytes_per_step"] == 0
    assert "c60_tz" in d["config"]["workload"] and "nbf=1800 naux=4740 nocc=180" in d["config"]["workload"]


def _res(nq, sub_h, sub_k, plane_bytes, convert_bytes, half_ops, k_ops, value):
    st = dict(half_kind=1, kgemm_kind=1, half_moduli=12, half_i8_chunks=11, half_i8_cached=0, half_i8_resident_rows=594,
              half_i8_convert_bytes=convert_bytes, half_i8_ops=half_ops, half_i8_plane_bytes=plane_bytes, kgemm_i8_ops=k_ops,
              kgemm_moduli=13, q_begin=0, q_end=nq, half_flops=2.0 * nq * 2316568 * 180, half_bytes=8.0 * nq * 2316568,
              kgemm_flops=1800.0 * 1801 * nq * 180, j_bytes=8.0 * nq * 1159184)
    return dict(st_dev=st, sub={"ms_half_i8": sub_h, "ms_kgemm_i8": sub_k}, cfg=dict(nbf=1800, nocc=180, nmat=1), Crl=None,
                parts={"ms_j": 7.9, "ms_half": sum(sub_h), "ms_kgemm": sum(sub_k), "ms_allreduce": 0.0}, reduce_kind="none", value=value)


def test_int8_tables_follow_the_stats():
    """Numbers of profiles/r02_bench_c60_n1_final2.json: the residue GEMM of the half transform is reported against HBM with
    planes + gathered C^T + residue bytes as its traffic (what ncu counted on the 592-row shard), the tensor view beside it."""
    import bench

    res = _res(4740, [28.083, 2.49, 30.114, 11.743], [7.002, 17.753, 0.788], 141350400000.0, 153440268864.0, 5.428e13, 3.885e13, 106.1)
    kt, half_tf = bench.kernel_table(res, 37.04, 6453.4, "measured")
    h = kt["half_transform"]["int8_arm"]
    assert h["resident_row_blocks"] == 594 and h["chunks"] == 11 and h["moduli"] == 12
    assert abs(h["convert"]["gbs"] - 153440268864.0 / 28.083e-3 / 1e9) < 1e-6
    g = h["gemm"]
    # 12 moduli x 1800 x 4740 x 192 residue bytes + the gathered C^T (planes / (38 q-tiles x 128 rows) x 192 columns)
    want = 141350400000.0 + 141350400000.0 / (38 * 128) * 192 + 12.0 * 1800 * 4740 * 192
    assert abs(g["bytes"] - want) < 1.0
    assert 0.84 < g["frac_of_hbm_peak"] < 0.88 and 0.62 < g["frac_of_int8_peak"] < 0.66
    assert kt["int8_peak"]["sustained_tops"] > 0
    k = kt["k_gemm"]["int8_arm"]
    assert abs(k["gemm"]["int8_tops"] - 3.885e13 / 17.753e-3 / 1e12) < 1e-6
    assert half_tf > 37.04  # FP64-equivalent rate of the residue arm is above the FP64 pipe's ceiling


MAIN_SCRIPT = r"""
import json, os, sys
sys.path.insert(0, os.environ["B2_ROOT"])
import bench

class FakeDist:
    rank, world, local_rank = 0, 1, 0
    def nccl_id(self, E): return None
    def barrier(self): pass
    def max(self, x): return x
    def all_true(self, f): return f
    def close(self): pass

def fake_measure(args, ds, name, steps, warmup, headline):
    cfg, keep, amp, Cl, Crl, D = bench.make_inputs(args, name)
    nq = cfg["naux"]
    st = dict(half_kind=1, kgemm_kind=1, half_moduli=12, half_i8_chunks=11, half_i8_cached=0, half_i8_resident_rows=594,
              half_i8_convert_bytes=153440268864.0, half_i8_ops=5.428e13, half_i8_plane_bytes=141350400000.0, kgemm_i8_ops=3.885e13,
              kgemm_moduli=13, q_begin=0, q_end=nq, half_flops=2.0 * nq * 2316568 * 180, half_bytes=8.0 * nq * 2316568,
              kgemm_flops=1800.0 * 1801 * nq * 180, j_bytes=8.0 * nq * 1159184, reduce_kind=0, hbm_tensor_bytes=87.9e9, hbm_work_bytes=97e9)
    pk = {"sustained_tflops": 37.04, "burst_tflops": 37.1}
    ab = {"ms_total": 205.8, "ms_j": 7.6, "ms_half": 116.7, "ms_kgemm": 81.3, "max_abs_dK_vs_default_arms": 9e-14}
    return dict(cfg=cfg, keep=keep, amp=amp, Cl=Cl, Crl=Crl, value=106.1, wall_ms=106.3,
                parts={"ms_j": 7.9, "ms_half": 72.5, "ms_kgemm": 25.5, "ms_allreduce": 0.0},
                sub={"ms_half_i8": [28.083, 2.49, 30.114, 11.743], "ms_kgemm_i8": [7.002, 17.753, 0.788]}, st_dev=st, e2e_ms=107.25,
                e2e_parts={"ms_h2d": 0.6, "ms_d2h": 1.0}, launches=580, clk={"sm_mhz": 1571.0, "sm_max_mhz": 1965.0, "reasons": ["sw_power_cap"]},
                arms_equal=True, arms_diff={}, run_equal=True, ranks_equal=True,
                spot={"ok": True, "max_scaled_J": 4.2e-14, "max_scaled_K": 2.6e-14, "seconds": 1.8}, pk_dmma=pk, pk_dfma=pk, layout_s=0.02,
                fill_s=0.07, reduce_kind="none (one GPU)", setup=None, fp64_arms=ab, h2d=28512000, d2h=51840000)

bench.Dist = FakeDist
bench.measure = fake_measure
bench.cublas_dgemm_calibration = lambda n, k: {"k_gemm_shape": {"tflops": 33.7}}
bench.cpu_baseline = lambda *a, **k: ({"value": 56648.0, "unit": "ms", "cores": 16, "kind": "reference", "sample": "fake"}, [])
sys.argv = ["bench.py", "--no-extra"]
sys.exit(bench.main())
"""


def test_main_assembles_the_line_from_a_recorded_measurement(tmp_path):
    """bench.main() with the measurement replaced by the numbers of profiles/r02_bench_c60_n1_final2.json (no GPU here): the
    contract keys, the dominant-kernel roofline with the roof that binds it, the FP64 roofline beside it, the dtype note."""
    script = tmp_path / "main.py"
    script.write_text(MAIN_SCRIPT)
    r = subprocess.run([sys.executable, str(script)], env=dict(os.environ, B2_ROOT=ROOT), capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stdout[-800:] + r.stderr[-1500:]
    lines = [x for x in r.stdout.splitlines() if x.strip()]
    assert len(lines) == 1
    d = json.loads(lines[0])
    for k in ("metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling", "vs_baseline", "dtype",
              "data", "config", "e2e", "gpu_launches", "roofline", "cpu_baseline", "clocks"):
        assert k in d, k
    assert d["value"] == 106.1 and d["e2e"]["value"] == 107.25 and d["e2e"]["h2d_bytes_per_step"] > 0 and d["gpu_launches"] == 580
    rf = d["roofline"]
    assert rf["bound"] == "hbm" and rf["unit"] == "GB/s" and "i8h_gemm_kernel" in rf["kernel"]
    assert abs(rf["frac"] - rf["achieved"] / rf["peak"]) < 1e-12 and 0.8 < rf["frac"] < 0.9
    assert 0.6 < rf["tensor_view"]["frac"] < 0.7
    f64 = d["roofline_fp64"]
    assert f64["bound"] == "tensor" and f64["unit"] == "TFLOP/s" and abs(f64["frac"] - f64["achieved"] / f64["peak"]) < 1e-12
    assert 0.9 < f64["frac"] < 0.93  # K3 on the DMMA pipe, from the fp64_arms builds
    assert d["value_fp64_arms_ms"] == 205.8 and "INT8 tensor cores" in d["dtype_note"] and d["dtype"] == "f64"
    assert d["cpu_baseline"]["kind"] == "reference" and d["kernels"]["half_transform"]["int8_arm"]["resident_row_blocks"] == 594
