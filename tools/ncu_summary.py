"""Summarise an .ncu-rep (read HERE with `ncu -i ... --page raw --csv`) into the JSON kept under profiles/:
per captured kernel the duration, DRAM bytes, tensor/DMMA pipe activity, occupancy facts and the top stall reasons.
Usage: python tools/ncu_summary.py gpurun_out/prof.ncu-rep profiles/rNN_ncu_xxx.json "command that was profiled" """
import csv
import io
import json
import subprocess
import sys

KEEP = [
    "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
    "sm__throughput.avg.pct_of_peak_sustained_elapsed", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
    "sm__pipe_tensor_subpipe_dmma_cycles_active.avg.pct_of_peak_sustained_active",
    "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed",
    "sm__pipe_tensor_cycles_active.min.pct_of_peak_sustained_elapsed",
    "sm__pipe_tensor_cycles_active.max.pct_of_peak_sustained_elapsed",
    "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread", "launch__grid_size",
    "launch__block_size", "sm__cycles_elapsed.avg.per_second", "lts__t_sector_hit_rate.pct",
    "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "smsp__issue_active.avg.pct_of_peak_sustained_active",
    "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio",
]


def main():
    rep, out, cmd = sys.argv[1], sys.argv[2], (sys.argv[3] if len(sys.argv) > 3 else "")
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True, check=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    head, units, data = rows[0], rows[1], rows[2:]
    col = {name: i for i, name in enumerate(head)}
    kernels = {}
    for r in data:
        name = r[col["Kernel Name"]]
        kernels[name] = {k: f"{r[col[k]]} {units[col[k]]}".strip() for k in KEEP if k in col}
    json.dump({"command": cmd, "kernels": kernels}, open(out, "w"), indent=1)
    for k, v in kernels.items():
        print(k[:50], v.get("gpu__time_duration.sum"), "dmma", v.get("sm__pipe_tensor_subpipe_dmma_cycles_active.avg.pct_of_peak_sustained_active"),
              "dram", v.get("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed"))


if __name__ == "__main__":
    main()
