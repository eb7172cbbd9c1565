#!/bin/bash
# One gpurun call that produces the round's numbers of record (everything lands in gpurun_out/, copied to profiles/ by hand):
#   tools/record_r02.sh [tag]
#   1. pytest -m gpu ($TESTS: default the whole suite)   -> pytest_gpu_<tag>.log
#   2. bench.py (default: C60, N = 1)                    -> bench_c60_n1_<tag>.json
#   3. bench.py on the 8-GPU shard of C60 (592 rows)     -> bench_c60q8_n1_<tag>.json   (residue planes resident across builds)
#   4. ncu launch list of the bench command              -> launches_c60_<tag>.csv      (gpu__time_duration only: no replay)
#   5. ncu --set full of the INT8 arms on the shard      -> ncu_i8_q8_<tag>.ncu-rep     ($NCU_K kernel regex, $NCU_C launches)
# SKIP_TESTS / SKIP_LIST / SKIP_NCU_FULL=1 leave a step out; LIST_STEPS / LIST_WARMUP shorten the launch list.
tag=${1:-final}
out=gpurun_out
mkdir -p $out
t0=$(date +%s)
stamp() { echo "[$(( $(date +%s) - t0 )) s] $*" | tee -a $out/record_$tag.log; }
stamp start; nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,memory.used --format=csv,noheader | tee -a $out/record_$tag.log
if [ -z "$SKIP_TESTS" ]; then
  timeout 420 python -m pytest ${TESTS:-tests} -m gpu -x -q --durations=8 > $out/pytest_gpu_$tag.log 2>&1
  stamp "pytest rc=$? $(tail -1 $out/pytest_gpu_$tag.log)"
fi
timeout 400 python bench.py > $out/bench_c60_n1_$tag.json 2> $out/bench_c60_n1_$tag.err
stamp "bench rc=$?"; python tools/show_bench.py $out/bench_c60_n1_$tag.json 2>&1 | tee -a $out/record_$tag.log
timeout 200 python bench.py --workload c60_tz_q8 --no-extra --no-cpu-baseline > $out/bench_c60q8_n1_$tag.json 2> $out/bench_c60q8_n1_$tag.err
stamp "bench q8 rc=$?"; python tools/show_bench.py $out/bench_c60q8_n1_$tag.json 2>&1 | tee -a $out/record_$tag.log
if [ -z "$SKIP_LIST" ]; then
  timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 900 --csv --log-file $out/launches_c60_$tag.csv \
    python bench.py --steps ${LIST_STEPS:-2} --warmup ${LIST_WARMUP:-3} --no-extra --no-cpu-baseline --no-spot --skip-probes --no-ab > $out/launches_bench_$tag.json 2> $out/launches_bench_$tag.err
  stamp "launch list rc=$? lines=$(wc -l < $out/launches_c60_$tag.csv)"
fi
if [ -z "$SKIP_NCU_FULL" ]; then
  timeout 300 ncu --set full --clock-control none --import-source on -k regex:"${NCU_K:-i8}" -c ${NCU_C:-14} -f -o $out/ncu_i8_q8_$tag \
    python bench.py --workload c60_tz_q8 --steps 1 --warmup 3 --no-extra --no-cpu-baseline --no-spot --skip-probes --no-ab > $out/ncu_i8_q8_$tag.json 2> $out/ncu_i8_q8_$tag.err
  stamp "ncu full rc=$? $(ls -la $out/ncu_i8_q8_$tag.ncu-rep 2>/dev/null | awk '{print $5}') bytes"
fi
stamp done
