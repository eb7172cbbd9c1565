#!/bin/bash
# compute-sanitizer passes over tools/sanitize_case.py: memcheck, initcheck, and racecheck on the
# all-TMA half transform (B200JK_CGATHER=1) and on the cp.async gather path (B200JK_CGATHER=0).  Summaries -> gpurun_out/.
out=${1:-gpurun_out}; mkdir -p $out
CS=/usr/local/cuda/bin/compute-sanitizer
run() { # name, tool args..., env
  name=$1; shift
  echo "== $name"; ( timeout 600 env "$@" ) > $out/r02_san_$name.log 2>&1
  grep -E "ERROR SUMMARY|RACECHECK SUMMARY|sanitize_case ok|hazards" $out/r02_san_$name.log | tail -3
}
run memcheck         B200JK_CGATHER=0 $CS --tool memcheck --print-limit 20 python tools/sanitize_case.py
run memcheck_tma     B200JK_CGATHER=1 $CS --tool memcheck --print-limit 20 python tools/sanitize_case.py
run initcheck        B200JK_CGATHER=0 $CS --tool initcheck --print-limit 20 python tools/sanitize_case.py
run racecheck_tma    B200JK_CGATHER=1 SAN_NBF=100 SAN_NAUX=70 $CS --tool racecheck --racecheck-report analysis --print-limit 20 python tools/sanitize_case.py
run racecheck_gather B200JK_CGATHER=0 SAN_NBF=100 SAN_NAUX=70 $CS --tool racecheck --racecheck-report analysis --print-limit 20 python tools/sanitize_case.py
# the INT8-tensor-core arms forced on the same small case (tcgen05 / TMEM / bulk-copy kernels).  Run so far: the three tools over
# smoke() with both arms (profiles/r02_san_*_smoke_both_arms.log):
#   compute-sanitizer --tool memcheck|initcheck|racecheck python -c "import __graft_entry__ as g; g.smoke()"
run memcheck_i8      B200JK_HALF=i8 B200JK_KGEMM=i8 $CS --tool memcheck --print-limit 20 python tools/sanitize_case.py
run initcheck_i8     B200JK_HALF=i8 B200JK_KGEMM=i8 $CS --tool initcheck --print-limit 20 python tools/sanitize_case.py
