#!/bin/bash
# A/B matrix of the K GEMM switches: triangular diagonal tiles on/off x split-K reduction folded / separate.
# usage: tools/ab_kgemm.sh <workload> [steps]   -> one line per arm with ms of the three kernels
wl=${1:-c60_tz_q8}; steps=${2:-5}
for tri in 1 0; do for red in fold separate; do
  B200JK_KTRI=$tri B200JK_KREDUCE=$red python bench.py --workload $wl --steps $steps --warmup 3 --no-extra --no-spot --no-cpu-baseline --skip-probes 2>/dev/null \
   | python -c "import sys,json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); k=d['kernels']; print('$wl tri=$tri reduce=$red value %.3f half %.3f kgemm %.3f j %.3f' % (d['value'], k['half_transform']['ms'], k['k_gemm']['ms'], k['j_sweeps']['ms']))"
done; done
