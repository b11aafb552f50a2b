#!/bin/bash
# Tuning knobs (MVWARP_*) exist in the development build only: exports MVWARP_LIB=mvdet_b200/libmvwarp_dev.so (python -m mvdet_b200.build --dev)
export MVWARP_LIB=${MVWARP_LIB:-$PWD/mvdet_b200/libmvwarp_dev.so}
B="python bench.py --steps 20 --warmup 5 --no-e2e --no-cpu-baseline"
for cfg in "16 15" "24 23" "32 31" "40 39" "48 47" "20 19" "28 27"; do
  set -- $cfg
  echo -n "BUDGET=$1 HEAVY=$2  "; MVWARP_TASK_BUDGET=$1 MVWARP_HEAVY_LEN=$2 $B | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(d['kernel_ms'], round(d['roofline_bwd']['frac'],3), d['plan'])"
done
