#!/bin/bash
# Tuning knobs (MVWARP_*) exist in the development build only: exports MVWARP_LIB=mvdet_b200/libmvwarp_dev.so (python -m mvdet_b200.build --dev)
export MVWARP_LIB=${MVWARP_LIB:-$PWD/mvdet_b200/libmvwarp_dev.so}
# development: sweep kernel tuning variants on one GPU; prints kernel times per variant
B="python bench.py --steps 20 --warmup 5 --no-e2e --no-cpu-baseline"
for bv in ${BWD_VARIANTS:-0}; do
  echo -n "BWD_VARIANT=$bv  "; MVWARP_BWD_VARIANT=$bv $B $EXTRA | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(d['kernel_ms'], round(d['roofline_bwd']['frac'],3), d['value'])"
done
for fv in ${FWD_VARIANTS:-}; do
  echo -n "FWD_VARIANT=$fv  "; MVWARP_FWD_VARIANT=$fv $B $EXTRA --mode fwd | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(d['kernel_ms'], round(d['roofline']['frac'],3))"
done
