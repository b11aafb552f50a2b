#!/bin/bash
# development: forward of fused up-sample plans -- the quad kernel's occupancy variants (MVWARP_TBL_VARIANT 8 = <UNROLL 2, 2 CTAs per SM>,
# 9 = <2,3>, 10 = <1,6>, 11 = <1,3>, 12 = <1,4>) against the shipped CSR tile kernel (0 / 2) and the CSR task kernel (1); needs the development library
export MVWARP_LIB=${MVWARP_LIB:-$PWD/mvdet_b200/libmvwarp_dev.so}
B="python bench.py --steps 20 --warmup 3 --no-e2e --no-cpu-baseline --no-conv1 --no-sharded-conv1 --fuse-upsample"
run() { echo -n "TBL_VARIANT=$1 ${*:2} :: "; MVWARP_TBL_VARIANT=$1 $B "${@:2}" | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('fwd %.4f ms (%.3f) min %.4f'%(d['kernel_ms']['fwd'], d['roofline_fwd']['frac'], d['kernel_ms_min']['fwd']), 'bwd %.4f ms (%.3f)'%(d['kernel_ms']['bwd'], d['roofline_bwd']['frac']))"; }
for v in ${FP32_VARIANTS:-0 8 9 10 11 12}; do run $v --dtype fp32; done
for v in ${BF16_VARIANTS:-0 8 10 12}; do run $v --dtype bf16; done
run 8 --dtype fp32 --batch 4
run 0 --dtype fp32 --batch 4
run 8 --dtype fp32 --rig multiviewx
run 0 --dtype fp32 --rig multiviewx
