#!/bin/bash
# development: backward timings across configurations (one line each)
B="python bench.py --steps 20 --warmup 5 --no-e2e --no-cpu-baseline --no-conv1 --no-sharded-conv1"
run() { echo -n "$* :: "; $B "$@" | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('value %.1f'%d['value'], 'fwd %.4f ms (%.3f)'%(d['kernel_ms_median']['fwd'], d['roofline_fwd']['frac']), 'bwd %.4f ms median, %.4f min (%.3f)'%(d['kernel_ms_median']['bwd'], d['kernel_ms_min']['bwd'], d['roofline_bwd']['frac']))"; }
run --dtype fp32
run --dtype fp32 --batch 4
run --dtype bf16 --batch 8
run --dtype fp32 --rig multiviewx
run --dtype bf16 --batch 8 --rig multiviewx
run --dtype fp32 --fuse-upsample
run --dtype fp32 --layout nchw
