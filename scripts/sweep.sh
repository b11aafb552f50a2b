#!/bin/bash
# development: layout x dtype x batch x rig matrix on one GPU (kernel times, roofline fractions)
B="python bench.py --steps 10 --warmup 3 --no-e2e --no-cpu-baseline"
run() { echo -n "$* :: "; $B "$@" | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('value %.1f frames/s'%d['value'], 'fwd %.3f ms (%.3f)'%(d['kernel_ms']['fwd'], d['roofline_fwd']['frac']), 'bwd %.3f ms (%.3f)'%(d['kernel_ms']['bwd'], d['roofline_bwd']['frac']))"; }
for a in "$@"; do run $a; done
