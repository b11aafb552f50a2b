#!/bin/bash
# BASELINE.json configs[4]: batch sweep at the Wildtrack and MultiviewX shapes on one GPU (per-GPU numbers; with
# view sharding the N-GPU job is N independent copies of this).  Memory is capped near 100 GB per run: fp32
# fwd+bwd stops at B=16, bf16 fwd+bwd at B=32, B=64 runs forward-only in bf16 (hard part 9).
B="python bench.py --steps 10 --warmup 3 --no-e2e --no-cpu-baseline"
line() { $B "$@" | python -c "
import sys,json; d=json.loads(sys.stdin.read()); c=d['config']
f=d.get('roofline_fwd', d['roofline']); b=d.get('roofline_bwd')
print('| %s | %s | %d | %s | %.1f | %.3f (%.0f %%) | %s |' % (c['rig'], d['dtype'][:4], c['batch_per_gpu'], c['mode'], d['value'], f['ms'], 100*f['frac'], ('%.3f (%.0f %%)' % (b['ms'], 100*b['frac'])) if b else '-'))"; }
echo "| rig | dtype | B | mode | frames/s | fwd ms (of HBM peak) | bwd ms (of HBM peak) |"
echo "|---|---|---|---|---|---|---|"
for rig in wildtrack multiviewx; do
  for b in 1 2 4 8 16; do line --rig $rig --dtype fp32 --batch $b; done
  for b in 1 2 4 8 16 32; do line --rig $rig --dtype bf16 --batch $b; done
  line --rig $rig --dtype bf16 --batch 64 --mode fwd
done
