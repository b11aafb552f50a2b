#!/bin/bash
# Multi-GPU evidence (run with gpurun --gpus N): the driver-style bench line at N ranks (weak-scaling warp + the strong-scaling
# sharded conv1 leg), the multi-rank parity test, and -- at N = 8 -- BASELINE.json configs[4]'s 8-GPU batch sweep.
#   gpurun --gpus 8 --timeout 1500 -- 'bash scripts/gpu_round2_multi.sh r02 8'
TAG=${1:-r02}; N=${2:-8}
O=gpurun_out; mkdir -p $O
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1"
python -m pytest tests/test_gpu_multi.py -m gpu -q > $O/${TAG}_pytest_multi_${N}gpu.log 2>&1; echo "pytest multi rc=$?"; tail -2 $O/${TAG}_pytest_multi_${N}gpu.log
$TR --master-port 29521 bench.py --gpus $N --steps 20 --warmup 5 2> $O/${TAG}_bench_${N}gpu.err | grep '^{' > $O/${TAG}_bench_${N}gpu.json; echo "bench rc=$?"
python bench.py --impl reference --gpus $N --steps 2 --warmup 1 > /dev/null 2>&1; echo "ref arm rc=$?"
if [ "$N" = 8 ]; then
  line() { $TR --master-port 29522 bench.py --gpus $N --steps 10 --warmup 3 --no-e2e --no-conv1 --no-sharded-conv1 "$@" 2>/dev/null | grep '^{' | python -c "
import sys,json; d=json.loads(sys.stdin.read()); c=d['config']
f=d.get('roofline_fwd', d['roofline']); b=d.get('roofline_bwd')
print('| %s | %s | %d | %d | %s | %.1f | %.3f (%.0f %%) | %s |' % (c['rig'], d['dtype'][:4], d['n_gpus'], c['batch_per_gpu'], c['mode'], d['value'], f['ms'], 100*f['frac'], ('%.3f (%.0f %%)' % (b['ms'], 100*b['frac'])) if b else '-'))"; }
  { echo "| rig | dtype | GPUs | B per GPU | mode | frames/s (all GPUs) | fwd ms, slowest rank (of HBM peak) | bwd ms, slowest rank (of HBM peak) |"
    echo "|---|---|---|---|---|---|---|---|"
    for rig in wildtrack multiviewx; do
      for b in 1 4 16; do line --rig $rig --dtype fp32 --batch $b; done
      for b in 1 8 32; do line --rig $rig --dtype bf16 --batch $b; done
      line --rig $rig --dtype bf16 --batch 64 --mode fwd
    done; } > $O/${TAG}_batch_sweep_8gpu.md
  cat $O/${TAG}_batch_sweep_8gpu.md
fi
python - <<PY
import json
d=json.load(open("$O/${TAG}_bench_${N}gpu.json"))
print("value", d["value"], "e2e", d["e2e"]["value"], "kernel_ms", d["kernel_ms"])
print(json.dumps(d.get("sharded_conv1"), indent=1))
PY
