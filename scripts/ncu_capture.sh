#!/bin/bash
# launch list + one full ncu capture of the hot kernels (run on the GPU box through gpurun)
# usage: bash scripts/ncu_capture.sh TAG [extra bench args]
TAG=${1:-cur}; shift
CMD="python bench.py --steps 3 --warmup 3 --no-e2e --no-cpu-baseline $*"
$CMD > gpurun_out/plain_$TAG.log 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none -c 40 --csv --log-file gpurun_out/launches_$TAG.csv $CMD > gpurun_out/ncu1.log 2>&1
$CMD > gpurun_out/plain2_$TAG.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:'fwd_|bwd_' -s 6 -c 2 -f -o gpurun_out/prof_$TAG $CMD > gpurun_out/ncu2.log 2>&1
tail -2 gpurun_out/ncu2.log
