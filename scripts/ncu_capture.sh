#!/bin/bash
# launch list + one full ncu capture of the two hot kernels (run on the GPU box through gpurun)
TAG=${1:-cur}
CMD="python bench.py --steps 3 --warmup 3 --no-e2e --no-cpu-baseline"
$CMD > gpurun_out/plain.log 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none -c 40 --csv --log-file gpurun_out/launches_$TAG.csv $CMD > gpurun_out/ncu1.log 2>&1
$CMD > gpurun_out/plain2.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:'fwd_cl_kernel|bwd_cl_' -s 9 -c 3 -f -o gpurun_out/prof_$TAG $CMD > gpurun_out/ncu2.log 2>&1
tail -2 gpurun_out/ncu2.log
