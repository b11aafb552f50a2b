#!/bin/bash
# ncu capture of the fused warp->conv1 kernel (tensor-pipe utilisation, north_star (d))
CMD="python scripts/conv1_bench.py 1 2"
$CMD > gpurun_out/conv1_plain.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:'warp_conv1_pipe' -s 1 -c 1 -f -o gpurun_out/prof_conv1 $CMD > gpurun_out/ncu_conv1.log 2>&1
tail -2 gpurun_out/ncu_conv1.log; cat gpurun_out/conv1_plain.log | tail -1
