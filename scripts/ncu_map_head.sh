#!/bin/bash
# ncu capture of the three map-head kernels + the tail reduce (tensor-pipe utilisation, DRAM traffic), Wildtrack B=1
CMD="python scripts/map_head_bench.py 1 2"
$CMD > gpurun_out/map_head_plain.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:"warp_conv1|conv3x3_tc|conv3x3_head|tail_reduce" -s 4 -c 4 -f -o gpurun_out/prof_map_head $CMD > gpurun_out/ncu_map_head.log 2>&1
tail -2 gpurun_out/ncu_map_head.log; tail -1 gpurun_out/map_head_plain.log
