#!/bin/bash
# Last check of round 2 on one GPU, in the driver's order: tests, smoke, our bench arm; then ncu evidence of the kernels added late in the
# round (fused up-sample conv1, two-pass head).  The warp kernels' sources are unchanged since profiles/traffic.json was captured
# (bench.py checks the hash), so the default capture is not repeated.   usage: gpurun --timeout 900 -- 'bash scripts/gpu_round2_final2.sh r02g'
TAG=${1:-r02g}
O=gpurun_out; mkdir -p $O
python -m pytest tests -m gpu -q > $O/${TAG}_pytest.log 2>&1; echo "pytest rc=$?" | tee $O/${TAG}_rc.txt
python -c "import __graft_entry__ as g; g.smoke()" > $O/${TAG}_smoke.log 2>&1; echo "smoke rc=$?" | tee -a $O/${TAG}_rc.txt
python bench.py > $O/${TAG}_bench.json 2> $O/${TAG}_bench.err; echo "bench rc=$?" | tee -a $O/${TAG}_rc.txt
CMD="python bench.py --steps 3 --warmup 3 --no-e2e --no-cpu-baseline --no-conv1 --no-sharded-conv1"
$CMD > $O/plain_$TAG.log 2>&1 && ncu --metrics gpu__time_duration.sum --clock-control none -c 40 --csv --log-file $O/${TAG}_ncu_launches.csv $CMD > $O/ncu1.log 2>&1
UP="python scripts/ncu_map_head_up.py 1 2"
$UP > $O/${TAG}_map_head_up_plain.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:'warp_conv1_direct|conv3x3_tc2|head_partial|head_stencil|tail_reduce' -s 10 -c 5 -f -o $O/${TAG}_prof_map_head_up $UP > $O/ncu_up.log 2>&1
python scripts/ncu_summary.py $O/${TAG}_prof_map_head_up.ncu-rep > $O/${TAG}_ncu_full_map_head_up.txt 2>> $O/ncu_up.log; rm -f $O/${TAG}_prof_map_head_up.ncu-rep
python scripts/ncu_map_head_up.py 8 3 >> $O/${TAG}_map_head_up_plain.log 2>&1
tail -3 $O/${TAG}_pytest.log; cat $O/${TAG}_rc.txt; cat $O/${TAG}_map_head_up_plain.log; tail -c 600 $O/${TAG}_bench.json
