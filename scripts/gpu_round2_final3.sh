O=gpurun_out; TAG=r02i
python bench.py > $O/${TAG}_bench.json 2> $O/${TAG}_bench.err; echo "bench rc=$?"
CMD="python bench.py --steps 3 --warmup 3 --no-e2e --no-cpu-baseline --no-conv1 --no-sharded-conv1 --dtype bf16 --batch 8"
$CMD > $O/plain_${TAG}_bf16.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:'fwd_|bwd_' -s 6 -c 2 -f -o $O/${TAG}_prof_bf16_b8 $CMD > $O/ncu_bf16.log 2>&1
python scripts/ncu_summary.py $O/${TAG}_prof_bf16_b8.ncu-rep > $O/${TAG}_ncu_full_bf16_b8.txt 2>> $O/ncu_bf16.log; rm -f $O/${TAG}_prof_bf16_b8.ncu-rep
python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-conv1 --no-sharded-conv1 --no-e2e --dtype bf16 --batch 8 > $O/${TAG}_bench_bf16_b8.json 2>/dev/null
tail -c 300 $O/${TAG}_bench.json; grep -c . $O/${TAG}_ncu_full_bf16_b8.txt
