#!/bin/bash
# Round-2 evidence in one gpurun call (everything lands in gpurun_out/, which gpurun merges back):
#   gpurun --timeout 2400 -- 'bash scripts/gpu_round2.sh r02'
# Order: tests, smoke, reference arm, our arm (never under a profiler), sweeps, then the ncu passes (each only after its command
# exited 0 without ncu).  Back in the dev container: scripts/ncu_summary.py on the .ncu-rep files -> profiles/.
TAG=${1:-r02}
O=gpurun_out
mkdir -p $O
python -m pytest tests -m gpu -q > $O/${TAG}_pytest.log 2>&1; echo "pytest rc=$?" | tee $O/${TAG}_rc.txt
python -c "import __graft_entry__ as g; g.smoke()" > $O/${TAG}_smoke.log 2>&1; echo "smoke rc=$?" | tee -a $O/${TAG}_rc.txt
python bench.py --impl reference --steps 5 --warmup 1 > $O/${TAG}_bench_reference_arm.json 2> $O/${TAG}_bench_ref.err; echo "reference rc=$?" | tee -a $O/${TAG}_rc.txt
python bench.py > $O/${TAG}_bench.json 2> $O/${TAG}_bench.err; echo "bench rc=$?" | tee -a $O/${TAG}_rc.txt
B="python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-conv1 --no-sharded-conv1"
$B --no-e2e --layout nchw > $O/${TAG}_bench_nchw.json 2>/dev/null
$B --no-e2e --fuse-upsample > $O/${TAG}_bench_upsample_cl.json 2>/dev/null
$B --no-e2e --fuse-upsample --layout nchw > $O/${TAG}_bench_upsample_nchw.json 2>/dev/null
$B --no-e2e --dtype bf16 --batch 8 > $O/${TAG}_bench_bf16_b8.json 2>/dev/null
$B --e2e-dense > $O/${TAG}_bench_e2e_dense.json 2>/dev/null
bash scripts/batch_sweep.sh > $O/${TAG}_batch_sweep.md 2> $O/${TAG}_batch_sweep.err
python scripts/train_step_bench.py 8 5 > $O/${TAG}_train_step_b8.json 2>&1
python scripts/train_step_bench.py 1 10 > $O/${TAG}_train_step_b1.json 2>&1
python scripts/train_step_breakdown.py 8 > $O/${TAG}_train_step_breakdown.json 2>&1
python scripts/map_head_bench.py > $O/${TAG}_map_head_bench.txt 2>&1
python scripts/detector_head_bench.py > $O/${TAG}_detector_head_bench.txt 2>&1
# ---- ncu: launch list of the default command, then full captures of the hot kernels per configuration -------------------
CMD="python bench.py --steps 3 --warmup 3 --no-e2e --no-cpu-baseline --no-conv1 --no-sharded-conv1"
$CMD > $O/plain_$TAG.log 2>&1 && ncu --metrics gpu__time_duration.sum --clock-control none -c 40 --csv --log-file $O/${TAG}_ncu_launches.csv $CMD > $O/ncu1.log 2>&1
cap() {  # name, kernel regex, skip, count, bench args...
  local name=$1 rx=$2 skip=$3 cnt=$4; shift 4
  $CMD "$@" > $O/plain_${TAG}_$name.log 2>&1 && \
  ncu --set full --clock-control none --import-source on -k regex:"$rx" -s $skip -c $cnt -f -o $O/${TAG}_prof_$name $CMD "$@" > $O/ncu_$name.log 2>&1
  tail -1 $O/ncu_$name.log
  # summarise on the box (gpurun_out/ may carry at most 64 MiB back): per-kernel text summary, DRAM bytes into a traffic json;
  # only the default configuration's report itself travels
  python scripts/ncu_summary.py $O/${TAG}_prof_$name.ncu-rep --traffic $O/${TAG}_traffic_$name.json > $O/${TAG}_ncu_full_$name.txt 2>> $O/ncu_$name.log
  [ "$name" = default ] || rm -f $O/${TAG}_prof_$name.ncu-rep
}
cap default 'fwd_|bwd_' 6 2
cap bf16_b8 'fwd_|bwd_' 6 2 --dtype bf16 --batch 8
cap nchw 'fwd_|bwd_' 6 2 --layout nchw
cap upsample 'fwd_|bwd_' 6 2 --fuse-upsample
cap upsample_nchw 'fwd_|bwd_|transpose' 12 4 --fuse-upsample --layout nchw
TS="python scripts/train_step_bench.py 1 2"
$TS > $O/plain_${TAG}_train.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:'warp_conv1|conv3x3_tc2|relu_bwd|conv1_tail' -s 12 -c 10 -f -o $O/${TAG}_prof_train $TS > $O/ncu_train.log 2>&1
python scripts/ncu_summary.py $O/${TAG}_prof_train.ncu-rep > $O/${TAG}_ncu_full_train.txt 2>> $O/ncu_train.log; rm -f $O/${TAG}_prof_train.ncu-rep
du -sh $O
tail -3 $O/${TAG}_pytest.log; cat $O/${TAG}_rc.txt
