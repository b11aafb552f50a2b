#!/bin/bash
# Final check of the round on one GPU, in the driver's order: tests, smoke, reference arm, our arm; then the launch list and ONE full
# ncu capture of the default command (-> profiles/traffic.json via scripts/make_traffic.py).  ~5 minutes.
TAG=${1:-r02f}
O=gpurun_out; mkdir -p $O
python -m pytest tests -m gpu -q > $O/${TAG}_pytest.log 2>&1; echo "pytest rc=$?" | tee $O/${TAG}_rc.txt
python -c "import __graft_entry__ as g; g.smoke()" > $O/${TAG}_smoke.log 2>&1; echo "smoke rc=$?" | tee -a $O/${TAG}_rc.txt
python bench.py --impl reference --steps 5 --warmup 1 > $O/${TAG}_bench_reference_arm.json 2> $O/${TAG}_bench_ref.err; echo "reference rc=$?" | tee -a $O/${TAG}_rc.txt
python bench.py > $O/${TAG}_bench.json 2> $O/${TAG}_bench.err; echo "bench rc=$?" | tee -a $O/${TAG}_rc.txt
CMD="python bench.py --steps 3 --warmup 3 --no-e2e --no-cpu-baseline --no-conv1 --no-sharded-conv1"
$CMD > $O/plain_$TAG.log 2>&1 && ncu --metrics gpu__time_duration.sum --clock-control none -c 40 --csv --log-file $O/${TAG}_ncu_launches.csv $CMD > $O/ncu1.log 2>&1
$CMD > $O/plain2_$TAG.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:'fwd_|bwd_' -s 6 -c 2 -f -o $O/${TAG}_prof_default $CMD > $O/ncu2.log 2>&1
python scripts/ncu_summary.py $O/${TAG}_prof_default.ncu-rep --traffic $O/${TAG}_traffic_default.json > $O/${TAG}_ncu_full_default.txt 2>> $O/ncu2.log
tail -3 $O/${TAG}_pytest.log; cat $O/${TAG}_rc.txt; tail -c 400 $O/${TAG}_bench.json
