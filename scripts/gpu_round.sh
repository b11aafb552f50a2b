#!/bin/bash
# One gpurun call that regenerates the round's evidence on a B200 box (everything lands in gpurun_out/, which gpurun merges back):
#   gpurun --timeout 900 -- 'bash scripts/gpu_round.sh r02'
# then, back in the dev container (no GPU needed):
#   python scripts/ncu_summary.py gpurun_out/prof_r02.ncu-rep > profiles/r02_ncu_full.txt     # per-kernel summary, DRAM bytes -> profiles/traffic.json
#   cp gpurun_out/launches_r02.csv profiles/r02_ncu_launches.csv ; cp gpurun_out/r02_bench*.json profiles/
#   bash scripts/sass_listing.sh profiles                                                        # cuobjdump works without a GPU
# Order follows the driver's round end: tests, smoke, reference arm, our arm -- and only then the two ncu passes
# (a number printed under ncu is never a bench value; each ncu pass runs only after its command exited 0 without ncu).
TAG=${1:-cur}
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q > gpurun_out/${TAG}_pytest.log 2>&1; echo "pytest rc=$?" | tee gpurun_out/${TAG}_rc.txt
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/${TAG}_smoke.log 2>&1; echo "smoke rc=$?" | tee -a gpurun_out/${TAG}_rc.txt
python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/${TAG}_bench_reference_arm.json 2> gpurun_out/${TAG}_bench_ref.err
echo "reference rc=$?" | tee -a gpurun_out/${TAG}_rc.txt
python bench.py > gpurun_out/${TAG}_bench.json 2> gpurun_out/${TAG}_bench.err; echo "bench rc=$?" | tee -a gpurun_out/${TAG}_rc.txt
bash scripts/ncu_capture.sh ${TAG}
tail -3 gpurun_out/${TAG}_pytest.log
