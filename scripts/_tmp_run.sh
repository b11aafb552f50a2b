P='import json,sys
for l in sys.stdin:
    if l.startswith("{"):
        d=json.loads(l); print(d["config"]["layout"][:14], d["dtype"][:4], d["config"]["rig"], "B", d["config"]["batch_per_gpu"], round(d["value"],1), d["kernel_ms_median"], round(d["roofline_fwd"]["frac"],3), round(d["roofline_bwd"]["frac"],3))'
B="python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-conv1 --no-sharded-conv1 --no-e2e --layout nchw"
$B 2>/dev/null | python -c "$P"
$B --dtype bf16 2>/dev/null | python -c "$P"
$B --batch 4 2>/dev/null | python -c "$P"
$B --rig multiviewx 2>/dev/null | python -c "$P"
echo old kernels; MVWARP_LIB=mvdet_b200/libmvwarp_dev.so MVWARP_NCHW_VARIANT=1 $B 2>/dev/null | python -c "$P"
MVWARP_LIB=mvdet_b200/libmvwarp_dev.so MVWARP_NCHW_VARIANT=1 $B --dtype bf16 2>/dev/null | python -c "$P"
