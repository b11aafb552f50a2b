P='import json,sys
for l in sys.stdin:
    if l.startswith("{"):
        d=json.loads(l); print(d["config"]["layout"][:14], d["dtype"][:4], "B", d["config"]["batch_per_gpu"], round(d["value"],1), d["kernel_ms_median"], round(d["roofline_fwd"]["frac"],3), round(d["roofline_bwd"]["frac"],3))'
python -m pytest tests/test_gpu_parity.py -m gpu -q -k "upsampl or staged" 2>&1 | tail -3
B="python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-conv1 --no-sharded-conv1 --no-e2e --fuse-upsample"
for v in 0 4 5 6 7; do echo "tbl variant $v"; MVWARP_LIB=mvdet_b200/libmvwarp_dev.so MVWARP_TBL_VARIANT=$v $B 2>/dev/null | python -c "$P"; MVWARP_LIB=mvdet_b200/libmvwarp_dev.so MVWARP_TBL_VARIANT=$v $B --batch 8 2>/dev/null | python -c "$P"; done
compute-sanitizer --tool memcheck python scripts/memcheck_case.py 2>&1 | tail -5
