#!/usr/bin/env python
"""profiles/traffic.json from one evidence run: the per-kernel DRAM bytes of the ncu capture (scripts/ncu_summary.py --traffic, which
also records the hash of the kernel sources) joined with the algorithmic bytes of the SAME workload from that run's bench line.
bench.py copies a kernel's figure into roofline.traffic only when both the workload and the source hash match the run at hand.
usage: python scripts/make_traffic.py gpurun_out/r02_traffic_default.json gpurun_out/r02_bench.json profiles/traffic.json"""
import json, sys
traffic, bench, out = json.load(open(sys.argv[1])), json.load(open(sys.argv[2])), sys.argv[3]
algo = {}
for key in ("roofline", "roofline_fwd", "roofline_bwd"):
    r = bench.get(key)
    if r and "kernel" in r:
        algo[r["kernel"]] = r["algorithmic_bytes"]
res = {}
for k, t in traffic.items():
    if k in algo:
        res[k] = dict(t, algorithmic_bytes=algo[k], workload=bench["config"]["workload"])
json.dump(res, open(out, "w"), indent=1)
print(json.dumps(res, indent=1))
