#!/usr/bin/env python
"""Summarise an .ncu-rep (read here, no GPU): per-kernel headline metrics + stall breakdown.
usage: python scripts/ncu_summary.py gpurun_out/prof_x.ncu-rep [> profiles/rNN_x_summary.txt]"""
import csv, io, subprocess, sys
rep = sys.argv[1]
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, units = rows[0], rows[1]
idx = {h: i for i, h in enumerate(hdr)}
WANT = ['gpu__time_duration.sum', 'dram__bytes_read.sum', 'dram__bytes_write.sum',
        'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed', 'lts__throughput.avg.pct_of_peak_sustained_elapsed',
        'l1tex__throughput.avg.pct_of_peak_sustained_elapsed', 'sm__throughput.avg.pct_of_peak_sustained_elapsed',
        'lts__t_sector_hit_rate.pct', 'l1tex__t_sector_hit_rate.pct', 'lts__t_sectors_srcunit_tex_op_read.sum',
        'lts__t_sectors_srcunit_tex_op_write.sum', 'l1tex__t_sectors_pipe_lsu_mem_global_op_ld.sum',
        'l1tex__t_requests_pipe_lsu_mem_global_op_ld.sum', 'sm__warps_active.avg.pct_of_peak_sustained_active',
        'launch__registers_per_thread', 'launch__occupancy_limit_registers', 'launch__waves_per_multiprocessor',
        'launch__grid_size', 'launch__block_size', 'launch__shared_mem_per_block_static',
        'smsp__inst_executed.sum', 'smsp__issue_active.avg.pct_of_peak_sustained_active',
        'smsp__warps_eligible.avg.per_cycle_active', 'sm__cycles_active.avg', 'sm__cycles_active.max', 'sm__cycles_active.min',
        'sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active',
        'l1tex__data_pipe_lsu_wavefronts.sum', 'l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed',
        'l1tex__t_sectors_pipe_lsu_mem_global_op_st.sum', 'l1tex__t_requests_pipe_lsu_mem_global_op_st.sum',
        'smsp__inst_executed_pipe_tensor.sum', 'sm__inst_executed_pipe_tma.sum']
for r in rows[2:]:
    print("=" * 100)
    print(r[idx['Kernel Name']], " grid", r[idx.get('Grid Size', 0)], "block", r[idx.get('Block Size', 0)])
    for w in WANT:
        if w in idx:
            print(f"  {w:72s} {r[idx[w]]:>18s} {units[idx[w]]}")
    st = []
    for h, i in idx.items():
        if h.startswith('smsp__average_warps_issue_stalled_') and h.endswith('_per_issue_active.ratio'):
            try:
                st.append((float(r[i].replace(',', '')), h[len('smsp__average_warps_issue_stalled_'):-len('_per_issue_active.ratio')]))
            except ValueError:
                pass
    st.sort(reverse=True)
    print("  warp stall cycles per issued instruction (top):", ", ".join(f"{n}={v:.2f}" for v, n in st[:7]))

# --traffic OUT.json ALGO_JSON: also write per-kernel DRAM traffic (read+write bytes per launch, averaged over the
# captured launches) keyed by short kernel name, for bench.py's roofline.traffic
if "--traffic" in sys.argv:
    import json, re, os
    out = sys.argv[sys.argv.index("--traffic") + 1]
    def to_bytes(v, u):
        mult = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "Tbyte": 1e12}[u]
        return float(v.replace(",", "")) * mult
    acc = {}
    for r in rows[2:]:
        name = re.sub(r"^void\s+", "", r[idx['Kernel Name']])
        short = re.match(r"(?:mvw::)?([A-Za-z0-9_]+)", name).group(1)
        rd = to_bytes(r[idx['dram__bytes_read.sum']], units[idx['dram__bytes_read.sum']])
        wr = to_bytes(r[idx['dram__bytes_write.sum']], units[idx['dram__bytes_write.sum']])
        acc.setdefault(short, []).append(rd + wr)
    prev = json.load(open(out)) if os.path.exists(out) else {}
    import hashlib
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    h = hashlib.sha256()   # same rule as bench.py kernel_source_hash(): development-only blocks do not count
    for f in ("mvwarp_kernels.cu", "mvwarp_coords.h", "mvwarp_internal.h"):
        h.update(re.sub(r"#ifdef MVWARP_DEV.*?#endif[^\n]*\n", "", open(os.path.join(root, "mvdet_b200", "csrc", f)).read(), flags=re.S).encode())
    for k, v in acc.items():
        prev[k] = {"dram_bytes": sum(v) / len(v), "launches": len(v), "report": os.path.basename(rep), "kernel_source_sha256": h.hexdigest()}
    json.dump(prev, open(out, "w"), indent=1)
