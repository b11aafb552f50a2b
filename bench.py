#!/usr/bin/env python
"""bench.py -- Wildtrack-shape warped frames/s of the warp_views hot path on B200.

  python bench.py --gpus 1 --steps 20 --warmup 5                      # our arm (CUDA kernels through the C ABI)
  python bench.py --impl reference --steps 3 --warmup 1               # reference arm: the CPU oracle on host cores
  python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port P \
         bench.py --gpus N --steps K --warmup W                        # view-sharded, one rank per GPU

One "step" = one pass of the hot path over one batch: forward gather-warp (+ view concat + coord_map)
and the deterministic CSR backward, on synthetic features of the Wildtrack shape
(7 views x 512 ch x 270x480 -> 3586 x 120x360).  Prints ONE JSON line (rank 0).
"""
import argparse
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

METRIC = "Wildtrack-shape warped frames/s (7 views->120x360), warp_views forward+backward"
UNIT = "frames/s"


def parse_args():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", choices=["ours", "reference"], default="ours")
    ap.add_argument("--rig", default="wildtrack", choices=["wildtrack", "multiviewx"])
    ap.add_argument("--batch", type=int, default=1, help="frames per step per GPU")
    ap.add_argument("--channels", type=int, default=512)
    ap.add_argument("--dtype", default="fp32", choices=["fp32", "bf16"])
    ap.add_argument("--layout", default="cl", choices=["cl", "nchw"],
                    help="physical layout of features/output: cl = channel-contiguous (fast path), nchw = reference layout")
    ap.add_argument("--mode", default="fwdbwd", choices=["fwdbwd", "fwd"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--e2e-dense", action="store_true",
                    help="end-to-end leg: copy the WHOLE feature / feature-gradient tensors across PCIe (round-1 behaviour) instead of "
                         "only the texels the warp can touch (mvwarp_copy_touched_texels)")
    ap.add_argument("--fuse-upsample", action="store_true",
                    help="extra configuration (SURVEY.md 8f rank 1): feed the trunk's 90x160 feature map and fuse "
                         "F.interpolate(-> 270x480) into the warp; NOT the headline workload")
    ap.add_argument("--out-pitch", type=int, default=0,
                    help="experiment: channel pitch of the channels-last output / grad_out (0 = dense V*C+2)")
    ap.add_argument("--distinct-cameras", action="store_true",
                    help="N > 1: give rank r cameras [7r, 7r+7) of one 7N-camera ring (every rank then has a different "
                         "homography set, hence different U / nnz and a different amount of work) instead of N copies of the "
                         "N=1 ring (equal work per GPU: the weak-scaling default)")
    ap.add_argument("--ref-channels", type=int, default=512,
                    help="reference arm: channels per step (512 = whole frames, no extrapolation; smaller values are a "
                         "bounded sample scaled linearly and labelled as such)")
    ap.add_argument("--no-conv1", action="store_true", help="skip the fused warp->conv1 (tensor-core) leg")
    ap.add_argument("--no-sharded-conv1", action="store_true",
                    help="skip the view-sharded conv1 leg (BASELINE.json configs[3]: 32-camera rig, partial sums reduced over NVLink)")
    return ap.parse_args()


# --------------------------------------------------------------------------------------------------
# clocks: sampled DURING the timed region through NVML (same counters as the nvidia-smi query line)
# --------------------------------------------------------------------------------------------------
_SAMPLER_CODE = r"""
import sys, time
import pynvml as nv
nv.nvmlInit()
h = nv.nvmlDeviceGetHandleByIndex(int(sys.argv[1]))
reasons = getattr(nv, "nvmlDeviceGetCurrentClocksEventReasons", None) or nv.nvmlDeviceGetCurrentClocksThrottleReasons
print("max", nv.nvmlDeviceGetMaxClockInfo(h, nv.NVML_CLOCK_SM), flush=True)
while True:
    print(time.time(), nv.nvmlDeviceGetClockInfo(h, nv.NVML_CLOCK_SM), reasons(h), flush=True)
    time.sleep(0.002)
"""


class ClockSampler:
    """SM clock + clock-event reasons sampled through NVML by a child process (no GIL contention with the
    launching thread) while the GPU works; `summary(t0, t1)` keeps the samples taken inside the timed region."""
    REASONS = {0x2: "applications_clocks_setting", 0x4: "sw_power_cap", 0x8: "hw_slowdown",
               0x10: "sync_boost", 0x20: "sw_thermal_slowdown", 0x40: "hw_thermal_slowdown",
               0x80: "hw_power_brake_slowdown", 0x100: "display_clock_setting"}

    def __init__(self, index: int):
        import subprocess
        try:
            self.proc = subprocess.Popen([sys.executable, "-c", _SAMPLER_CODE, str(index)], stdout=subprocess.PIPE,
                                         stderr=subprocess.DEVNULL, text=True)
            first = self.proc.stdout.readline().split()
            self.max_mhz = int(first[1]) if first and first[0] == "max" else None
            self.ok = self.max_mhz is not None
        except Exception:  # pragma: no cover
            self.proc, self.max_mhz, self.ok = None, None, False

    def stop(self):
        if self.proc is None:
            return []
        self.proc.terminate()
        out, _ = self.proc.communicate()
        rows = []
        for line in out.splitlines():
            parts = line.split()
            if len(parts) == 3:
                rows.append((float(parts[0]), int(parts[1]), int(parts[2])))
        self.proc = None
        self.rows = rows
        return rows

    def summary(self, t0: float, t1: float):
        rows = getattr(self, "rows", None)
        if rows is None:
            rows = self.stop()
        if not self.ok or not rows:
            return {"sm_mhz": None, "sm_max_mhz": self.max_mhz, "reasons": ["nvml_unavailable"]}
        inside = [r for r in rows if t0 <= r[0] <= t1] or [min(rows, key=lambda r: abs(r[0] - 0.5 * (t0 + t1)))]
        mhz = sorted(r[1] for r in inside)
        mask = 0
        for r in inside:
            mask |= r[2]
        busy = sorted(r[1] for r in rows if r[0] >= t0)   # timed region + the end-to-end leg after it
        return {"sm_mhz": mhz[len(mhz) // 2], "sm_max_mhz": self.max_mhz,
                "reasons": sorted(n for b, n in self.REASONS.items() if mask & b), "samples": len(inside),
                "sm_mhz_busy_median": busy[len(busy) // 2] if busy else None, "samples_busy": len(busy)}


def host_cores() -> int:
    try:
        return len(os.sched_getaffinity(0))
    except AttributeError:  # pragma: no cover
        return os.cpu_count() or 1


def measured_peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        with open(path) as f:
            return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    return 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"


# --------------------------------------------------------------------------------------------------
# reference arm / cpu baseline: the CPU oracle (restated kornia path) on the host cores
# --------------------------------------------------------------------------------------------------
def cpu_reference_run(spec, channels, steps, warmup, mode, threads=None):
    """Times the oracle (V x warp_perspective + cat + coord_map, autograd backward) on `channels` of the
    512 channels of one frame; cost is linear in channels, so frames/s = (channels/512) / seconds."""
    import torch
    from mvdet_b200.geometry import synthetic_rig
    from oracle import kornia_warp as K
    # torchrun exports OMP_NUM_THREADS=1; the CPU arm uses every core this process may run on unless told otherwise
    torch.set_num_threads(threads or host_cores())
    pm = synthetic_rig(spec)
    g = torch.Generator().manual_seed(0)
    x = torch.randn(1, spec.num_cam, channels, *spec.feat_shape, generator=g)
    go = torch.randn(1, spec.num_cam * channels + 2, *spec.grid_shape, generator=torch.Generator().manual_seed(1))
    times = []
    for i in range(warmup + steps):
        t0 = time.perf_counter()
        if mode == "fwd":
            with torch.no_grad():
                K.warp_views_oracle(x, pm, spec.grid_shape)
        else:
            xr = x.requires_grad_(True)
            out = K.warp_views_oracle(xr, pm, spec.grid_shape)
            torch.autograd.grad(out, xr, go)
        dt = time.perf_counter() - t0
        if i >= warmup:
            times.append(dt)
    sec = sum(times) / len(times)
    srt = sorted(times)
    return (channels / 512.0) / sec, sec, torch.get_num_threads(), {"median": srt[len(srt) // 2], "min": srt[0]}


def reference_config(args, spec, channels, threads):
    """The reference arm's OWN configuration: the same frames as the arm under test (same `workload` string), but the truth
    about how the reference runs them -- NCHW-contiguous fp32 tensors on the host, kornia-style per-camera loop, CPU threads."""
    Hs, Ws = spec.feat_shape
    Hg, Wg = spec.grid_shape
    cfg = {
        "workload": f"{spec.name}: B=1 x {spec.num_cam} views x 512ch x {Hs}x{Ws} -> {spec.num_cam * 512 + 2}ch x {Hg}x{Wg}, "
                    f"fp32, {args.mode}",
        "rig": spec.name, "batch_per_gpu": 1, "views_per_gpu": spec.num_cam, "channels": 512,
        "layout": "NCHW contiguous (the layout persp_trans_detector.py:65,69 hands to kornia)", "mode": args.mode,
        "device": "cpu", "threads": threads,
        "parallelism": "single process on the host cores (rank 0 only); the reference has no multi-GPU path for the warp",
        "channels_per_step": channels,
        "sample": "whole frames (512/512 channels per step), no extrapolation" if channels == 512 else
                  f"{channels}/512 channels per step, frames/s scaled linearly",
    }
    return cfg


def run_reference(args):
    import torch
    from mvdet_b200.geometry import RIGS
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return  # the reference is a single-process CPU path: rank 0 alone runs it
    spec = RIGS[args.rig]
    value, sec, threads, stats = cpu_reference_run(spec, args.ref_channels, args.steps, args.warmup, args.mode)
    whole = args.ref_channels == 512
    sample = (f"{'whole' if whole else str(args.ref_channels) + '/512 channels of'} {spec.name} frame{'s' if whole else ''} per step "
              f"({spec.num_cam} views x {args.ref_channels} ch, fp32, NCHW), oracle "
              f"{'fwd+autograd bwd' if args.mode == 'fwdbwd' else 'fwd'}"
              f"{'' if whole else ', scaled linearly to full frames'}")
    line = {
        "impl": "reference", "metric": METRIC if args.mode == "fwdbwd" else METRIC.replace("forward+backward", "forward"),
        "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": sec * 1e3, "ms_per_step_median": stats["median"] * 1e3, "ms_per_step_min": stats["min"] * 1e3,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f32", "data": "synthetic",
        "config": reference_config(args, spec, args.ref_channels, threads),
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": threads, "kind": "port", "sample": sample},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
        "host": {"cpu_count": os.cpu_count(), "torch_threads": threads},
    }
    print(json.dumps(line), flush=True)


def parallelism_note(args, spec, world):
    if world == 1:
        return f"one GPU warps all {spec.num_cam} cameras (views shard across GPUs with no data-path collective: --gpus N)"
    how = (f"cameras [{spec.num_cam}r, {spec.num_cam}r+{spec.num_cam}) of one {spec.num_cam * world}-camera ring (work differs per rank)"
           if getattr(args, "distinct_cameras", False) and world > 1 else
           f"ring r of {world} identical {spec.num_cam}-camera rings (the N=1 work on every GPU)")
    return (f"DATA-PARALLEL REPLICAS of the view-sharded warp: each of {world} ranks warps its own {spec.num_cam} cameras - {how}; the warp "
            f"needs no data-path collective, so `value` is replica throughput (weak scaling, near 1.0 by construction). The path that "
            f"COMMUNICATES -- one rig's cameras partitioned across the ranks, conv1 partial sums reduce-scattered over NVLink -- is the "
            f"`sharded_conv1` key of this line (strong scaling)")


def workload_config(args, spec, world):
    Hs, Ws = spec.feat_shape
    Hg, Wg = spec.grid_shape
    if getattr(args, "fuse_upsample", False):
        return {
            "workload": f"{spec.name} with the up-sample fused: B={args.batch} x {spec.num_cam} views x {args.channels}ch x "
                        f"{Hs // 3}x{Ws // 3} (-> virtual {Hs}x{Ws}) -> {spec.num_cam * args.channels + 2}ch x {Hg}x{Wg}, "
                        f"{args.dtype}, {args.mode}",
            "rig": spec.name, "batch_per_gpu": args.batch, "views_per_gpu": spec.num_cam, "channels": args.channels,
            "layout": "channels_last" if args.layout == "cl" else "NCHW contiguous", "mode": args.mode,
            "parallelism": parallelism_note(args, spec, world),
            "l2": "output (>= 0.3 GB per step) larger than the 126 MB L2; the low-res features (206 MB fp32) are not",
            "note": "extra configuration, not BASELINE.json's metric: frames are the same, the op reads the trunk output directly"}
    return {
        "workload": f"{spec.name}: B={args.batch} x {spec.num_cam} views x {args.channels}ch x {Hs}x{Ws} -> "
                    f"{spec.num_cam * args.channels + 2}ch x {Hg}x{Wg}, {args.dtype}, {args.mode}",
        "rig": spec.name, "batch_per_gpu": args.batch, "views_per_gpu": spec.num_cam, "channels": args.channels,
        "layout": "channels_last (NHWC physical, logical [B,V,C,H,W])" if args.layout == "cl" else "NCHW contiguous",
        "mode": args.mode,
        "parallelism": parallelism_note(args, spec, world),
        "l2": "inputs (>= 1.2 GB per step) larger than the 126 MB L2; no flush needed",
    }


def guarded(fn, *a):
    """An extra leg must never cost the headline line: report its failure instead of raising."""
    try:
        return fn(*a)
    except Exception as e:  # noqa: BLE001
        import traceback
        return {"error": f"{type(e).__name__}: {e}", "where": traceback.format_exc().splitlines()[-3:]}


def tensor_peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        with open(path) as f:
            d = json.load(f)
        return float(d["bf16_tflops_sustained"]), float(d["bf16_tflops"]), "measured (MEASURED_PEAKS.json)"
    return 1500.0, 1500.0, "fallback (B200_PROFILING.md)"


def _event_times(fn, steps, warmup, sync):
    import torch
    for _ in range(warmup):
        fn()
    sync()
    evs = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(steps)]
    for a, b in evs:
        a.record()
        fn()
        b.record()
    sync()
    t = sorted(a.elapsed_time(b) for a, b in evs)
    return {"mean": sum(t) / len(t), "median": t[len(t) // 2], "min": t[0]}


def run_conv1_fused(args, spec, pm, dev, rank, world, dist):
    """north_star (c), BASELINE.json configs[2]: warp_views fused into map_classifier[0] (persp_trans_detector.py:51,81) as one
    tcgen05 implicit GEMM (mvwarp_warp_conv1_fwd_ex), bf16 operands / fp32 accumulate, same rig and batch as the headline
    workload; cuDNN's un-fused equivalent (warp_views bf16 + F.conv2d) is timed beside it.  Every rank runs its own copy."""
    import torch
    import torch.nn.functional as F
    from mvdet_b200 import conv1, ops, plan as P
    V, Cc, B = spec.num_cam, args.channels, args.batch
    Hs, Ws = spec.feat_shape
    Hg, Wg = spec.grid_shape
    gen = torch.Generator(device=dev).manual_seed(100 + rank)
    feats = torch.randn((B, V, Hs, Ws, Cc), device=dev, generator=gen).to(torch.bfloat16).permute(0, 1, 4, 2, 3)
    weight = torch.randn((512, V * Cc + 2, 3, 3), device=dev, generator=gen) * 0.02
    bias = torch.randn((512,), device=dev, generator=gen)
    wpack = conv1.pack_conv1_weight(weight, V, Cc)
    sync = torch.cuda.synchronize
    steps = max(5, min(args.steps, 20))
    t = _event_times(lambda: conv1.warp_conv1(feats, pm, (Hg, Wg), wpack, bias, relu=True), steps, 3, sync)
    flops = 2.0 * B * Hg * Wg * 512 * (V * Cc + 2) * 9
    sustained, burst, src = tensor_peaks()
    # the un-fused path on library kernels: our warp (bf16, channels-last) + cuDNN's convolution + ReLU
    plan = P.get_plan(pm, (Hs, Ws), (Hg, Wg), dev)
    w_cl = weight.to(torch.bfloat16).contiguous(memory_format=torch.channels_last)
    b16 = bias.to(torch.bfloat16)
    with torch.no_grad():
        tc = _event_times(lambda: F.relu_(F.conv2d(ops.warp_views_with_plan(feats, plan), w_cl, b16, padding=1)), steps, 3, sync)
    res = {"kernel": "warp_conv1_direct_kernel (tcgen05 cta_group::2, TMA weight boxes, taps read in place) + conv1_tail_reduce_kernel",
           "workload": f"{spec.name} B={B}: relu(conv3x3(warp_views(feats bf16)) + bias), M={B * Hg * Wg} N=512 K={(V * Cc + 2) * 9}",
           "ms": t["mean"], "ms_median": t["median"], "ms_min": t["min"], "steps": steps,
           "tflops": flops / (t["median"] * 1e-3) / 1e12, "flops_per_step": flops,
           "tflops_note": "nominal: the DENSE flop count over the measured time; the kernel skips the K blocks of views that reach neither "
                          "tile of a CTA pair (12.5 % of them on the Wildtrack rig, 22 % on MultiviewX / the 32-camera rig), so the figure can "
                          "exceed the tensor peak",
           "frac_of_bf16_sustained": flops / (t["median"] * 1e-3) / 1e12 / sustained,
           "frac_of_bf16_burst": flops / (t["median"] * 1e-3) / 1e12 / burst,
           "peak_bf16_tflops": {"sustained": sustained, "burst": burst, "source": src},
           "vs_cudnn_ms": tc["median"], "speedup_vs_cudnn_unfused": tc["median"] / t["median"],
           "frames_per_s": B / (t["median"] * 1e-3), "dtype": "bf16 operands, fp32 accumulate"}
    # persp_trans_detector.py:65 fused in as well: the kernel reads the trunk's LOW-RES map (1/3 resolution) and composes up-sample
    # and warp per halo cell; beside it the two-step path it replaces (bf16 F.interpolate of all views, then the fused kernel)
    if Hs % 3 == 0 and Ws % 3 == 0:
        del feats
        low = torch.randn((B, V, Hs // 3, Ws // 3, Cc), device=dev, generator=gen).to(torch.bfloat16).permute(0, 1, 4, 2, 3)

        def two_step():
            up = F.interpolate(low.flatten(0, 1), (Hs, Ws), mode="bilinear").unflatten(0, (B, V))
            return conv1.warp_conv1(up, pm, (Hg, Wg), wpack, bias, relu=True)
        with torch.no_grad():
            tu = _event_times(lambda: conv1.warp_conv1(low, pm, (Hg, Wg), wpack, bias, relu=True, upsample_shape=(Hs, Ws)), steps, 3, sync)
            t2 = _event_times(two_step, steps, 3, sync)
            a, b2 = conv1.warp_conv1(low, pm, (Hg, Wg), wpack, bias, relu=True, upsample_shape=(Hs, Ws)), two_step()
            diff = float((a - b2).norm() / b2.norm())
        res["upsample_fused"] = {
            "workload": f"relu(conv3x3(warp_views(F.interpolate(low {Hs // 3}x{Ws // 3} -> {Hs}x{Ws}, bilinear))) + bias): low-res map in, "
                        "up-sample + warp composed per halo cell into a 3x3 low-res window (warp_conv1_direct_kernel<UPS>)",
            "ms_median": tu["median"], "ms_min": tu["min"], "tflops": flops / (tu["median"] * 1e-3) / 1e12,
            "two_step_ms_median": t2["median"], "speedup_vs_two_step": t2["median"] / tu["median"],
            "l2_rel_diff_vs_two_step": diff, "frames_per_s": B / (tu["median"] * 1e-3)}
    if world > 1:
        x = torch.tensor([res["ms_median"]], device=dev, dtype=torch.float64)
        dist.all_reduce(x, op=dist.ReduceOp.MAX)
        res["ms_median_max_over_ranks"] = float(x.item())
    return res


def run_conv1_train_step(args, spec, pm, dev):
    """BASELINE.json configs[2]: the Wildtrack TRAINING step at B=8 in bf16 -- relu(map_classifier[0](warp_views(feats))) forward
    + backward with gradients for the features, the weight and the bias.  mvdet_b200.conv1.warp_conv1_trainable: fused tcgen05
    forward, data gradient on the library's TMA-fed tcgen05 convolution, weight gradient on cuDNN, deterministic CSR warp backward;
    beside it the un-fused path (warp_views + F.conv2d under autograd, i.e. cuDNN for all three convolutions)."""
    import torch
    import torch.nn.functional as F
    from mvdet_b200 import conv1, ops
    V, Cc, B = spec.num_cam, 512, 8
    Hs, Ws = spec.feat_shape
    Hg, Wg = spec.grid_shape
    gen = torch.Generator(device=dev).manual_seed(300)
    feats = torch.randn((B, V, Hs, Ws, Cc), device=dev, generator=gen).to(torch.bfloat16).permute(0, 1, 4, 2, 3).requires_grad_(True)
    weight = (torch.randn((512, V * Cc + 2, 3, 3), device=dev, generator=gen) * 0.02).requires_grad_(True)
    bias = torch.randn((512,), device=dev, generator=gen).requires_grad_(True)
    go = torch.randn((B, Hg, Wg, 512), device=dev, generator=gen).permute(0, 3, 1, 2)
    wpack = conv1.pack_conv1_weight(weight, V, Cc)
    w16 = weight.detach().to(torch.bfloat16).contiguous(memory_format=torch.channels_last).requires_grad_(True)
    b16 = bias.detach().to(torch.bfloat16).requires_grad_(True)
    go16 = go.to(torch.bfloat16)

    def fused():
        out = conv1.warp_conv1_trainable(feats, weight, bias, pm, (Hg, Wg), wpack=wpack)
        torch.autograd.grad(out, [feats, weight, bias], go)

    def unfused():
        out = F.relu(F.conv2d(ops.warp_views(feats, pm, (Hg, Wg)), w16, b16, padding=1))
        torch.autograd.grad(out, [feats, w16, b16], go16)

    tf = _event_times(fused, 5, 2, torch.cuda.synchronize)
    tu = _event_times(unfused, 5, 2, torch.cuda.synchronize)
    flops = 3 * 2.0 * B * Hg * Wg * 512 * (V * Cc + 2) * 9          # forward + data gradient + weight gradient
    res = {"workload": f"{spec.name} B={B} bf16 training step: fwd + bwd of relu(conv1(warp_views(feats))), grads for feats / weight / bias",
           "ms": tf["median"], "ms_min": tf["min"], "frames_per_s": B / (tf["median"] * 1e-3),
           "tflops": flops / (tf["median"] * 1e-3) / 1e12, "unfused_cudnn_ms": tu["median"],
           "speedup_vs_unfused": tu["median"] / tf["median"],
           "kernels": "forward: warp_conv1_direct_kernel (tcgen05); ReLU mask + bf16 + bias gradient: relu_bwd_cast_kernel; dgrad: "
                      "conv3x3_tc2_kernel x V (tcgen05 + TMA); wgrad: 3 batched cuBLAS GEMMs on shifted views of zero-bordered "
                      "buffers; warp backward: bwd_cl_kernel (bf16, 16-byte aligned rows); no cuDNN"}
    # the same step from the trunk's LOW-RES output (persp_trans_detector.py:65 inside the op, forward and backward): the gradient
    # comes back at low resolution; beside it torch's F.interpolate + warp_views + cuDNN under autograd
    if Hs % 3 == 0 and Ws % 3 == 0:
        del feats
        torch.cuda.empty_cache()
        low = torch.randn((B, V, Hs // 3, Ws // 3, Cc), device=dev, generator=gen).to(torch.bfloat16).permute(0, 1, 4, 2, 3).requires_grad_(True)

        def fused_up():
            out = conv1.warp_conv1_trainable(low, weight, bias, pm, (Hg, Wg), wpack=wpack, upsample_shape=(Hs, Ws))
            torch.autograd.grad(out, [low, weight, bias], go)

        def unfused_up():   # per camera, as the reference does it (:62-66; one call on all B*V maps exceeds torch's 2^31-element limit at B=8)
            ups = [F.interpolate(low[:, v], (Hs, Ws), mode="bilinear") for v in range(V)]
            out = F.relu(F.conv2d(ops.warp_views_list(ups, pm, (Hg, Wg)), w16, b16, padding=1))
            torch.autograd.grad(out, [low, w16, b16], go16)

        tfu = _event_times(fused_up, 5, 2, torch.cuda.synchronize)
        try:
            tuu = _event_times(unfused_up, 3, 1, torch.cuda.synchronize)
        except Exception as e:  # noqa: BLE001  (the comparison path must not cost the measured one)
            tuu = {"median": float("nan"), "error": f"{type(e).__name__}: {e}"}
        res["from_lowres"] = {"workload": f"the same step from the {Hs // 3}x{Ws // 3} trunk output: F.interpolate (persp_trans_detector.py:65) fused into "
                                          "warp_conv1_trainable(upsample_shape=), low-res gradient out",
                              "ms": tfu["median"], "ms_min": tfu["min"], "frames_per_s": B / (tfu["median"] * 1e-3),
                              "unfused_interpolate_cudnn_ms": tuu["median"], "speedup_vs_unfused": tuu["median"] / tfu["median"],
                              **({"unfused_error": tuu["error"]} if "error" in tuu else {})}
    return res


def run_map_head(args, spec, pm, dev):
    """SURVEY.md 8(f) rank 2 (+ rank 1): the whole ground-plane head in inference -- map_classifier(warp_views(F.interpolate(feats)))
    (persp_trans_detector.py:65, 68-81) -- as MapHead's kernels on the trunk's LOW-RES output (fused up-sample conv1 on tcgen05, TMA-fed
    tcgen05 dilated conv, two-pass 512 -> 1 head), against the same modules on library kernels (torch's F.interpolate, our warp_views,
    three cuDNN convolutions; bf16 channels-last)."""
    import torch
    import torch.nn as nn
    import torch.nn.functional as F
    from mvdet_b200 import map_head as MH, ops
    V, Cc, B = spec.num_cam, args.channels, args.batch
    Hs, Ws = spec.feat_shape
    Hg, Wg = spec.grid_shape
    if Hs % 3 or Ws % 3:
        return {"skipped": "feature map is not a 3x up-sample"}
    torch.manual_seed(400)
    gen = torch.Generator(device=dev).manual_seed(400)
    low = torch.randn((B, V, Hs // 3, Ws // 3, Cc), device=dev, generator=gen).to(torch.bfloat16).permute(0, 1, 4, 2, 3)
    ref = nn.Sequential(nn.Conv2d(Cc * V + 2, 512, 3, padding=1), nn.ReLU(), nn.Conv2d(512, 512, 3, padding=2, dilation=2), nn.ReLU(),
                        nn.Conv2d(512, 1, 3, padding=4, dilation=4, bias=False)).to(dev)
    head = MH.MapHead({"map_classifier." + k: v for k, v in ref.state_dict().items()}, V, Cc, device=dev)
    cudnn = ref.to(torch.bfloat16).to(memory_format=torch.channels_last)
    sync = torch.cuda.synchronize
    steps = max(5, min(args.steps, 20))

    def library():
        ups = [F.interpolate(low[:, v], (Hs, Ws), mode="bilinear") for v in range(V)]       # :65, per camera like the reference
        return cudnn(ops.warp_views_list(ups, pm, (Hg, Wg)))

    with torch.no_grad():
        a = head(low, pm, (Hg, Wg), upsample_shape=(Hs, Ws))
        b = library().float()
        t = _event_times(lambda: head(low, pm, (Hg, Wg), upsample_shape=(Hs, Ws)), steps, 3, sync)
        tl = _event_times(library, max(3, steps // 2), 2, sync)
        up = torch.stack([F.interpolate(low[:, v], (Hs, Ws), mode="bilinear") for v in range(V)], dim=1)
        tf = _event_times(lambda: head(up, pm, (Hg, Wg)), steps, 3, sync)
    return {"workload": f"{spec.name} B={B} inference: map_classifier(warp_views(F.interpolate(trunk output {Hs // 3}x{Ws // 3}))) -> [B,1,{Hg},{Wg}]",
            "kernels": "warp_conv1_direct_kernel<UPS> (+ conv1_tail_reduce_kernel), conv3x3_tc2_kernel, head_partial_kernel, head_stencil_kernel",
            "ms_median": t["median"], "ms_min": t["min"], "frames_per_s": B / (t["median"] * 1e-3),
            "from_upsampled_features_ms": tf["median"], "library_interpolate_warp_cudnn_ms": tl["median"],
            "speedup_vs_library": tl["median"] / t["median"], "max_rel_diff_vs_library_bf16": float((a - b).abs().max() / b.abs().max()),
            "dtype": "bf16 activations and weights, fp32 accumulate"}


def run_sharded_conv1(args, dev, rank, world, dist):
    """BASELINE.json configs[3] / north_star's multi-GPU design, STRONG scaling: ONE 32-camera rig (512 ch x 270x480 -> the
    full-resolution 480x1440 grid), its cameras partitioned across the N ranks.  Each rank runs the fused warp -> conv1 kernel on
    its cameras and its slice of map_classifier[0]'s weight; the fp32 partial sums [1,512,480,1440] are reduce-scattered over
    NVLink (copy engines pushing finished row bands into symmetric-memory inboxes, mvdet_b200.sharding.conv1_fused_reduce_scatter_ce),
    bias + ReLU after the sum.  NCCL's all-reduce of the same partials is timed beside it.  N=1 runs the whole rig on one GPU."""
    import torch
    from mvdet_b200 import conv1, sharding as S
    from mvdet_b200.geometry import RIG32, synthetic_rig
    spec = RIG32
    V, Cc, B = spec.num_cam, 512, 1
    if V % world:
        return {"skipped": f"{V} cameras do not divide over {world} ranks"}
    Hs, Ws = spec.feat_shape
    Hg, Wg = spec.grid_shape
    pm = synthetic_rig(spec)
    shard = S.ViewShard.for_rank(V, rank, world)
    gen = torch.Generator(device=dev).manual_seed(200 + rank)
    feats = torch.randn((B, shard.count, Hs, Ws, Cc), device=dev, generator=gen).to(torch.bfloat16).permute(0, 1, 4, 2, 3)
    # one model: the same seed on every rank gives the same bias; each rank keeps only its input-channel slice of the weight
    wgen = torch.Generator(device=dev).manual_seed(7)
    bias = torch.randn((512,), device=dev, generator=wgen)
    wslice = torch.randn((512, shard.count * Cc + (2 if shard.owns_coord_map else 0), 3, 3), device=dev, generator=gen) * 0.01
    wpack = conv1.pack_conv1_weight(wslice, shard.count, Cc, with_coord=shard.owns_coord_map)
    del wslice

    def sync():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def reduce_max(vals):
        if world == 1:
            return vals
        x = torch.tensor(vals, device=dev, dtype=torch.float64)
        dist.all_reduce(x, op=dist.ReduceOp.MAX)
        return x.tolist()

    steps = 5
    flops_rank = 2.0 * B * Hg * Wg * 512 * (shard.count * Cc + (2 if shard.owns_coord_map else 0)) * 9
    gemm = _event_times(lambda: S.conv1_partial_fused(feats, pm, (Hg, Wg), wpack, bias, shard), steps, 2, sync)
    res = {"rig": f"{V} cameras x {Cc}ch x {Hs}x{Ws} -> 512ch x {Hg}x{Wg}, bf16 operands / fp32 partial sums, B={B}",
           "scaling": "strong (one rig, cameras partitioned)", "views_per_rank": shard.count, "n_gpus": world, "steps": steps}
    if world == 1:
        total = _event_times(lambda: conv1.warp_conv1(feats, pm, (Hg, Wg), wpack, bias, relu=True), steps, 2, sync)
        gm, tm = gemm["median"], total["median"]
        res.update({"path": "single GPU: fused warp -> conv1 over all cameras, no exchange", "gemm_ms": gm, "exchange_ms": 0.0,
                    "total_ms": tm, "bytes_over_nvlink": 0, "tflops_per_rank": flops_rank / (gm * 1e-3) / 1e12,
                    "frames_per_s": B / (tm * 1e-3)})
        return res
    inbox = S.Conv1Inbox((Hg, Wg), B, dev)
    work = torch.empty((B, Hg, Wg, 512), dtype=torch.float32, device=dev)
    ce = _event_times(lambda: S.conv1_fused_reduce_scatter_ce(feats, pm, (Hg, Wg), wpack, bias, shard, inbox, workspace=work),
                      steps, 2, sync)
    # parity of the exchange: my band against the NCCL all-reduce of the same partial sums
    band = S.conv1_fused_reduce_scatter_ce(feats, pm, (Hg, Wg), wpack, bias, shard, inbox, workspace=work)
    del work
    full = S.reduce_partials(S.conv1_partial_fused(feats, pm, (Hg, Wg), wpack, bias, shard), relu=True)
    y0, y1 = inbox.my_rows
    diff = float((band.float() - full[:, :, y0:y1]).abs().max() / full.abs().max())
    del band, full
    nccl = _event_times(lambda: S.conv1_fused_allreduce(feats, pm, (Hg, Wg), wpack, bias, shard, bands=6), steps, 2, sync)
    gm, tm, nm, diff = reduce_max([gemm["median"], ce["median"], nccl["median"], diff])
    partial_bytes = B * 512 * Hg * Wg * 4
    res.update({"path": "reduce_scatter_ce: band-wise fused GEMM in rotated owner order, finished bands pushed to their owner's "
                        "symmetric-memory inbox by peer-to-peer cudaMemcpyAsync (copy engines over NVLink), one barrier, "
                        "mvwarp_inbox_reduce (slot sum + bias + ReLU)",
                "gemm_ms": gm, "exchange_ms": tm - gm, "total_ms": tm,
                "bytes_over_nvlink": int(partial_bytes * (world - 1) / world),
                "tflops_per_rank": flops_rank / (gm * 1e-3) / 1e12, "frames_per_s": B / (tm * 1e-3),
                "nccl_allreduce_overlapped_total_ms": nm,
                "exchange_vs_nccl_allreduce_max_rel_diff": diff,
                "limiter": "the GEMM holds every SM; what is exposed is the last band's copy, one barrier and the inbox reduce",
                "result": "each rank ends with its row band of relu(conv1 + bias)"})
    return res


# --------------------------------------------------------------------------------------------------
# our arm
# --------------------------------------------------------------------------------------------------
def run_ours(args):
    import torch
    import torch.distributed as dist
    from mvdet_b200 import ops, plan as P
    from mvdet_b200.geometry import RIGS, RigSpec, synthetic_rig

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py (impl=ours) needs a CUDA device; there is no CPU fallback")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)

    base = RIGS[args.rig]
    V, Cc, B = base.num_cam, args.channels, args.batch
    if args.distinct_cameras and world > 1:
        # rank r owns cameras [r*V, (r+1)*V) of ONE (world*V)-camera ring generated by the same recipe: the ranks' homographies
        # (and so their unique-texel counts, CSR sizes and kernel times) all differ from the N=1 rig's and from each other
        spec = RigSpec(base.name, base.num_cam * world, base.img_shape, base.worldgrid_shape, base.img_reduce,
                       base.grid_reduce, base.channels, base.seed)
        pm = synthetic_rig(spec)[rank * V:(rank + 1) * V]
    else:
        # weak scaling = the N=1 work on every GPU: a (world*V)-camera installation made of `world` copies of the N=1 ring,
        # rank r warping ring r (its own features, its own channel slice; no data-path collective)
        spec = base
        pm = synthetic_rig(base)
    Hs, Ws = base.feat_shape
    Hg, Wg = base.grid_shape
    if args.fuse_upsample:
        up_hw = (Hs, Ws)
        Hs, Ws = Hs // 3, Ws // 3               # the trunk's stride-8 output (90x160), persp_trans_detector.py:41,65
        plan = P.get_plan_upsampled(pm, (Hs, Ws), up_hw, (Hg, Wg), dev)
    else:
        plan = P.get_plan(pm, (Hs, Ws), (Hg, Wg), dev)
    dtype = torch.float32 if args.dtype == "fp32" else torch.bfloat16
    elem = 4 if args.dtype == "fp32" else 2
    ctot = V * Cc + 2

    gen = torch.Generator(device=dev).manual_seed(rank)
    do_bwd = args.mode == "fwdbwd"
    if args.layout == "cl":
        feats = torch.randn((B, V, Hs, Ws, Cc), device=dev, generator=gen).to(dtype).permute(0, 1, 4, 2, 3)
        pitch = max(args.out_pitch, ctot)
        out = torch.empty((B, Hg, Wg, pitch), device=dev, dtype=dtype)[..., :ctot].permute(0, 3, 1, 2)
        grad_out = grad_feats = None
        if do_bwd or not args.no_e2e:
            grad_out = torch.randn((B, Hg, Wg, pitch), device=dev, generator=gen).to(dtype)[..., :ctot].permute(0, 3, 1, 2)
        if do_bwd:
            grad_feats = torch.empty((B, V, Hs, Ws, Cc), device=dev, dtype=dtype).permute(0, 1, 4, 2, 3)
    else:
        feats = torch.randn((B, V, Cc, Hs, Ws), device=dev, generator=gen).to(dtype)
        # NCHW trunk + fused up-sample: the op's result (and so the gradient cuDNN hands back) is channels-last memory format
        mf = torch.channels_last if args.fuse_upsample else torch.contiguous_format
        out = torch.empty((B, ctot, Hg, Wg), device=dev, dtype=dtype).contiguous(memory_format=mf)
        grad_out = grad_feats = None
        if do_bwd or not args.no_e2e:
            grad_out = torch.randn((B, ctot, Hg, Wg), device=dev, generator=gen).to(dtype).contiguous(memory_format=mf)
        if do_bwd:
            grad_feats = torch.empty((B, V, Cc, Hs, Ws), device=dev, dtype=dtype)

    # NCHW trunk + fused up-sample plan: the op transposes the small low-res map once (mvwarp_layout_transpose) and runs the
    # channel-contiguous kernels; time that whole route (transposes included), not the raw planar kernels
    routed = args.fuse_upsample and args.layout == "nchw"
    cl_out = out.stride(1) == 1

    def fwd():
        if routed:
            return ops._warp_views_plan(feats, ops._register(plan), True, cl_out)
        ops.warp_views_fwd_into(plan, feats, out, 0, 0, V * Cc)

    def bwd():
        if routed:
            return ops._warp_views_plan_bwd(grad_out, ops._register(plan), Cc, False)
        ops.warp_views_bwd_into(plan, grad_out, grad_feats, 0, 0)

    def step():
        fwd()
        if do_bwd:
            bwd()

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    for _ in range(max(args.warmup, 3)):
        step()
    barrier()

    # ---- timed region: K steps, CUDA events on the launching (current) stream ------------------------
    ev = [[torch.cuda.Event(enable_timing=True) for _ in range(3)] for _ in range(args.steps)]
    t_begin, t_end = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    clocks = ClockSampler(local)
    barrier()
    wall0 = time.time()
    t_begin.record()
    for k in range(args.steps):
        ev[k][0].record()
        fwd()
        ev[k][1].record()
        if do_bwd:
            bwd()
        ev[k][2].record()
    t_end.record()
    barrier()
    wall1 = time.time()
    total_ms = t_begin.elapsed_time(t_end)
    fwd_ms = sum(e[0].elapsed_time(e[1]) for e in ev) / args.steps
    bwd_ms = sum(e[1].elapsed_time(e[2]) for e in ev) / args.steps if do_bwd else 0.0
    per_step = sorted(e[0].elapsed_time(e[2]) for e in ev)          # median + min next to the mean
    per_fwd = sorted(e[0].elapsed_time(e[1]) for e in ev)
    per_bwd = sorted(e[1].elapsed_time(e[2]) for e in ev)
    med = lambda v: v[len(v) // 2]
    dist_stats = [med(per_step), per_step[0], med(per_fwd), per_fwd[0], med(per_bwd), per_bwd[0]]
    if world > 1:
        t = torch.tensor([total_ms, fwd_ms, bwd_ms] + dist_stats, device=dev, dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        total_ms, fwd_ms, bwd_ms, *dist_stats = t.tolist()
    ms_per_step = total_ms / args.steps
    frames_per_step = B * world
    value = frames_per_step / (ms_per_step * 1e-3)

    # ---- roofline of the dominant kernel -----------------------------------------------------------
    peak, peak_src = measured_peaks()
    fwd_bytes = B * plan.fwd_bytes(Cc, elem)
    bwd_bytes = B * plan.bwd_bytes(Cc, elem)
    kfwd = ("fwd_table_tile_kernel" if args.layout == "cl" else "transpose_tiles_kernel + fwd_table_tile_kernel") if args.fuse_upsample else \
           ("fwd_cl_kernel" if args.layout == "cl" else "fwd_nchw_kernel")
    roof_fwd = {"kernel": kfwd, "bound": "hbm",
                "achieved": fwd_bytes / (fwd_ms * 1e-3) / 1e9, "peak": peak, "unit": "GB/s",
                "algorithmic_bytes": fwd_bytes, "ms": fwd_ms, "traffic": None}
    roof_fwd["frac"] = roof_fwd["achieved"] / peak
    roofline = roof_fwd
    extra = {}
    if do_bwd:
        roof_bwd = {"kernel": "bwd_cl_kernel" if args.layout == "cl" else ("bwd_cl_kernel + transpose_tiles_kernel" if routed else "bwd_generic_kernel"),
                    "bound": "hbm",
                    "achieved": bwd_bytes / (bwd_ms * 1e-3) / 1e9, "peak": peak, "unit": "GB/s",
                    "algorithmic_bytes": bwd_bytes, "ms": bwd_ms, "traffic": None}
        roof_bwd["frac"] = roof_bwd["achieved"] / peak
        roofline = roof_bwd if bwd_ms >= fwd_ms else roof_fwd
        extra["roofline_fwd"], extra["roofline_bwd"] = roof_fwd, roof_bwd
        extra["roofline_step"] = {"achieved": (fwd_bytes + bwd_bytes) / (ms_per_step * 1e-3) / 1e9, "peak": peak,
                                  "frac": (fwd_bytes + bwd_bytes) / (ms_per_step * 1e-3) / 1e9 / peak, "unit": "GB/s"}
    roofline["peak_source"] = peak_src
    _load_traffic(roofline, extra)

    # ---- e2e: host buffers through the public op, copies inside the timed region ---------------------
    e2e = None
    if not args.no_e2e:
        e2e = run_e2e(args, plan, pm, feats, grad_out, dev, world, dist if world > 1 else None)
    clocks.stop()

    # ---- extra legs, measured in-run so that BENCH / SCALE carry them: the tensor-core path and the communicating path ----
    del feats, out, grad_out, grad_feats
    torch.cuda.empty_cache()
    legs = {}
    if not args.no_conv1 and not args.fuse_upsample and args.layout == "cl" and Cc % 64 == 0:
        legs["conv1_fused"] = guarded(run_conv1_fused, args, base, pm, dev, rank, world, dist if world > 1 else None)
        torch.cuda.empty_cache()
        if world == 1 and Cc == 512:
            legs["conv1_train_step"] = guarded(run_conv1_train_step, args, base, pm, dev)
            torch.cuda.empty_cache()
        if world == 1 and Cc % 64 == 0:
            legs["map_head"] = guarded(run_map_head, args, base, pm, dev)
            torch.cuda.empty_cache()
    if not args.no_sharded_conv1 and not args.fuse_upsample and args.layout == "cl":
        legs["sharded_conv1"] = guarded(run_sharded_conv1, args, dev, rank, world, dist if world > 1 else None)
        torch.cuda.empty_cache()

    # ---- CPU baseline on this box's host cores (rank 0, N=1 only) ------------------------------------
    cpu_baseline = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        # whole 512-channel frames on all host threads (~3 s per step: no extrapolation); the single-thread figure (the
        # reference forces OMP_NUM_THREADS=1, main.py:3) on a 64-channel sample scaled linearly, to keep the run bounded
        v_all, sec_all, thr, _ = cpu_reference_run(base, 512, 2, 1, args.mode)
        v_one, sec_one, _, _ = cpu_reference_run(base, 64, 1, 1, args.mode, threads=1)
        cpu_baseline = {"value": v_all, "unit": UNIT, "cores": thr, "kind": "port",
                        "sample": f"2 whole {base.name} frames (512 channels, fp32 NCHW) after 1 warm-up, oracle (restated kornia "
                                  f"warp x V + cat + coord_map{', autograd bwd' if do_bwd else ''}), all host threads; "
                                  f"single_thread_value from 64/512 channels scaled linearly",
                        "single_thread_value": v_one, "cpu_count": os.cpu_count()}

    if rank == 0:
        line = {
            "metric": METRIC if do_bwd else METRIC.replace("forward+backward", "forward"),
            "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": max(args.warmup, 3),
            "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f32" if args.dtype == "fp32" else "bf16 (fp32 coordinates/weights/accumulate)", "data": "synthetic",
            "config": workload_config(args, base, world),
            "roofline": roofline, **extra,
            "cpu_baseline": cpu_baseline, "e2e": e2e,
            "gpu_launches": args.steps * (2 if do_bwd else 1) * (2 if routed else 1),
            "clocks": clocks.summary(wall0, wall1),
            "plan": {k: plan.stats[k] for k in ("unique_texels", "nnz", "valid_cells", "max_row", "device_bytes")},
            "kernel_ms": {"fwd": fwd_ms, "bwd": bwd_ms},
            "ms_per_step_median": dist_stats[0], "ms_per_step_min": dist_stats[1],
            "kernel_ms_median": {"fwd": dist_stats[2], "bwd": dist_stats[4]},
            "kernel_ms_min": {"fwd": dist_stats[3], "bwd": dist_stats[5]},
            **legs,
        }
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


def kernel_source_hash():
    """sha256 of the warp kernels' sources AS THE PRODUCT LIBRARY COMPILES THEM: `#ifdef MVWARP_DEV ... #endif` blocks (tuning knobs,
    experimental kernels of the development build) are cut out first, so experiments there do not invalidate a capture."""
    import hashlib, re
    h = hashlib.sha256()
    for f in ("mvwarp_kernels.cu", "mvwarp_coords.h", "mvwarp_internal.h"):
        with open(os.path.join(ROOT, "mvdet_b200", "csrc", f)) as fh:
            h.update(re.sub(r"#ifdef MVWARP_DEV.*?#endif[^\n]*\n", "", fh.read(), flags=re.S).encode())
    return h.hexdigest()


def _load_traffic(roofline, extra):
    """DRAM traffic per launch from the committed ncu capture (profiles/traffic.json).  The capture records the hash of the kernel
    sources it was taken from and the workload's algorithmic bytes: if either differs from this run's, the figure would be stale
    and `traffic` stays null (with the reason) instead of silently quoting an old measurement."""
    path = os.path.join(ROOT, "profiles", "traffic.json")
    if not os.path.exists(path):
        return
    try:
        with open(path) as f:
            traffic = json.load(f)
        src = kernel_source_hash()
        for r in [roofline] + [v for k, v in extra.items() if k.startswith("roofline_") and "kernel" in v]:
            t = traffic.get(r["kernel"])
            if not t:
                continue
            if t.get("algorithmic_bytes") != r["algorithmic_bytes"]:
                r["traffic_source"] = "null: the committed capture is of another workload"
            elif t.get("kernel_source_sha256") != src:
                r["traffic_source"] = "null: kernel sources changed since the committed ncu capture (profiles/traffic.json is stale)"
            else:
                r["traffic"] = t["dram_bytes"]
                r["traffic_source"] = ("ncu --set full dram__bytes_read.sum + dram__bytes_write.sum, profiles/" + t["report"] +
                                       ", kernel sources " + src[:12])
    except Exception:
        pass


def run_e2e(args, plan, pm, feats_dev, grad_out_dev, dev, world, dist):
    """Same step through the public op with HOST buffers: every step copies the features (and grad_out) from
    pinned host memory, runs warp_views (+ its autograd backward) and reads the results back to pinned host
    memory.  A stream of frames is double-buffered over three CUDA streams (H2D, compute, D2H) so the two PCIe
    directions and the kernels overlap; every byte of every step still crosses the bus inside the timed region."""
    import torch
    from mvdet_b200 import ops
    do_bwd = args.mode == "fwdbwd"
    B = feats_dev.shape[0]

    def pinned_like(t):  # pinned host mirror with the same physical layout as the device tensor
        if t.dim() == 5 and t.stride(2) == 1:
            return torch.empty(t.permute(0, 1, 3, 4, 2).shape, dtype=t.dtype, pin_memory=True).permute(0, 1, 4, 2, 3)
        if t.dim() == 4 and t.stride(1) == 1:
            return torch.empty(t.permute(0, 2, 3, 1).shape, dtype=t.dtype, pin_memory=True).permute(0, 3, 1, 2)
        return torch.empty(t.shape, dtype=t.dtype, pin_memory=True)

    h_feats, h_go = pinned_like(feats_dev), pinned_like(grad_out_dev)
    h_feats.copy_(feats_dev)
    h_go.copy_(grad_out_dev)
    h_out = pinned_like(grad_out_dev)
    h_gf = pinned_like(feats_dev) if do_bwd else None
    NBUF = 2
    d_feats = [torch.empty_like(feats_dev) for _ in range(NBUF)]
    d_go = [torch.empty_like(grad_out_dev) for _ in range(NBUF)]
    s_in, s_comp, s_out = (torch.cuda.Stream(dev) for _ in range(3))
    s_in_k, s_out_k = torch.cuda.Stream(dev), torch.cuda.Stream(dev)   # the gather / scatter kernels run beside the copy engines
    ev_in_k = [torch.cuda.Event() for _ in range(NBUF)]
    ev_in = [torch.cuda.Event() for _ in range(NBUF)]
    ev_comp = [torch.cuda.Event() for _ in range(NBUF)]     # compute of the step using buffer i is done
    ev_out = [torch.cuda.Event() for _ in range(NBUF)]      # its results are on the host
    steps = max(4, min(args.steps, 32))   # a stream of frames: pipeline fill + drain cost one step's worth of copies in total

    sparse = not args.e2e_dense and feats_dev.stride(2) == 1
    if sparse and do_bwd:
        h_gf.zero_()                                         # zeroed ONCE: rows outside the touched set are zero in every step's gradient

    last = {}

    def one(k):
        i = k % NBUF
        if sparse:
            with torch.cuda.stream(s_in_k):                  # every texel the warp can read (rig-constant set), gathered across PCIe by
                s_in_k.wait_event(ev_comp[i])                # a kernel, concurrently with the copy engine's dense grad_out copy below
                ops.copy_touched_texels(plan, h_feats, d_feats[i])
                ev_in_k[i].record(s_in_k)
        with torch.cuda.stream(s_in):
            s_in.wait_event(ev_comp[i])                      # inputs of step k-2 no longer needed
            if not sparse:
                d_feats[i].copy_(h_feats, non_blocking=True)
            if do_bwd:
                d_go[i].copy_(h_go, non_blocking=True)
            ev_in[i].record(s_in)
        with torch.cuda.stream(s_comp):
            s_comp.wait_event(ev_in[i])
            if sparse:
                s_comp.wait_event(ev_in_k[i])
            x = d_feats[i].requires_grad_(do_bwd)
            # the public op, plan-cache lookup included (the matrices are hashed on every call)
            if plan.upsampled:
                out = ops.warp_views_upsampled(x, plan.src_hw, pm, plan.grid_hw)
            else:
                out = ops.warp_views(x, pm, plan.grid_hw)
            gx = torch.autograd.grad(out, x, d_go[i])[0] if do_bwd else None
            last["gx"] = gx
            d_feats[i].requires_grad_(False)
            ev_comp[i].record(s_comp)
        with torch.cuda.stream(s_out):
            s_out.wait_event(ev_comp[i])
            h_out.copy_(out.detach(), non_blocking=True)
            out.record_stream(s_out)
            if do_bwd and not sparse:
                h_gf.copy_(gx, non_blocking=True)
                gx.record_stream(s_out)
            ev_out[i].record(s_out)
        if do_bwd and sparse:
            with torch.cuda.stream(s_out_k):                 # every non-zero row of the feature gradient, scattered across PCIe
                s_out_k.wait_event(ev_comp[i])
                ops.copy_touched_texels(plan, gx, h_gf)
                gx.record_stream(s_out_k)

    for k in range(NBUF):                                    # warm-up: fills the pipeline, primes the allocator
        one(k)
    torch.cuda.synchronize()
    if dist is not None:
        dist.barrier()
    t0 = time.perf_counter()
    for k in range(steps):
        one(k)
    torch.cuda.synchronize()
    dt = time.perf_counter() - t0
    if dist is not None:
        t = torch.tensor([dt], device=dev, dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        dt = float(t.item())
    nbytes = lambda t: t.numel() * t.element_size()
    feat_bytes = nbytes(h_feats)
    if sparse:                                               # bytes actually moved: U touched texels x C channels per frame
        feat_bytes = B * plan.stats["unique_texels"] * feats_dev.shape[2] * feats_dev.element_size()
    h2d = feat_bytes + (nbytes(h_go) if do_bwd else 0)
    d2h = nbytes(h_out) + (feat_bytes if do_bwd else 0)
    if do_bwd and sparse:                                    # the host gradient must equal the dense one (untouched rows: zero)
        check = bool(torch.equal(h_gf, last["gx"].cpu()))
    else:
        check = None
    return {"value": B * world * steps / dt, "unit": UNIT, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
            "steps": steps, "ms_per_step": dt / steps * 1e3,
            "transfer": ("sparse: only the texels some grid cell samples cross PCIe (U of V*Hs*Ws = "
                         f"{plan.stats['unique_texels']} of {plan.V * plan.feat_hw[0] * plan.feat_hw[1]}, rig-constant) -- gathered from / "
                         "scattered into mapped pinned host memory by mvwarp_copy_touched_texels; out and grad_out are dense copies; the "
                         "host gradient tensor is zeroed once" if sparse else "dense: whole tensors, cudaMemcpyAsync"),
            "dense_bytes_each_way": nbytes(h_feats) + nbytes(h_out),
            "host_gradient_equals_dense": check,
            "pcie_gbs_each_way": [h2d * steps / dt / 1e9, d2h * steps / dt / 1e9],
            "api": "mvdet_b200.ops.copy_touched_texels + mvdet_b200.warp_views (plan-cache lookup included) + torch.autograd.grad on pinned host buffers; H2D / compute / D2H "
                   "double-buffered on three streams, all copies inside the timed region"}


def main():
    args = parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
