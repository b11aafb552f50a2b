#!/usr/bin/env python
"""Context for the roofline fractions: what plain torch kernels reach on this GPU for the three traffic mixes the warp
kernels have -- pure write (memset), copy (1 read : 1 write), pure read (sum) -- CUDA-event timed, buffers >> L2."""
import json, torch
dev = "cuda:0"
n = 1 << 29                      # 2 GiB of fp32
a = torch.empty(n, dtype=torch.float32, device=dev)
b = torch.empty(n, dtype=torch.float32, device=dev)
def timeit(fn, iters=10):
    for _ in range(3): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters
gb = n * 4 / 1e9
out = {"bytes_per_buffer_GB": gb,
       "memset_GBps": gb / timeit(lambda: a.zero_()) * 1e3,
       "copy_GBps_read_plus_write": 2 * gb / timeit(lambda: b.copy_(a)) * 1e3,
       "read_sum_GBps": gb / timeit(lambda: a.sum()) * 1e3}
# 1 read : 3 write mix (the backward's): write three buffers from one
c = torch.empty(n // 4, dtype=torch.float32, device=dev)
def r1w3():
    torch.add(c, 1.0, out=a[: n // 4]); torch.add(c, 2.0, out=a[n // 4: n // 2]); torch.add(c, 3.0, out=a[n // 2: 3 * n // 4])
out["r1_w3_GBps_three_launches"] = (n // 4 * 4 * 6 / 1e9) / timeit(r1w3) * 1e3
print(json.dumps(out))
