#!/usr/bin/env python
"""development: all-reduce bandwidth probe (what does NCCL deliver between the GPUs of this box?)"""
import os, sys, time, torch, torch.distributed as dist
rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
for mb in (64, 708, 1024):
    x = torch.ones(mb * 1000 * 1000 // 4, device="cuda")
    for _ in range(3): dist.all_reduce(x)
    torch.cuda.synchronize(); dist.barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(10): dist.all_reduce(x)
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 10
    if rank == 0:
        print(f"all_reduce {mb} MB x {world} ranks: {ms:.3f} ms, busbw {2*(world-1)/world*mb/1e3/(ms/1e3):.1f} GB/s", flush=True)
if rank == 0:
    print("p2p access 0->1:", torch.cuda.can_device_access_peer(0, 1))
dist.destroy_process_group()
