#!/usr/bin/env python
"""Probe: torch symmetric memory on this box (peer pointers for P2P stores from our kernels)."""
import os, torch, torch.distributed as dist
import torch.distributed._symmetric_memory as symm
rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
dist.init_process_group("nccl", device_id=dev)
t = symm.empty(1 << 20, dtype=torch.float32, device=dev)
hdl = symm.rendezvous(t, dist.group.WORLD)
print(rank, "buffer_ptrs", [hex(p) for p in hdl.buffer_ptrs], "world", hdl.world_size, "rank", hdl.rank, flush=True)
t.fill_(rank + 1.0)
hdl.barrier()
peer = hdl.get_buffer((rank + 1) % world, (1 << 20,), torch.float32)
print(rank, "peer value", float(peer[0]), "multicast_ptr", hex(getattr(hdl, "multicast_ptr", 0) or 0), flush=True)
hdl.barrier()
dist.destroy_process_group()
