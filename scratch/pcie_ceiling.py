"""Platform ceiling of the host-buffer (e2e) measurement: every rank copies a 161 MB pinned buffer host->device and another
device->host at the same time (two streams), nothing else -- the aggregate PCIe / host-memory throughput the box can give N
GPUs at once.  torchrun --nproc-per-node N scratch/pcie_ceiling.py"""
import json
import os
import time

import torch
import torch.distributed as dist

rank, world = int(os.environ.get("RANK", "0")), int(os.environ.get("WORLD_SIZE", "1"))
torch.cuda.set_device(int(os.environ.get("LOCAL_RANK", "0")))
if world > 1:
    dist.init_process_group("gloo")
n = 161_000_000 // 8
h_in = torch.empty(n, dtype=torch.float64).pin_memory()
h_out = torch.empty(n, dtype=torch.float64).pin_memory()
d_in = torch.empty(n, dtype=torch.float64, device="cuda")
d_out = torch.zeros(n, dtype=torch.float64, device="cuda")
s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()


def once(reps):
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    t0 = time.time()
    for _ in range(reps):
        with torch.cuda.stream(s1):
            d_in.copy_(h_in, non_blocking=True)
        with torch.cuda.stream(s2):
            h_out.copy_(d_out, non_blocking=True)
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    return time.time() - t0


once(3)
reps = 20
dt = once(reps)
if rank == 0:
    per_gpu = 2 * n * 8 * reps / dt / 1e9
    print(json.dumps({"n_gpus": world, "bytes_each_way_per_copy": n * 8, "GBps_per_gpu_both_directions": per_gpu,
                      "GBps_aggregate": per_gpu * world, "ms_per_round_trip_pair": 1e3 * dt / reps}))
if world > 1:
    dist.barrier()
    dist.destroy_process_group()
