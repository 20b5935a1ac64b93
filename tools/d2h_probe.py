"""Raw device->host copy bandwidth of N ranks at once (judge item: is the 8-GPU end-to-end figure bound by the host?).
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port 29511 tools/d2h_probe.py
Every rank pins 1 GiB of host memory, copies 1 GiB device -> host K times with one cudaMemcpyAsync per copy (CUDA events), first
all ranks together (barrier before), then one rank at a time.  Rank 0 prints one JSON line."""
import json
import os
import time

import torch
import torch.distributed as dist

rank = int(os.environ.get("RANK", "0"))
world = int(os.environ.get("WORLD_SIZE", "1"))
local = int(os.environ.get("LOCAL_RANK", "0"))
torch.cuda.set_device(local)
if world > 1:
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
N = 1 << 30
K = 8
dev = torch.empty(N, dtype=torch.uint8, device="cuda")
dev.random_(0, 255)
host = torch.empty(N, dtype=torch.uint8, pin_memory=True)
host.zero_()  # first touch from this rank's CPUs


def timed(k):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(k):
        host.copy_(dev, non_blocking=True)
    e1.record()
    torch.cuda.synchronize()
    return k * N / (e0.elapsed_time(e1) * 1e-3) / 1e9


def barrier():
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()


timed(2)
barrier()
t0 = time.perf_counter()
together = timed(K)
barrier()
wall = time.perf_counter() - t0
alone = 0.0
for r in range(world):
    barrier()
    if r == rank:
        alone = timed(K)
barrier()
vals = torch.tensor([together, alone], device="cuda", dtype=torch.float64)
if world > 1:
    allv = [torch.zeros_like(vals) for _ in range(world)]
    dist.all_gather(allv, vals)
else:
    allv = [vals]
if rank == 0:
    tog = [round(float(v[0]), 2) for v in allv]
    alo = [round(float(v[1]), 2) for v in allv]
    print(json.dumps({"probe": "pinned D2H, 1 GiB x %d per rank, one cudaMemcpyAsync per copy" % K, "ranks": world,
                      "GBps_per_rank_all_ranks_together": tog, "GBps_aggregate_together": round(sum(tog), 2),
                      "GBps_aggregate_wall_clock": round(world * K * N / wall / 1e9, 2),
                      "GBps_per_rank_alone": alo, "cpus": os.cpu_count()}))
if world > 1:
    dist.destroy_process_group()
