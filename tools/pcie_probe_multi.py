#!/usr/bin/env python
"""Host-fabric probe under torchrun: every rank copies pinned host memory to / from ITS GPU at the same time (barrier-aligned), so the
aggregate shows where the host side saturates, independently of any kernel of this library.  One JSON line on rank 0.
usage: python -m torch.distributed.run --nproc-per-node N --master-addr 127.0.0.1 tools/pcie_probe_multi.py"""
import json, os, time
import torch
import torch.distributed as dist
rank, world, local = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
if world > 1:
    dist.init_process_group("nccl", device_id=dev)
ncpu = os.cpu_count() or 1
per = max(ncpu // world, 1)
cores = list(range(local * per, min((local + 1) * per, ncpu)))
try:
    os.sched_setaffinity(0, cores)      # before the pinned allocation: first touch on the rank's own cores
except OSError:
    cores = []
SZ = 1 << 30
h_in = torch.empty(SZ, dtype=torch.uint8).pin_memory(); h_in.fill_(1)
h_out = torch.empty(SZ // 2, dtype=torch.uint8).pin_memory()
d_in = torch.empty(SZ, dtype=torch.uint8, device=dev); d_out = torch.empty(SZ // 2, dtype=torch.uint8, device=dev)
s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()

def timed(fn, n=4):
    fn(); torch.cuda.synchronize()
    if world > 1: dist.barrier()
    torch.cuda.synchronize(); t0 = time.perf_counter()
    for _ in range(n): fn()
    torch.cuda.synchronize()
    dt = torch.tensor([(time.perf_counter() - t0) / n], device=dev)
    if world > 1: dist.all_reduce(dt, op=dist.ReduceOp.MAX)
    return float(dt.item())

def both():
    with torch.cuda.stream(s1): d_in.copy_(h_in, non_blocking=True)
    with torch.cuda.stream(s2): h_out.copy_(d_out, non_blocking=True)

t_h2d = timed(lambda: d_in.copy_(h_in, non_blocking=True))
t_d2h = timed(lambda: h_out.copy_(d_out, non_blocking=True))
t_both = timed(both)
if rank == 0:
    print(json.dumps({"ranks": world, "cores_per_rank": len(cores), "host_cores": ncpu,
                      "h2d_alone_gbs_per_rank": SZ / t_h2d / 1e9, "h2d_alone_gbs_total": world * SZ / t_h2d / 1e9,
                      "d2h_alone_gbs_per_rank": SZ / 2 / t_d2h / 1e9, "d2h_alone_gbs_total": world * SZ / 2 / t_d2h / 1e9,
                      "both_h2d_gbs_total": world * SZ / t_both / 1e9, "both_d2h_gbs_total": world * SZ / 2 / t_both / 1e9,
                      "note": "max over ranks of the wall time of 4 barrier-aligned 1 GiB (H2D) / 0.5 GiB (D2H) pinned copies"}))
if world > 1:
    dist.destroy_process_group()
