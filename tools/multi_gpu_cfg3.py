#!/usr/bin/env python
"""BASELINE config 3 under torchrun: 10 000 BCSS-shaped tiles (C = 4, 3 scales x flip, gt in {0..4}) in contiguous shards
over the ranks, one fused kernel per rank, ONE all-reduce (NCCL) of the int64 [4,4] confusion matrix; rank 0 checks that the
merged matrix equals the matrix of the unsharded run exactly and prints one JSON line.
  python -m torch.distributed.run --nnodes=1 --nproc-per-node G --master-addr 127.0.0.1 --master-port P tools/multi_gpu_cfg3.py"""
import json, os, sys
import torch
import torch.distributed as dist
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, ROOT)
from pistoseg_b200 import dist as pdist, ops, synthetic
from pistoseg_b200._lib import DECIDE_SOFTMAX

world, rank, local = int(os.environ.get("WORLD_SIZE", "1")), int(os.environ.get("RANK", "0")), int(os.environ.get("LOCAL_RANK", "0"))
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
if world > 1:
    dist.init_process_group("nccl", device_id=dev)
N = 10000
base = synthetic.cfg3(N=1000)                       # seeded 1000-tile block, identical on every rank
rep = lambda t: t.to(dev).repeat((10,) + (1,) * (t.dim() - 1)).contiguous()
views, gt = [rep(v) for v in base["views"]], rep(base["gt"])
lo, hi = pdist.shard_range(N, rank, world)


def run(a, b, conf):
    return ops.fuse_argmax_confusion([v[a:b] for v in views], base["codes"], (224, 224), decide=DECIDE_SOFTMAX, gt=gt[a:b], conf=conf)


conf = ops.new_confusion(4, dev)
for _ in range(3):
    run(lo, hi, conf)
conf.zero_()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
if world > 1:
    dist.barrier()
torch.cuda.synchronize(); e0.record()
steps = 20
for _ in range(steps):
    conf.zero_()
    run(lo, hi, conf)
    pdist.all_reduce_confusion(conf)               # the one collective of the path
e1.record(); torch.cuda.synchronize()
ms = torch.tensor([e0.elapsed_time(e1)], device=dev)
if world > 1:
    dist.all_reduce(ms, op=dist.ReduceOp.MAX)
if rank == 0:
    full = ops.new_confusion(4, dev)
    run(0, N, full)
    torch.cuda.synchronize()
    exact = bool(torch.equal(full, conf))
    print(json.dumps({"config": "cfg3: 10k BCSS-shaped tiles, C=4, V=6, gt, confusion all-reduce", "n_gpus": world, "tiles_per_s": N * steps / (float(ms.item()) * 1e-3),
                      "merged_equals_single_gpu_matrix": exact, "total_counted": int(conf.sum().item()), "collective": "all_reduce(SUM) int64[16] per step"}))
    assert exact
if world > 1:
    dist.destroy_process_group()
