"""N > 1 host logic on CPU (gloo, world_size 2): sharding covers every unit exactly once and the merged confusion
matrix equals the single-rank one (integer sums are order-independent)."""
import os
import socket

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from oracle import confusion as oconf
from pistoseg_b200 import dist as pdist


def _free_port():
    s = socket.socket(); s.bind(("127.0.0.1", 0)); p = s.getsockname()[1]; s.close(); return p


def _worker(rank, world, port, n, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"; os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    g = torch.Generator().manual_seed(0)
    pred = torch.randint(0, 4, (n, 32, 32), generator=g, dtype=torch.uint8)
    gt = torch.randint(0, 5, (n, 32, 32), generator=g, dtype=torch.uint8)
    lo, hi = pdist.shard_range(n, rank, world)
    conf = torch.from_numpy(oconf.generate_matrix(pred[lo:hi].numpy(), gt[lo:hi].numpy(), 4).astype(np.int64))
    pdist.all_reduce_confusion(conf)
    if rank == 0:
        q.put((conf.numpy(), oconf.generate_matrix(pred.numpy(), gt.numpy(), 4)))
    dist.destroy_process_group()


def test_sharded_confusion_equals_single_rank():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, 37, q)) for r in range(2)]
    for p in procs: p.start()
    merged, single = q.get(timeout=120)
    for p in procs: p.join(timeout=60)
    assert np.array_equal(merged, single)


def test_shard_range_partitions():
    for n in (0, 1, 7, 10000, 16384):
        for world in (1, 2, 3, 8):
            seen = []
            for r in range(world):
                lo, hi = pdist.shard_range(n, r, world)
                seen += list(range(lo, hi))
            assert seen == list(range(n))


def test_shard_by_key_keeps_images_together():
    keys = [f"{i % 5:02d}" for i in range(40)]
    parts = [pdist.shard_by_key(keys, r, 3) for r in range(3)]
    assert sorted(sum(parts, [])) == list(range(40))
    for r, part in enumerate(parts):
        for other in parts[:r]:
            assert not ({keys[i] for i in part} & {keys[i] for i in other})
