import json, os, sys, time
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, ROOT)
from pistoseg_b200 import mosaic, ops
dev = torch.device("cuda:0")
rng = np.random.default_rng(0)
P = 2048
imgs = [rng.integers(0, 256, (224, 224, 3), dtype=np.uint8) for _ in range(64)]; imgs = [imgs[i % 64] for i in range(P)]
bgs = [(rng.random((224, 224)) < 0.1).astype(np.uint8) * 255 for _ in range(64)]; bgs = [bgs[i % 64] for i in range(P)]
pool = mosaic.TilePool(imgs, rng.integers(0, 3, P).astype(np.uint8), bgs, device=dev)
for pn, ps in ((4, 56),):
    planner = mosaic.MosaicPlanner(pool, pn, ps, reject_bg=False)
    plans, cells = planner.plans(range(256))
    reps = int(sys.argv[1]) if len(sys.argv) > 1 else 32
    plans = np.tile(plans, reps); cells = np.tile(cells, (reps, 1, 1))
    pl = torch.from_numpy(plans.view(np.uint8).reshape(-1)).to(dev); ce = torch.from_numpy(np.ascontiguousarray(cells).view(np.uint8).reshape(-1)).to(dev)
    for _ in range(3): ops.mosaic_gather(pool.dev, pl, ce, pn, ps, 3)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize(); e0.record()
    for _ in range(5): ops.mosaic_gather(pool.dev, pl, ce, pn, ps, 3)
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 5
    print(pn, ps, len(plans), "mosaics", ms, "ms", len(plans) / ms * 1e3, "mosaics/s", len(plans) * 401408 / ms / 1e6, "GB/s")
