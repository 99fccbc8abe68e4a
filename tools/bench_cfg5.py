#!/usr/bin/env python
"""BASELINE config 5 (large tiles, 5 scales x flip, C = 4, gt + labels + confusion): tiles/s for T = 512 / 1024 / 2048."""
import os, sys, json, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, ROOT)
from pistoseg_b200 import ops, synthetic
from pistoseg_b200._lib import DECIDE_SOFTMAX
dev = torch.device("cuda:0")
for T, N in ((512, 512), (1024, 128), (2048, 32)):
    cfg = synthetic.cfg5(N=8, T=T)
    rep = lambda t: t.to(dev).repeat((N // 8,) + (1,) * (t.dim() - 1)).contiguous()
    views = [rep(v) for v in cfg["views"]]; gt = rep(cfg["gt"]); conf = ops.new_confusion(4, dev)
    fn = lambda: ops.fuse_argmax_confusion(views, cfg["codes"], (T, T), decide=DECIDE_SOFTMAX, gt=gt, conf=conf)
    for _ in range(2): fn()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize(); e0.record()
    for _ in range(5): fn()
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 5
    b = 4 * 4 * sum(v.shape[2] * v.shape[3] for v in views) + 2 * T * T
    print(json.dumps({"T": T, "tiles_per_s": N / ms * 1e3, "GB/s": N * b / ms / 1e6}))
    del views, gt
