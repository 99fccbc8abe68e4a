#!/usr/bin/env python
"""Per-tile-class rates of the filtered kernels: all tiles with 2 / 3 classes present (one / two difference fields), with and
without the 32x32 export.  usage: bench_k.py impl,impl,..."""
import json, os, sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, ROOT)
from pistoseg_b200 import ops, synthetic
from pistoseg_b200._lib import DECIDE_SOFTMAX, MASK_FILL
dev = torch.device("cuda:0")
impls = [int(x) for x in (sys.argv[1] if len(sys.argv) > 1 else "4,6").split(",")]
N = 8192

def rep(t, n):
    r = (n + t.shape[0] - 1) // t.shape[0]
    return t.to(dev).repeat((r,) + (1,) * (t.dim() - 1))[:n].contiguous()

def timeit(fn, steps=10, warmup=3):
    for _ in range(warmup): fn()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize(); e0.record()
    for _ in range(steps): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / steps

cfg = synthetic.cfg2(N=1024, single_frac=0.0)
views = [rep(v, N) for v in cfg["views"]]
bg = rep(cfg["bg"], N)
for name, pres in (("P=2", [1, 1, 0]), ("P=3", [1, 1, 1])):
    present = torch.tensor(pres, dtype=torch.uint8).repeat(N, 1).to(dev)
    for low in (None, (32, 32)):
        row = {}
        for impl in impls:
            fn = lambda: ops.fuse_argmax_confusion(views, cfg["codes"], (224, 224), impl=impl, present=present, bg=bg, mask_mode=MASK_FILL,
                                                   decide=DECIDE_SOFTMAX, bg_match=1, bg_label=3, lowres=low)
            try: row[impl] = round(N / timeit(fn) * 1e3 / 1e6, 3)
            except Exception: row[impl] = None
        print(name, "export" if low else "no export", "Mtiles/s by impl:", json.dumps(row), flush=True)
