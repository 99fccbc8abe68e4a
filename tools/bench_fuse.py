#!/usr/bin/env python
"""Quick A/B of the fusion kernels: tiles/s per (config, impl).  usage: bench_fuse.py [impl,impl,...] [N]"""
import json, os, sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, ROOT)
from pistoseg_b200 import ops, synthetic
from pistoseg_b200._lib import DECIDE_SOFTMAX, MASK_FILL
dev = torch.device("cuda:0")
impls = [int(x) for x in (sys.argv[1] if len(sys.argv) > 1 else "2,3,4").split(",")]
N = int(sys.argv[2]) if len(sys.argv) > 2 else 8192


def rep(t, n):
    r = (n + t.shape[0] - 1) // t.shape[0]
    return t.to(dev).repeat((r,) + (1,) * (t.dim() - 1))[:n].contiguous()


def timeit(fn, steps=30, warmup=3):
    for _ in range(warmup): fn()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize(); e0.record()
    for _ in range(steps): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / steps


cases = [
    ("cfg2 mix", synthetic.cfg2(N=1024), dict(mask_mode=MASK_FILL, decide=DECIDE_SOFTMAX, bg_match=1, bg_label=3, lowres=(32, 32))),
    ("cfg2 multi", synthetic.cfg2(N=1024, single_frac=0.0), dict(mask_mode=MASK_FILL, decide=DECIDE_SOFTMAX, bg_match=1, bg_label=3, lowres=(32, 32))),
    ("cfg2 multi nolow", synthetic.cfg2(N=1024, single_frac=0.0), dict(mask_mode=MASK_FILL, decide=DECIDE_SOFTMAX, bg_match=1, bg_label=3)),
    ("cfg2 single", synthetic.cfg2(N=1024, single_frac=1.0), dict(mask_mode=MASK_FILL, decide=DECIDE_SOFTMAX, bg_match=1, bg_label=3, lowres=(32, 32))),
    ("cfg1", synthetic.cfg1(N=1024), dict(decide=DECIDE_SOFTMAX, bg_match=1, bg_label=3)),
    ("cfg3", synthetic.cfg3(N=1000), dict(decide=DECIDE_SOFTMAX)),
]
for name, cfg, kw in cases:
    views = [rep(v, N) for v in cfg["views"]]
    args = {k: (rep(cfg[k], N) if cfg.get(k) is not None else None) for k in ("present", "bg", "gt")}
    conf = ops.new_confusion(cfg["C"], dev) if args["gt"] is not None else None
    row = {}
    for impl in impls:
        fn = lambda: ops.fuse_argmax_confusion(views, cfg["codes"], (224, 224), conf=conf, impl=impl, **args, **kw)
        try:
            ms = timeit(fn)
        except Exception as e:  # no instantiation of this kernel for the shape
            row[impl] = None
            continue
        row[impl] = round(N / ms * 1e3 / 1e6, 3)
    print(name, "Mtiles/s by impl:", json.dumps(row), flush=True)
