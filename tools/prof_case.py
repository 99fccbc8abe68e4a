#!/usr/bin/env python
"""Run one fusion case a few times (for ncu).  usage: prof_case.py {cfg2|cfg2multi|cfg2single|cfg1|cfg3} impl [N]"""
import os, sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, ROOT)
from pistoseg_b200 import ops, synthetic
from pistoseg_b200._lib import DECIDE_RAW, DECIDE_SOFTMAX, FUSE_PROB_MEAN, MASK_FILL
dev = torch.device("cuda:0")
name, impl = sys.argv[1], int(sys.argv[2])
N = int(sys.argv[3]) if len(sys.argv) > 3 else 4096
kw2 = dict(mask_mode=MASK_FILL, decide=DECIDE_SOFTMAX, bg_match=1, bg_label=3, lowres=(32, 32))
cfg, kw = {"cfg2": (lambda: synthetic.cfg2(N=1024), kw2), "cfg2multi": (lambda: synthetic.cfg2(N=1024, single_frac=0.0), kw2),
           "cfg2multinolow": (lambda: synthetic.cfg2(N=1024, single_frac=0.0), dict(mask_mode=MASK_FILL, decide=DECIDE_SOFTMAX, bg_match=1, bg_label=3)),
           "cfg2prob": (lambda: synthetic.cfg2(N=1024), dict(fuse_mode=FUSE_PROB_MEAN, decide=DECIDE_RAW, bg_match=1, bg_label=3)),
           "cfg2single": (lambda: synthetic.cfg2(N=1024, single_frac=1.0), kw2),
           "cfg1": (lambda: synthetic.cfg1(N=1024), dict(decide=DECIDE_SOFTMAX, bg_match=1, bg_label=3)),
           "cfg3": (lambda: synthetic.cfg3(N=1000), dict(decide=DECIDE_SOFTMAX)),
           "cfg5": (lambda: synthetic.cfg5(N=8, T=512), dict(decide=DECIDE_SOFTMAX))}[name]
cfg = cfg()


def rep(t, n):
    r = (n + t.shape[0] - 1) // t.shape[0]
    return t.to(dev).repeat((r,) + (1,) * (t.dim() - 1))[:n].contiguous()


views = [rep(v, N) for v in cfg["views"]]
args = {k: (rep(cfg[k], N) if cfg.get(k) is not None else None) for k in ("present", "bg", "gt")}
conf = ops.new_confusion(cfg["C"], dev) if args["gt"] is not None else None
for _ in range(4):
    ops.fuse_argmax_confusion(views, cfg["codes"], (cfg["T"], cfg["T"]), conf=conf, impl=impl, **args, **kw)
torch.cuda.synchronize()
print("ok")
