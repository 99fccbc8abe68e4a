#!/usr/bin/env python
"""Stress of the shape-specialised kernel's tile hand-over (producer / compute / fixer warps): batch sizes around multiples of the SM
count, every output compared bit for bit with the generic one-thread-per-pixel kernel.  usage: stress_static.py [rounds]"""
import os, sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, ROOT)
from pistoseg_b200 import ops, synthetic
from pistoseg_b200._lib import DECIDE_SOFTMAX, IMPL_GENERIC, MASK_FILL
dev = torch.device("cuda:0")
rounds = int(sys.argv[1]) if len(sys.argv) > 1 else 2
sms = torch.cuda.get_device_properties(0).multi_processor_count
sizes = [1, 2, 3, sms - 1, sms, sms + 1, 2 * sms - 1, 2 * sms, 2 * sms + 1, 3 * sms + 7, 1000]
bad = 0
for r in range(rounds):
    for N in sizes:
        for name, cfg, kw in (("cfg2", synthetic.cfg2(N=N), dict(mask_mode=MASK_FILL, decide=DECIDE_SOFTMAX, bg_match=1, bg_label=3, lowres=(32, 32))),
                              ("cfg1", synthetic.cfg1(N=N), dict(decide=DECIDE_SOFTMAX, bg_match=1, bg_label=3)),
                              ("cfg3", synthetic.cfg3(N=N), dict(decide=DECIDE_SOFTMAX))):
            views = [v.to(dev) for v in cfg["views"]]
            args = {k: (cfg[k].to(dev) if cfg.get(k) is not None else None) for k in ("present", "bg", "gt")}
            outs = []
            for impl in (0, IMPL_GENERIC):
                conf = ops.new_confusion(cfg["C"], dev) if args["gt"] is not None else None
                o = ops.fuse_argmax_confusion(views, cfg["codes"], (224, 224), conf=conf, impl=impl, **args, **kw)
                outs.append((o, conf))
            (a, ca), (b, cb) = outs
            ok = torch.equal(a["labels"], b["labels"]) and (ca is None or torch.equal(ca, cb)) and ("lowres" not in a or torch.equal(a["lowres"], b["lowres"]))
            if not ok:
                bad += 1
                print("MISMATCH", name, N, "round", r, int((a["labels"] != b["labels"]).sum()))
torch.cuda.synchronize()
print("stress done:", rounds, "rounds,", len(sizes) * 3 * rounds, "cases,", bad, "mismatches")
sys.exit(1 if bad else 0)
