#!/usr/bin/env python
"""Small invocations of every hot kernel (few tiles, all code paths, every CTA takes more than one tile): the input of a
compute-sanitizer memcheck / racecheck run where the tool is available (it is closed on the build pool), and a quick
"does every path still launch" check otherwise.  usage: sanitize_case.py [cfg2|cfg1|cfg3|cfg5|modef|all]"""
import os, sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, ROOT)
from pistoseg_b200 import ops, synthetic, tta
from pistoseg_b200._lib import DECIDE_SOFTMAX, MASK_FILL
dev = torch.device("cuda:0")
which = sys.argv[1] if len(sys.argv) > 1 else "all"
N = 300   # > 148 CTAs: every CTA takes more than one tile (buffer reuse, scheduler hand-over)
if which in ("all", "cfg2"):
    c = synthetic.cfg2(N=N)
    ops.fuse_argmax_confusion([v.to(dev) for v in c["views"]], c["codes"], (224, 224), mask_mode=MASK_FILL, decide=DECIDE_SOFTMAX,
                              present=c["present"], bg=c["bg"], bg_match=1, bg_label=3, lowres=(32, 32))
if which in ("all", "cfg1"):
    c = synthetic.cfg1(N=N)
    ops.fuse_argmax_confusion([v.to(dev) for v in c["views"]], c["codes"], (224, 224), decide=DECIDE_SOFTMAX, bg=c["bg"], gt=c["gt"],
                              bg_match=1, bg_label=3, conf=ops.new_confusion(3, dev))
if which in ("all", "cfg3"):
    c = synthetic.cfg3(N=N)
    ops.fuse_argmax_confusion([v.to(dev) for v in c["views"]], c["codes"], (224, 224), decide=DECIDE_SOFTMAX, gt=c["gt"], conf=ops.new_confusion(4, dev))
if which in ("all", "cfg5"):
    c = synthetic.cfg5(N=2, T=512)
    ops.fuse_argmax_confusion([v.to(dev) for v in c["views"]], c["codes"], (512, 512), decide=DECIDE_SOFTMAX, gt=c["gt"], conf=ops.new_confusion(4, dev))
if which in ("all", "modef"):
    g = torch.Generator().manual_seed(1)
    views = [torch.randn((40, 3, 224, 224), generator=g).to(dev) for _ in range(8)]
    codes = [tta.deaug_code(h, a) for h, a in tta.aliases.d4_transform()]
    bg = (torch.rand((40, 224, 224), generator=g) < 0.15).to(torch.uint8).to(dev)
    ops.fuse_argmax_confusion(views, codes, (224, 224), mask_mode=MASK_FILL, present=synthetic.make_present(40, 3, 5).to(dev), bg=bg, bg_match=1, bg_label=3, lowres=(32, 32))
torch.cuda.synchronize()
print("done", which)
