#!/usr/bin/env python
"""Mode F (8 full-resolution d4 views, the literal infer_pseudo_masks.py:96,121 path): tiles/s and fraction of the HBM peak."""
import json, os, sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, ROOT)
from pistoseg_b200 import ops, synthetic, tta
from pistoseg_b200._lib import MASK_FILL
dev = torch.device("cuda:0")
g = torch.Generator().manual_seed(1)
N = int(sys.argv[1]) if len(sys.argv) > 1 else 256
CC = int(sys.argv[2]) if len(sys.argv) > 2 else 3
views = [torch.randn((N, CC, 224, 224), generator=g, dtype=torch.float32).to(dev) for _ in range(8)]
codes = [tta.deaug_code(h, a) for h, a in tta.aliases.d4_transform()]
bg = (torch.rand((N, 224, 224), generator=g) < 0.15).to(torch.uint8).to(dev)
pres = synthetic.make_present(N, CC, 5).to(dev)
fn = lambda: ops.fuse_argmax_confusion(views, codes, (224, 224), mask_mode=MASK_FILL, present=pres, bg=bg, bg_match=1, bg_label=CC, lowres=(32, 32))
for _ in range(3): fn()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
torch.cuda.synchronize(); e0.record()
for _ in range(10): fn()
e1.record(); torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / 10
b = 8 * CC * 224 * 224 * 4 + 2 * 224 * 224 + 4096 * CC
print(json.dumps({"C": CC, "modeF_tiles_per_s": N / ms * 1e3, "GB/s": N * b / ms / 1e6, "ms": ms}))
