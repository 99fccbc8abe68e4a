#!/usr/bin/env python
"""mIoUMask.forward shape (one full-resolution view + gt -> confusion, loss.py:55-67): tiles/s and HBM fraction."""
import json, os, sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, ROOT)
from pistoseg_b200 import ops
dev = torch.device("cuda:0")
g = torch.Generator().manual_seed(1)
NT = int(sys.argv[1]) if len(sys.argv) > 1 else 1024
x = torch.randn((NT, 3, 224, 224), generator=g).to(dev); gt = torch.randint(0, 4, (NT, 224, 224), generator=g, dtype=torch.uint8).to(dev)
conf = ops.new_confusion(3, dev)
fn = lambda: ops.fuse_argmax_confusion([x], [0], (224, 224), gt=gt, conf=conf, want_labels=False)
for _ in range(3): fn()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
torch.cuda.synchronize(); e0.record()
for _ in range(20): fn()
e1.record(); torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / 20
print(json.dumps({"identity_tiles_per_s": NT / ms * 1e3, "GB/s": NT * (3 * 224 * 224 * 4 + 224 * 224) / ms / 1e6}))
