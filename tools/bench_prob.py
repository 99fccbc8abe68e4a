import sys, torch
sys.path.insert(0, '.')
from pistoseg_b200 import ops, synthetic
from pistoseg_b200._lib import *
dev = torch.device('cuda:0')
cfg = synthetic.cfg2(N=1024); N = 8192
rep = lambda t: t.to(dev).repeat((8,) + (1,) * (t.dim() - 1)).contiguous()
views = [rep(v) for v in cfg['views']]; bg = rep(cfg['bg'])
fn = lambda: ops.fuse_argmax_confusion(views, cfg['codes'], (224, 224), fuse_mode=FUSE_PROB_MEAN, decide=DECIDE_RAW, bg=bg, bg_match=1, bg_label=3)
for _ in range(3): fn()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
torch.cuda.synchronize(); e0.record()
for _ in range(10): fn()
e1.record(); torch.cuda.synchronize()
print('PROB_MEAN Mtiles/s', N * 10 / e0.elapsed_time(e1) / 1e3)
