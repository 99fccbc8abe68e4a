import sys, torch
sys.path.insert(0, '.')
from pistoseg_b200 import ops, synthetic, _lib
from pistoseg_b200._lib import DECIDE_SOFTMAX, MASK_FILL
dev = torch.device('cuda:0')
N = 8192
def rep(t, n):
    r = (n + t.shape[0] - 1) // t.shape[0]
    return t.to(dev).repeat((r,) + (1,) * (t.dim() - 1))[:n].contiguous()
cfg = synthetic.cfg2(N=1024, single_frac=0.0)
views = [rep(v, N) for v in cfg['views']]; bg = rep(cfg['bg'], N)
for name, present in (('varied', rep(cfg['present'], N)), ('110', torch.tensor([1,1,0], dtype=torch.uint8).repeat(N,1).to(dev)), ('101', torch.tensor([1,0,1], dtype=torch.uint8).repeat(N,1).to(dev)), ('011', torch.tensor([0,1,1], dtype=torch.uint8).repeat(N,1).to(dev))):
    fn = lambda: ops.fuse_argmax_confusion(views, cfg['codes'], (224,224), present=present, bg=bg, mask_mode=MASK_FILL, decide=DECIDE_SOFTMAX, bg_match=1, bg_label=3, lowres=(32,32))
    fn(); _lib.filter_stats(0, reset=True); fn(); torch.cuda.synchronize(); st = _lib.filter_stats(0, reset=True)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(10): fn()
    e1.record(); torch.cuda.synchronize()
    print(name, round(N*10/e0.elapsed_time(e1)/1e3, 3), 'Mtiles/s', st)
