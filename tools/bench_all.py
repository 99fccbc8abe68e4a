#!/usr/bin/env python
"""Secondary measurements (one JSON object per line -> gpurun_out/bench_all.jsonl): every BASELINE config and every kernel
of the library, device-resident, CUDA events, inputs larger than L2.  Not the contract line (that is bench.py)."""
import json
import os
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from pistoseg_b200 import _lib, mosaic, ops, synthetic  # noqa: E402
from pistoseg_b200._lib import DECIDE_RAW, DECIDE_SOFTMAX, FUSE_PROB_MEAN, IMPL_GENERIC, MASK_FILL  # noqa: E402

dev = torch.device("cuda:0")
PEAK = 6535.4
try:
    PEAK = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"]
except Exception:
    pass
out_path = os.path.join(ROOT, "gpurun_out", "bench_all.jsonl")
os.makedirs(os.path.dirname(out_path), exist_ok=True)
fout = open(out_path, "w")


def timeit(fn, steps=20, warmup=3):
    for _ in range(warmup):
        fn()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize(); e0.record()
    for _ in range(steps):
        fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / steps


def emit(name, units, unit_name, ms, bytes_per_unit, **extra):
    rate = units / (ms * 1e-3)
    gbs = rate * bytes_per_unit / 1e9
    rec = dict(name=name, value=rate, unit=f"{unit_name}/s", ms=ms, bytes_per_unit=bytes_per_unit, achieved_gbs=gbs, hbm_frac=gbs / PEAK, **extra)
    print(json.dumps(rec)); fout.write(json.dumps(rec) + "\n"); fout.flush()


def rep(t, n):
    r = (n + t.shape[0] - 1) // t.shape[0]
    return t.to(dev).repeat((r,) + (1,) * (t.dim() - 1))[:n].contiguous()


def fuse_case(name, cfg, N, steps=20, **kw):
    views = [rep(v, N) for v in cfg["views"]]
    args = dict(present=rep(cfg["present"], N) if cfg.get("present") is not None else None,
                bg=rep(cfg["bg"], N) if cfg.get("bg") is not None else None,
                gt=rep(cfg["gt"], N) if cfg.get("gt") is not None else None)
    T, C = cfg["T"], cfg["C"]
    conf = ops.new_confusion(C, dev) if args["gt"] is not None else None
    fn = lambda: ops.fuse_argmax_confusion(views, cfg["codes"], (T, T), conf=conf, **args, **kw)
    ms = timeit(fn, steps)
    b = 4 * C * sum(v.shape[2] * v.shape[3] for v in views) + T * T * (1 + (args["bg"] is not None) + (args["gt"] is not None)) + (4 * C * 1024 if kw.get("lowres") else 0)
    emit(name, N, "tiles", ms, b, views=[int(v.shape[2]) for v in views])


# ---- fused path, every BASELINE config ------------------------------------------------------------------------------
fuse_case("cfg1 single stride-8 view, bg+gt+labels+confusion (C=3)", synthetic.cfg1(N=1024), 16384, decide=DECIDE_SOFTMAX, bg_match=1, bg_label=3)
fuse_case("cfg2 pseudo-mask: 3 scales x flip, present, bg, labels, 32x32 (C=3)", synthetic.cfg2(N=1024), 16384, mask_mode=MASK_FILL, decide=DECIDE_SOFTMAX, bg_match=1, bg_label=3, lowres=(32, 32))
fuse_case("cfg2 all multi-label tiles", synthetic.cfg2(N=1024, single_frac=0.0), 16384, mask_mode=MASK_FILL, decide=DECIDE_SOFTMAX, bg_match=1, bg_label=3, lowres=(32, 32))
# data dependence of the filtered kernel: the same cfg-2 call on other input families, with the fraction of
# pixels that went through the exact pass and the number of tiles evaluated exactly as a whole (pisto_filter_stats)
def family_case(family, N=16384):
    gen = torch.Generator(device=dev).manual_seed(1234)
    sizes, codes = [21, 21, 28, 28, 35, 35], [0, 4, 0, 4, 0, 4]
    views = synthetic.family_views(family, N, sizes, gen, dev)
    base = synthetic.cfg2(N=1024)
    present, bg = rep(base["present"], N), rep(base["bg"], N)
    fn = lambda: ops.fuse_argmax_confusion(views, codes, (224, 224), mask_mode=MASK_FILL, decide=DECIDE_SOFTMAX, present=present, bg=bg, bg_match=1,
                                           bg_label=3, lowres=(32, 32))
    fn(); _lib.filter_stats(0, reset=True)
    fn(); st = _lib.filter_stats(0, reset=True)
    ms = timeit(fn, 20)
    emit(f"cfg2 on '{family}' logits (40% single-label tiles)", N, "tiles", ms, 171440,
         exact_pixel_frac=st["exact_pixels"] / max(st["multi_tiles"] * 224 * 224, 1), whole_tile_exact_frac=st["exact_tiles"] / max(st["multi_tiles"], 1),
         multi_tiles=st["multi_tiles"])


for fam in ("gauss", "smooth", "quantized", "neartie", "extreme"):
    family_case(fam)
fuse_case("cfg3 BCSS: 3 scales x flip, gt, labels, confusion (C=4)", synthetic.cfg3(N=1000), 10000, decide=DECIDE_SOFTMAX)
fuse_case("cfg2 PROB_MEAN fusion (softmax per view)", synthetic.cfg2(N=1024), 8192, fuse_mode=FUSE_PROB_MEAN, decide=DECIDE_RAW, bg_match=1, bg_label=3)
for T, N in ((512, 512), (1024, 128), (2048, 32)):
    fuse_case(f"cfg5 large tile T={T}: 5 scales x flip, gt, labels, confusion (C=4)", synthetic.cfg5(N=8, T=T), N, steps=5, decide=DECIDE_SOFTMAX)
# reference-literal Mode F: 8 full-resolution d4 views
from pistoseg_b200 import tta  # noqa: E402
g = torch.Generator().manual_seed(1)
N = 256
views = [torch.randn((N, 3, 224, 224), generator=g, dtype=torch.float32).to(dev) for _ in range(8)]
codes = [tta.deaug_code(h, a) for h, a in tta.aliases.d4_transform()]
bg = (torch.rand((N, 224, 224), generator=g) < 0.15).to(torch.uint8).to(dev)
pres = synthetic.make_present(N, 3, 5).to(dev)
ms = timeit(lambda: ops.fuse_argmax_confusion(views, codes, (224, 224), mask_mode=MASK_FILL, present=pres, bg=bg, bg_match=1, bg_label=3, lowres=(32, 32)), 5)
emit("modeF: 8 full-res d4 views (ttach literal), present, bg, labels, 32x32", N, "tiles", ms, 8 * 3 * 224 * 224 * 4 + 2 * 224 * 224 + 12288)
# mIoUMask.forward shape: single full-res view + gt -> confusion (loss.py:55-67)
x = torch.randn((1024, 3, 224, 224), generator=g).to(dev); gt = torch.randint(0, 4, (1024, 224, 224), generator=g, dtype=torch.uint8).to(dev)
conf = ops.new_confusion(3, dev)
ms = timeit(lambda: ops.fuse_argmax_confusion([x], [0], (224, 224), gt=gt, conf=conf, want_labels=False), 10)
emit("mIoUMask.forward: softmax+argmax+confusion of full-res logits (C=3)", 1024, "tiles", ms, 3 * 224 * 224 * 4 + 224 * 224)

# ---- confusion kernel -----------------------------------------------------------------------------------------------
n = 10_000 * 224 * 224
pred = torch.randint(0, 4, (n,), dtype=torch.uint8, device=dev); gtb = torch.randint(0, 5, (n,), dtype=torch.uint8, device=dev)
conf = ops.new_confusion(4, dev)
ms = timeit(lambda: ops.confusion_accumulate(pred, gtb, conf), 20)
emit("confusion_accumulate 10k BCSS tiles (C=4)", 10000, "tiles", ms, 2 * 224 * 224)
del pred, gtb
# ---- bilinear resize ------------------------------------------------------------------------------------------------
x = torch.randn((8192, 3, 28, 28), device=dev)
ms = timeit(lambda: ops.upsample_bilinear(x, (224, 224)), 10)
emit("upsample_bilinear f32 28->224 (C=3)", 8192, "tiles", ms, 3 * 4 * (28 * 28 + 224 * 224))
x64 = torch.randn((3, 2000, 2500), device=dev, dtype=torch.float64)
ms = timeit(lambda: ops.upsample_bilinear(x64, (1600, 2000)), 10)
emit("upsample_bilinear f64 2000x2500 -> 1600x2000 (C=3)", 1, "images", ms, 3 * 8 * (2000 * 2500 + 1600 * 2000))
# ---- big-mask stitch (segmentation_test.py:141-215) and OEEM CAM ensemble (prepare_seg_inputs.py:96-138) -------------------
from pistoseg_b200 import stitch  # noqa: E402
H_, W_ = 2000, 2500
scales_ = (0.75, 1.0, 1.25, 1.5, 1.75)
tiles_, poss_, crops_ = {}, {}, {}
tile_bytes = 0
for sc in scales_:
    hs, ws = int(H_ * sc), int(W_ * sc)
    ys = list(range(0, max(hs - 224, 0) + 1, 112)); xs = list(range(0, max(ws - 224, 0) + 1, 112))
    if ys[-1] != hs - 224: ys.append(hs - 224)
    if xs[-1] != ws - 224: xs.append(ws - 224)
    pos = [(y, x) for y in ys for x in xs]
    tiles_[sc] = torch.randn((len(pos), 3, 224, 224), generator=g).to(dev); crops_[sc] = None
    poss_[sc] = torch.tensor([[y, x, 224, 224] for y, x in pos], dtype=torch.int32, device=dev)   # fixed tile grid: built once
    tile_bytes += tiles_[sc].numel() * 4
gt_ = torch.randint(0, 4, (H_, W_), generator=g, dtype=torch.uint8).to(dev)


def big_mask():
    f = stitch.BigMaskFuser((H_, W_), 3, dev)
    for sc in scales_:
        f.add_tiles(tiles_[sc], sc, poss_[sc], crops_[sc])
    conf = ops.new_confusion(3, dev)
    return f.finish(gt=gt_, conf=conf)


ms = timeit(big_mask, 3, 1)
canvas_bytes = sum(int(H_ * sc) * int(W_ * sc) for sc in scales_) * 8 * 4 * 3 + H_ * W_ * 8 * 3 * (2 * len(scales_) + 1)
emit("big-mask fusion 2000x2500, 5 scales, 50% overlapping 224 tiles: softmax + f64 overlap-add + f64 resize + argmax + confusion", 1, "images", ms,
     tile_bytes + canvas_bytes, tiles=sum(len(v) for v in poss_.values()))
del tiles_
# ---- mosaic ---------------------------------------------------------------------------------------------------------
rng = np.random.default_rng(0)
P = 2048
imgs = [rng.integers(0, 256, (224, 224, 3), dtype=np.uint8) for _ in range(64)]
imgs = [imgs[i % 64] for i in range(P)]
bgs = [(rng.random((224, 224)) < 0.1).astype(np.uint8) * 255 for _ in range(64)]
bgs = [bgs[i % 64] for i in range(P)]
pool = mosaic.TilePool(imgs, rng.integers(0, 3, P).astype(np.uint8), bgs, device=dev)
for pn, ps in ((4, 56), (2, 112), (7, 32)):
    planner = mosaic.MosaicPlanner(pool, pn, ps, reject_bg=True)
    N = 8192
    t0 = time.perf_counter(); plans = planner.quad_plans(range(N)); quad_s = time.perf_counter() - t0
    pool.integral_device(ps)
    cells = planner.cells_device(0, 1, N)
    ms_cells = timeit(lambda: planner.cells_device(0, 1, N), 5)
    pl = torch.from_numpy(plans.view(np.uint8).reshape(-1)).to(dev)
    ms = timeit(lambda: ops.mosaic_gather(pool.dev, pl, cells, pn, ps, 3), 5)
    emit(f"mosaic_gather {pn}x{ps} (224x224 image+mask, 80% warped quadrants)", N, "mosaics", ms, 2 * (224 * 224 * 4),
         plan_cells_device_us_per_mosaic=ms_cells / N * 1e3, plan_quads_host_us_per_mosaic=quad_s / N * 1e6,
         plan_quads_device_us_per_mosaic=timeit(lambda: planner.quads_device(0, 1, N), 5) / N * 1e3)
    torch.cuda.synchronize(); t0 = time.perf_counter()
    for r in range(3):
        mosaic.synthesize_range(pool, planner, r * N, 1, N)
    torch.cuda.synchronize(); dt = (time.perf_counter() - t0) / 3
    emit(f"mosaic end to end {pn}x{ps}: device quadrant plans + device cell plans + gather (wall clock)", N, "mosaics", dt * 1e3, 2 * (224 * 224 * 4))
fout.close()
