#!/usr/bin/env python
"""CPU baselines of the other BASELINE configs: the oracle's literal torch / numpy / cv2 restatement of the
reference, timed on this box's host cores.  One JSON object per line (-> gpurun_out/cpu_baselines.jsonl).  Reported only.
  cfg 1 / 3 / 5: oracle/pipeline.py (+ oracle/confusion.py) on a bounded sample, all host threads and the reference's own 2
  cfg 4        : oracle/mosaic.py with the real cv2.flip / cv2.warpAffine, single process and a process pool over all cores
                 (the notebook uses Pool(12), create_dataset.ipynb:552-560)"""
import json, os, sys, time
import multiprocessing as mp
import numpy as np
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, ROOT)
from oracle import confusion as oconf, mosaic as omosaic, pipeline
from pistoseg_b200 import synthetic

out = open(os.path.join(ROOT, "gpurun_out", "cpu_baselines.jsonl"), "w") if os.path.isdir(os.path.join(ROOT, "gpurun_out")) else None


def emit(**kw):
    line = json.dumps(kw); print(line, flush=True)
    if out: out.write(line + "\n"); out.flush()


def fuse_case(name, cfg, n):
    T = cfg["T"]
    for threads in (os.cpu_count(), 2):
        torch.set_num_threads(threads)
        views = [v[:n] for v in cfg["views"]]
        pres = cfg["present"][:n].numpy() if cfg.get("present") is not None else None
        bg = cfg["bg"][:n].numpy() if cfg.get("bg") is not None else None
        best = None
        for _ in range(2):
            t0 = time.perf_counter()
            labels, _ = pipeline.pseudo_mask_batch(views, cfg["codes"], (T, T), pres, bg)
            if cfg.get("gt") is not None:
                oconf.generate_matrix(labels, cfg["gt"][:n].numpy(), cfg["C"])
            dt = time.perf_counter() - t0
            best = dt if best is None else min(best, dt)
        emit(name=name, unit="tiles/s", value=n / best, sample=f"{n} tiles, best of 2", threads=threads, cores=os.cpu_count(), kind="port")


_POOL = {}


def _one(args):
    plan, cells, pn, ps = args
    return omosaic.synthesize(plan, cells, _POOL["imgs"], _POOL["bgs"], _POOL["labels"], pn, ps, use_cv2=True)[1].sum()


def mosaic_case(pn, ps, n_single, n_pool):
    from pistoseg_b200 import mosaic
    rng = np.random.default_rng(0)
    P = 512
    imgs = [rng.integers(0, 256, (224, 224, 3), dtype=np.uint8) for _ in range(P)]
    bgs = [(rng.random((224, 224)) < 0.1).astype(np.uint8) * 255 for _ in range(P)]
    labels = rng.integers(0, 3, P).astype(np.uint8)
    pool = mosaic.TilePool(imgs, labels, bgs, device="cuda")
    planner = mosaic.MosaicPlanner(pool, pn, ps, seed=2022, reject_bg=True)
    plans, cells = planner.plans(range(n_pool))
    jobs = []
    for k in range(n_pool):
        quads = []
        for q in range(4):
            qd = plans[k]["quad"][q]
            M = None
            if qd["warp"]:
                A = np.vstack([np.asarray(qd["minv"], np.float64).reshape(2, 3), [0, 0, 1]])
                M = np.linalg.inv(A)[:2]
            quads.append(dict(flip=int(qd["flip"]), warp=bool(qd["warp"]), crop_y=int(qd["crop_y"]), crop_x=int(qd["crop_x"]), M=M))
        cl = np.stack([cells[k]["tile"], cells[k]["cy"], cells[k]["cx"]], -1).astype(np.int64)
        jobs.append((dict(split_h=int(plans[k]["split_h"]), split_w=int(plans[k]["split_w"]), quads=quads), cl, pn, ps))
    _POOL.update(imgs=imgs, bgs=bgs, labels=labels)
    t0 = time.perf_counter()
    for j in jobs[:n_single]:
        _one(j)
    dt = time.perf_counter() - t0
    emit(name=f"cfg4 mosaic {pn}x{ps} (cv2 flip / warpAffine, numpy gather)", unit="mosaics/s", value=n_single / dt, sample=f"{n_single} mosaics", processes=1, cores=os.cpu_count(), kind="port")
    ctx = mp.get_context("fork")
    with ctx.Pool(os.cpu_count()) as pl:
        pl.map(_one, jobs[:os.cpu_count()])  # warm
        t0 = time.perf_counter()
        pl.map(_one, jobs, chunksize=8)
        dt = time.perf_counter() - t0
    emit(name=f"cfg4 mosaic {pn}x{ps} (cv2 flip / warpAffine, numpy gather)", unit="mosaics/s", value=n_pool / dt, sample=f"{n_pool} mosaics", processes=os.cpu_count(), cores=os.cpu_count(), kind="port")


if __name__ == "__main__":
    fuse_case("cfg1 (V=1, bg + gt, confusion)", synthetic.cfg1(N=256), 256)
    fuse_case("cfg3 (C=4, V=6, gt, confusion)", synthetic.cfg3(N=256), 256)
    fuse_case("cfg5 T=512 (C=4, V=10, gt, confusion)", synthetic.cfg5(N=8, T=512), 8)
    mosaic_case(4, 56, 200, 2000)
