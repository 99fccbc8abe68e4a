#!/usr/bin/env python
"""Generates tests/golden/*.npz by RUNNING THE REFERENCE'S OWN CODE (and the libraries it calls) in the build container.

  python tests/golden/make_golden.py          (needs /root/reference; the fixtures it writes are committed)

What is executed:
  * /root/reference/loss.py is imported as a module (torch + numpy only)                     -> miou_*.npz
  * get_mask_pred_and_entropy / interpolate_tensor are cut out of /root/reference/infer_pseudo_masks.py with `ast`
    (the file's top-level imports need ttach / lightning, which are not installed) and executed  -> pmask_*.npz
  * torch.nn.functional.interpolate (what interpolate_tensor calls)                             -> bilinear.npz
  * cv2.flip / cv2.warpAffine / cv2.getRotationMatrix2D / cv2.copyMakeBorder through oracle.mosaic(use_cv2=True)
    (albumentations itself is not installed: the plan -> pixels contract is pinned, not its RNG)  -> mosaic.npz
"""
import ast
import importlib.util
import os
import sys

import numpy as np
import torch
import torch.nn.functional as F

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
REF = "/root/reference"


def load_ref_loss():
    spec = importlib.util.spec_from_file_location("ref_loss", os.path.join(REF, "loss.py"))
    m = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(m)
    return m


def load_ref_functions(path, names):
    src = open(path).read()
    tree = ast.parse(src)
    ns = {"torch": torch, "np": np, "F": F}
    for node in tree.body:
        if isinstance(node, ast.FunctionDef) and node.name in names:
            exec(compile(ast.Module([node], []), path, "exec"), ns)
    return ns


def golden_miou():
    loss = load_ref_loss()
    g = torch.Generator().manual_seed(101)
    out = {}
    for tag, C, B, H, extra in (("luad", 3, 4, 32, 1), ("bcss", 4, 3, 24, 1), ("empty", 3, 2, 16, 1)):
        m = loss.mIoUMask(num_classes=C)
        logits1 = torch.randn((B, C, H, H), generator=g) * 2
        logits2 = torch.randn((B, C, H, H), generator=g) * 2
        mask1 = torch.randint(0, C + extra, (B, H, H), generator=g)
        mask2 = torch.randint(0, C + extra, (B, H, H), generator=g)
        if tag == "empty":  # class 2 never occurs in gt nor pred -> IoU nan -> 0, freq 0
            logits1[:, 2] = -50; logits2[:, 2] = -50
            mask1[mask1 == 2] = 0; mask2[mask2 == 2] = 1
        r1 = m(logits1, mask1)
        r2 = m(torch.softmax(logits2, 1), mask2, probs=True)
        out[f"{tag}_logits1"] = logits1.numpy(); out[f"{tag}_logits2"] = logits2.numpy()
        out[f"{tag}_mask1"] = mask1.numpy().astype(np.uint8); out[f"{tag}_mask2"] = mask2.numpy().astype(np.uint8)
        out[f"{tag}_cm"] = m.confusion_matrix.copy()
        out[f"{tag}_ret1"] = np.array(r1); out[f"{tag}_ret2"] = np.array(r2)
        out[f"{tag}_tissue"] = m.Tissue_Intersection_over_Union()
        out[f"{tag}_miou"] = np.array(m.Mean_Intersection_over_Union())
        out[f"{tag}_fwiou"] = np.array(m.Frequency_Weighted_Intersection_over_Union())
    np.savez_compressed(os.path.join(HERE, "miou.npz"), **out)


def golden_pmask():
    ns = load_ref_functions(os.path.join(REF, "infer_pseudo_masks.py"), {"get_mask_pred_and_entropy", "interpolate_tensor"})
    fn, interp = ns["get_mask_pred_and_entropy"], ns["interpolate_tensor"]
    g = torch.Generator().manual_seed(202)
    out = {}
    labels = [[1, 1, 0], [0, 1, 0], [1, 1, 1], [1, 0, 1], [0, 0, 1], [1, 0, 0, 1], [0, 1, 0, 0]]
    for i, lab in enumerate(labels):
        Cn = len(lab)
        logit = torch.randn((Cn, 56, 56), generator=g) * 2
        if i == 2:
            logit = torch.round(logit * 2) / 2  # exact ties
        tissue = np.where(torch.rand((56, 56), generator=g).numpy() < 0.2, 0.0, 127.0)
        out[f"logit{i}"] = logit.numpy().copy()
        out[f"low{i}"] = interp(logit, (8, 8)).numpy()  # 56 -> 8: the odd-factor gather, like 224 -> 32
        mutated = logit.clone()
        mask, ent = fn(mutated, tissue, list(lab))
        out[f"label{i}"] = np.array(lab); out[f"tissue{i}"] = tissue
        out[f"mask{i}"] = np.asarray(mask).astype(np.int64); out[f"entropy{i}"] = np.asarray(ent, dtype=np.float32)
        out[f"mutated{i}"] = mutated.numpy()
    np.savez_compressed(os.path.join(HERE, "pmask.npz"), **out)


def golden_bilinear():
    g = torch.Generator().manual_seed(303)
    out = {}
    cases = [((1, 2, 28, 28), (224, 224), "f4"), ((1, 1, 21, 21), (224, 224), "f4"), ((1, 1, 35, 35), (224, 224), "f4"),
             ((1, 1, 224, 224), (32, 32), "f4"), ((1, 1, 112, 112), (100, 90), "f4"), ((1, 2, 26, 31), (103, 125), "f4"),
             ((1, 2, 75, 70), (60, 56), "f8"), ((1, 2, 40, 44), (100, 90), "f8")]
    torch.set_num_threads(8)
    for i, (shp, size, dt) in enumerate(cases):
        x = (torch.randn(shp, generator=g) * 3).to(torch.float32 if dt == "f4" else torch.float64)
        out[f"x{i}"] = x.numpy(); out[f"size{i}"] = np.array(size)
        out[f"y{i}"] = F.interpolate(x, size, mode="bilinear").numpy()
    out["n"] = np.array(len(cases))
    np.savez_compressed(os.path.join(HERE, "bilinear.npz"), **out)


def golden_mosaic():
    from oracle import mosaic, warp_affine as wa
    rng = np.random.default_rng(404)
    out = {}
    for tag, pn, ps, with_bg, ncls in (("luad", 4, 16, True, 3), ("bcss", 2, 32, False, 4)):
        P = 12
        sizes = [(ps * 2, ps * 2)] * 8 + [(ps - 5, ps + 3), (ps + 4, ps - 6), (ps - 3, ps - 2), (ps * 3, ps + 1)]
        pool = [rng.integers(0, 256, (h, w, 3), dtype=np.uint8) for h, w in sizes]
        bgs = [(rng.random(t.shape[:2]) < 0.15).astype(np.uint8) * 255 for t in pool] if with_bg else None
        labels = rng.integers(0, ncls, P).astype(np.uint8)
        H = W = pn * ps
        imgs, masks, plans, cells_all = [], [], [], []
        for trial in range(6):
            h = int(rng.integers(H // 10, H * 4 // 10 + 1)) * 2; w = int(rng.integers(W // 10, W * 4 // 10 + 1)) * 2
            qs = [(h, w), (h, W - w), (H - h, w), (H - h, W - w)]
            quads = []; cells = np.zeros((4, pn * pn, 3), np.int64)
            for q in range(4):
                for c in range(pn * pn):
                    t = int(rng.integers(0, P)); th, tw = pool[t].shape[:2]
                    cells[q, c] = (t, int(rng.integers(0, max(th, ps) - ps + 1)), int(rng.integers(0, max(tw, ps) - ps + 1)))
                warp = bool(rng.random() < 0.8) if trial else (q % 2 == 0)
                par = (float(rng.uniform(-45, 45)), float(rng.uniform(0.8, 1.2)), float(rng.uniform(-.0625, .0625)), float(rng.uniform(-.0625, .0625)))
                M = wa.shift_scale_rotate_matrix(H, W, *par) if warp else None
                quads.append(dict(flip=(q if trial == 0 else int(rng.integers(0, 4))), warp=warp, M=M,
                                  crop_y=int(rng.integers(0, H - qs[q][0] + 1)), crop_x=int(rng.integers(0, W - qs[q][1] + 1))))
            plan = dict(split_h=h, split_w=w, quads=quads)
            img, msk = mosaic.synthesize(plan, cells, pool, bgs, labels, pn, ps, use_cv2=True)
            imgs.append(img); masks.append(msk); cells_all.append(cells)
            plans.append(np.array([h, w] + [v for qd in quads for v in (qd["flip"], int(qd["warp"]), qd["crop_y"], qd["crop_x"])], np.int64))
            out[f"{tag}_M{trial}"] = np.stack([qd["M"] if qd["M"] is not None else np.zeros((2, 3)) for qd in quads])
        out[f"{tag}_img"] = np.stack(imgs); out[f"{tag}_mask"] = np.stack(masks)
        out[f"{tag}_plan"] = np.stack(plans); out[f"{tag}_cells"] = np.stack(cells_all)
        out[f"{tag}_labels"] = labels
        out[f"{tag}_pool_hw"] = np.array(sizes)
        out[f"{tag}_pool"] = np.concatenate([t.reshape(-1) for t in pool])
        if with_bg:
            out[f"{tag}_bg"] = np.concatenate([t.reshape(-1) for t in bgs])
    np.savez_compressed(os.path.join(HERE, "mosaic.npz"), **out)


if __name__ == "__main__":
    golden_miou(); golden_pmask(); golden_bilinear(); golden_mosaic()
    for f in sorted(os.listdir(HERE)):
        if f.endswith(".npz"):
            print(f, os.path.getsize(os.path.join(HERE, f)))
