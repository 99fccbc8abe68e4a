"""Output writers of the pseudo-mask / test stages: palette PNGs, ``.pt`` logits.
These stay on the host (PIL / torch.save) but run in a small thread pool so that encoding overlaps the GPU."""
import os
from concurrent.futures import ThreadPoolExecutor

import numpy as np
import torch
from PIL import Image

PALETTE_WSSS4LUAD = [0, 64, 128, 64, 128, 0, 243, 152, 0, 255, 255, 255] + [0] * 252 * 3   # infer_pseudo_masks.py:143
PALETTE_BCSS = [255, 0, 0, 0, 255, 0, 0, 0, 255, 153, 0, 255, 255, 255, 255]              # infer_pseudo_masks.py:145-150


def palette_for(dataset):
    return PALETTE_WSSS4LUAD if dataset == "wsss4luad" else PALETTE_BCSS


def save_mask_png(mask_u8, path, palette, size_wh=None):
    """mode-'P' PNG; resize of a palette image is NEAREST whatever resample is asked (infer_pseudo_masks.py:151-153)."""
    im = Image.fromarray(np.ascontiguousarray(mask_u8, dtype=np.uint8), mode="P")
    im.putpalette(palette)
    if size_wh is not None and tuple(size_wh) != im.size:
        im = im.resize(tuple(size_wh), resample=Image.BILINEAR)
    im.save(path)


def save_rgb_png(img_u8, path):
    """RGB PNG of a mosaic image (create_dataset.ipynb:546)."""
    Image.fromarray(np.ascontiguousarray(img_u8, dtype=np.uint8), mode="RGB").save(path)


class AsyncWriter:
    def __init__(self, workers=8):
        self.pool = ThreadPoolExecutor(max_workers=workers)
        self.pending = []

    def submit(self, fn, *a, **k):
        self.pending.append(self.pool.submit(fn, *a, **k))
        if len(self.pending) > 4096:
            self.drain()

    def drain(self):
        for f in self.pending:
            f.result()
        self.pending = []

    def close(self):
        self.drain()
        self.pool.shutdown()


def save_logits_pt(t, path):
    torch.save(t, path)


def ensure_dirs(root, names):
    for n in names:
        os.makedirs(os.path.join(root, n), exist_ok=True)
