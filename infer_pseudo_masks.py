#!/usr/bin/env python
"""Stage 2 of the PistoSeg pipeline -- pseudo-mask inference -- with the post-processing on libpistoseg_b200.

Command line, output files and directory layout are those of the reference script (``run.sh:48``):
  <save-dir>/mask/<name>.png          mode-P PNG at the original tile size, labels 0..C-1, background = C
  <save-dir>/logits_32x32/<stem>.pt   float32 [C,32,32] TTA-merged logits
  <save-dir>/background-img/, <save-dir>/entropy/   (created, left empty, as in the reference)
The backbone (stage-1 ``MosaicModule`` checkpoint) and the dataset (``TrainDataset``) are the reference's own modules,
imported from the PistoSeg checkout this script is dropped into; everything after the backbone output of a batch is ONE
kernel launch (d4 merge, 32x32 export, label masking, softmax/argmax, background) instead of ~25 per tile.
"""
import argparse
import os
import sys
from pathlib import Path

import numpy as np
import torch

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

from pistoseg_b200 import io as pio
from pistoseg_b200 import postproc, tta

for _k in ("OMP_NUM_THREADS", "OPENBLAS_NUM_THREADS", "MKL_NUM_THREADS", "VECLIB_MAXIMUM_THREADS", "NUMEXPR_NUM_THREADS"):
    os.environ.setdefault(_k, "2")   # the reference pins 2 host threads (infer_pseudo_masks.py:22-28)


def parse_args(argv=None):
    ap = argparse.ArgumentParser(description="Infer Pseudo-Labels for dataset")
    ap.add_argument("--checkpoint", "-ckpt", help="the checkpoint path of stage 1 model")
    ap.add_argument("--train-data", default="./data/training")
    ap.add_argument("--save-dir", default="./pmask")
    ap.add_argument("--gpus", type=int)
    ap.add_argument("--dataset", type=str, default="wsss4luad")
    ap.add_argument("--batch-size", type=int, default=32)
    ap.add_argument("--num-workers", type=int, default=8)
    ap.add_argument("--pin-memory", action="store_true", default=False)
    ap.add_argument("--patch-size", type=int, default=256)
    return ap.parse_args(argv)


def label_from_name(name, dataset):
    """Image-level label vector encoded in the tile file name (utils.get_label / to_list; infer_pseudo_masks.py:130-134)."""
    if dataset == "wsss4luad":
        s = str(name).split("-")[-1].split(".")[0]          # "[1, 0, 1]"
        return [int(v) for v in s[1:-1].split(", ")]
    s = str(name).split("]")[0].split("[")[-1]              # "1010"
    return [int(s[0]), int(s[1]), int(s[2]), int(s[3])]


def find_checkpoint(ckpt_dir):
    for fn in os.listdir(ckpt_dir):
        if "epoch=" in fn:
            return os.path.join(ckpt_dir, fn)
    raise AssertionError(f"Cannot find a valid checkpoint file in {ckpt_dir}")


def main(args, model=None, dataset=None, original_size=None):
    """model / dataset / original_size can be injected (tests, other backbones); by default they are the reference's
    MosaicModule checkpoint, TrainDataset and PIL's view of the tile on disk."""
    device = torch.device("cuda", args.gpus if args.gpus is not None else 0)
    if model is None:
        from models.mosaic_module import MosaicModule     # reference module (backbone stays in PyTorch)
        import argparse
        try:  # torch >= 2.6: Lightning's internal torch.load runs with weights_only=True and must be allowed the checkpoint's Namespace
            torch.serialization.add_safe_globals([argparse.Namespace])
        except AttributeError:
            pass
        model = MosaicModule.load_from_checkpoint(args.checkpoint).cuda(device)
    if dataset is None:
        from dataset import TrainDataset                   # reference module
        dataset = TrainDataset(args)
    if original_size is None:
        from PIL import Image

        def original_size(name):
            return Image.open(Path(args.train_data) / name).size
    model = tta.SegmentationTTAWrapper(model, tta.aliases.d4_transform(), merge_mode="mean")
    gen = torch.Generator()
    gen.manual_seed(0)
    loader = torch.utils.data.DataLoader(dataset, batch_size=args.batch_size, num_workers=args.num_workers, pin_memory=args.pin_memory,
                                         shuffle=False, generator=gen)
    pio.ensure_dirs(args.save_dir, ["mask", "logits_32x32", "background-img", "entropy"])
    palette = pio.palette_for(args.dataset)
    writer = pio.AsyncWriter()
    model.eval()
    with torch.no_grad():
        for batch in loader:
            image = batch["image"].to(device, non_blocking=True)
            names = list(batch["name"])
            present = torch.tensor([label_from_name(n, args.dataset) for n in names], dtype=torch.uint8)
            tissue_is_bg = (batch["tissue"] == 0).to(torch.uint8)                 # dataset.py:85-88: 0 = background, 127 = tissue
            views, codes = model.views(image)                                      # 8 backbone forwards (PyTorch)
            out = postproc.pseudo_mask_batch(views, codes, image.shape[-2:], present, tissue_is_bg)   # 1 kernel
            labels = out["labels"].cpu().numpy()
            lowres = out["lowres"].cpu()
            for j, name in enumerate(names):
                writer.submit(pio.save_logits_pt, lowres[j].clone(), Path(args.save_dir) / "logits_32x32" / (name.split(".png")[0] + ".pt"))
                writer.submit(pio.save_mask_png, labels[j], Path(args.save_dir) / "mask" / name, palette, original_size(name))
    writer.close()


if __name__ == "__main__":
    args = parse_args()
    print("Saving to {}".format(args.save_dir))
    args.checkpoint = find_checkpoint(args.checkpoint)
    print(f"Loading checkpoint from {args.checkpoint}")
    try:
        import pytorch_lightning as pl
        pl.seed_everything(42, workers=True)
    except ImportError:
        torch.manual_seed(42); np.random.seed(42)
    main(args)
