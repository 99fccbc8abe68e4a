#!/usr/bin/env python
"""Stage 6 of the PistoSeg pipeline -- segmentation test + mIoU report -- with the post-processing on libpistoseg_b200.

Same command line, same ``<ckpt>/test/mask/*.png`` outputs, same ``segmentation_test.log`` lines and stdout report as the
reference script (``run.sh:64``).  Patch-level softmax/argmax/confusion is one fused kernel per batch; the WSSS4LUAD
big-mask path (multi-scale stitching in float64, resize, average, argmax, confusion) runs on device canvases
(pistoseg_b200/stitch.py).  Under ``torchrun`` tiles are sharded by image over the ranks and the confusion matrices are
merged with one all-reduce.
"""
import argparse
import logging
import os
import sys
from pathlib import Path

import numpy as np
import torch

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

from pistoseg_b200 import _lib, dist as pdist, io as pio, ops
from pistoseg_b200.metrics import mIoUMask
from pistoseg_b200.stitch import BigMaskFuser

for _k in ("OMP_NUM_THREADS", "OPENBLAS_NUM_THREADS", "MKL_NUM_THREADS", "VECLIB_MAXIMUM_THREADS", "NUMEXPR_NUM_THREADS"):
    os.environ.setdefault(_k, "2")


def parse_args(argv=None):
    ap = argparse.ArgumentParser()
    ap.add_argument("--dataset", type=str, default="wsss4luad")
    ap.add_argument("--checkpoint", "-ckpt", help="path to the checkpoint file")
    ap.add_argument("--patch-size", type=int, default=256)
    ap.add_argument("--test-data", default="./data/testing")
    ap.add_argument("--batch-size", type=int, default=64)
    ap.add_argument("--gpus", default=[1, ])
    ap.add_argument("--num-workers", type=int, default=8)
    ap.add_argument("--pin-memory", action="store_true", default=True)
    return ap.parse_args(argv)


def parse_tile_name(name):
    """{img}_{scale}_{y}_{x}-{label}.png (split_validation.ipynb; segmentation_test.py:160-162)."""
    parts = name.split("_")
    return parts[0], float(parts[1]), (int(parts[2]), int(parts[3].split("-")[0]))


def report(test_iou, big_iou, dataset):
    if dataset == "wsss4luad":
        logging.critical(f"Segmentation Test - Test mIoU (patch): {test_iou.Mean_Intersection_over_Union()}")
        logging.critical(f"Segmentation Test - Test fwIoU (patch): {test_iou.Frequency_Weighted_Intersection_over_Union()}")
        logging.critical(f"Segmentation Test - Test tissue IoU (patch): {test_iou.Tissue_Intersection_over_Union()}")
        print(f"mIoU(big mask): {big_iou.Mean_Intersection_over_Union()}")
        print(f"fwIoU: {big_iou.Frequency_Weighted_Intersection_over_Union()}")
        print(f"tIoU, sIoU, nIoU: {big_iou.Tissue_Intersection_over_Union()}")
        logging.critical(f"Segmentation Test - Test mIoU (big mask): {big_iou.Mean_Intersection_over_Union()}")
        logging.critical(f"Segmentation Test - Test fwIoU (big mask): {big_iou.Frequency_Weighted_Intersection_over_Union()}")
        logging.critical(f"MosaSegmentationic Test - Test tissue IoU (big mask): {big_iou.Tissue_Intersection_over_Union()}")  # sic (:227)
    else:
        print(f"mIoU(big mask): {test_iou.Mean_Intersection_over_Union()}")
        print(f"fwIoU: {test_iou.Frequency_Weighted_Intersection_over_Union()}")
        print(f"tmr, str, lym, nec: {test_iou.Tissue_Intersection_over_Union()}")
        logging.critical(f"Segmentation Test - Test mIoU (big mask): {test_iou.Mean_Intersection_over_Union()}")
        logging.critical(f"Segmentation Test - Test fwIoU (big mask): {test_iou.Frequency_Weighted_Intersection_over_Union()}")
        logging.critical(f"Segmentation Test - Test tissue IoU (big mask): {test_iou.Tissue_Intersection_over_Union()}")


def _trust_local_checkpoints():
    """torch >= 2.6 unpickles with weights_only=True by default and refuses the argparse.Namespace a Lightning checkpoint of this
    project carries (hyper_parameters['args']).  The checkpoints are the user's own training output, as in the reference
    (segmentation_test.py:94,106): allow-list the Namespace so that Lightning's internal torch.load works too."""
    import argparse
    try:
        torch.serialization.add_safe_globals([argparse.Namespace])
    except AttributeError:  # older torch: nothing to do
        pass


def main(args, model=None, dataset=None, image_size=None, load_gt=None):
    """model / dataset / image_size(image_idx)->(w,h) / load_gt(image_idx)->uint8 [h,w] can be injected; the defaults
    are the reference's SegmentationModule checkpoint, TestDataset and the PNGs next to the patch directory."""
    import torch.distributed as dist
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    device = torch.device("cuda", int(os.environ.get("LOCAL_RANK", "0")))
    if world > 1 and not dist.is_initialized():
        dist.init_process_group("nccl", device_id=device)
    if model is None:
        _trust_local_checkpoints()
        lib = torch.load(args.checkpoint, map_location="cpu", weights_only=False)  # a Lightning checkpoint: hyper_parameters['args'] is an argparse.Namespace
        model_args = lib["hyper_parameters"]["args"]
        for k, v in vars(args).items():
            setattr(model_args, k, v)
        args = model_args
        from models.segmentation_module import SegmentationModule   # reference module
        model = SegmentationModule.load_from_checkpoint(args.checkpoint, args=args).to(device)
    if dataset is None:
        from dataset import TestDataset                               # reference module
        dataset = TestDataset(args)
    from PIL import Image
    data_root = "/".join(str(args.test_data).split("/")[:-1])
    if image_size is None:
        def image_size(idx):
            return Image.open(os.path.join(data_root, "img", idx + ".png")).size
    if load_gt is None:
        def load_gt(idx):
            return np.array(Image.open(os.path.join(data_root, "mask", idx + ".png")))

    luad = args.dataset == "wsss4luad"
    C = 3 if luad else 4
    test_iou = mIoUMask(num_classes=C)
    print(f"Save dir: {args.save_dir}")
    os.makedirs(os.path.join(args.save_dir, "mask"), exist_ok=True)
    # tiles of one image must meet on one rank (they share a canvas): shard the dataset by image index
    indices = list(range(len(dataset)))
    if world > 1:
        keys = [str(dataset.test_image[i].name).split("_")[0] if hasattr(dataset, "test_image") else str(i) for i in indices]
        indices = pdist.shard_by_key(keys, rank, world)
        dataset = torch.utils.data.Subset(dataset, indices)
    loader = torch.utils.data.DataLoader(dataset, batch_size=args.batch_size, num_workers=args.num_workers, pin_memory=args.pin_memory, shuffle=False)
    fusers = {}
    writer = pio.AsyncWriter()
    palette = pio.palette_for(args.dataset)
    model.eval()
    with torch.no_grad():
        for i, data in enumerate(loader):
            image_batch, mask_batch, name_batch, oh_batch, ow_batch = data
            if i % 100 == 0:
                print(f"{i}/{len(loader)}")
                print(f"mIoU(patch): {test_iou.Mean_Intersection_over_Union()}")
                print(f"fwIoU: {test_iou.Frequency_Weighted_Intersection_over_Union()}")
                print(f"tIoU, sIoU, nIoU: {test_iou.Tissue_Intersection_over_Union()}")
            output = model(image_batch.to(device, non_blocking=True)).float().contiguous()
            if luad:
                test_iou.update(output, mask_batch)                     # softmax + argmax + confusion, one kernel
                groups = {}
                for j, name in enumerate(name_batch):
                    idx, scale, pos = parse_tile_name(name)
                    groups.setdefault((idx, scale), []).append((j, pos, (int(oh_batch[j]), int(ow_batch[j]))))
                for (idx, scale), items in groups.items():
                    if idx not in fusers:
                        w, h = image_size(idx)
                        fusers[idx] = BigMaskFuser((h, w), 3, device)
                    sel = torch.tensor([t[0] for t in items], device=device)
                    fusers[idx].add_tiles(output.index_select(0, sel), scale, [t[1] for t in items], [t[2] for t in items])
            else:
                # one pass over the logits: softmax-argmax into the confusion matrix (:137-139) and logit-argmax for the PNG (:182)
                out = ops.fuse_argmax_confusion([output], [0], output.shape[-2:], decide=_lib.DECIDE_SOFTMAX, gt=mask_batch.to(device).byte(),
                                                conf=test_iou._acc(device), want_labels=False, want_raw_labels=True)
                raw = out["labels_raw"].cpu().numpy()
                for j, name in enumerate(name_batch):
                    writer.submit(pio.save_mask_png, raw[j], os.path.join(args.save_dir, "mask", name), palette)
                del out
    big_iou = mIoUMask(num_classes=3, device=device)
    if luad:
        conf = big_iou._acc(device)
        for idx, fuser in fusers.items():
            gt = torch.from_numpy(load_gt(idx).astype(np.uint8)).to(device)
            res = fuser.finish(gt=gt, conf=conf, bg_match=3, bg_label=3)
            writer.submit(pio.save_mask_png, res["labels"].cpu().numpy(), os.path.join(args.save_dir, "mask", idx + ".png"), palette)
    test_iou.all_reduce(); big_iou.all_reduce()                     # one int64 [C,C] all-reduce each
    writer.close()
    if rank == 0:
        report(test_iou, big_iou, args.dataset)
    return test_iou, big_iou


if __name__ == "__main__":
    args = parse_args()
    try:
        import pytorch_lightning as pl
        pl.seed_everything(42)
    except ImportError:
        torch.manual_seed(42); np.random.seed(42)
    args.save_dir = os.path.join(args.checkpoint, "test")
    Path(args.save_dir).mkdir(exist_ok=True, parents=True)
    logging.basicConfig(level=logging.CRITICAL, filename=f"{args.checkpoint}/segmentation_test.log", filemode="w",
                        format="%(asctime)s - %(name)s - %(levelname)-9s - %(filename)-8s : %(lineno)s line - %(message)s - %(funcName)s",
                        datefmt="%Y-%m-%d %H:%M:%S")
    logging.critical(args)
    ckpt = None
    for filename in os.listdir(args.checkpoint):
        if "epoch=" in filename:
            ckpt = os.path.join(args.checkpoint, filename)
            break
    assert ckpt is not None, f"Cannot find a valid checkpoint file in {args.checkpoint}"
    args.checkpoint = ckpt
    print(f"Find best checkpoint file: {args.checkpoint}")
    main(args)
