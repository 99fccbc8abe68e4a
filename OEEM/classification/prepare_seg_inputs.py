#!/usr/bin/env python
"""OEEM stage ``prepare_seg_inputs`` -- multi-scale CAM ensemble -> 32 x 32 float64 CAMs per training image -- with everything after
the classifier on libpistoseg_b200 (``/root/reference`` layout: ``OEEM/classification/prepare_seg_inputs.py``, run from ``OEEM/``).

Command line, configuration file, input layout and output files are the reference's (``prepare_seg_inputs.py:26-79,138``):
  classification/configuration_<dataset>.yml            side_length, stride, scales, mean / std, network_image_size, num_of_class
  classification/weights/<dataset>/<ckpt>.pth           wideResNet checkpoint ('model' state dict with the 'module.' prefix)
  classification/<dataset>-<ckpt>_train_pseudo_mask/<stem>.npy    float64 [num_of_class, 32, 32]
The classifier (``network.wide_resnet``) and the tiling dataset (``dataset.TrainingSetCAM``) are the reference's own modules, imported
from the OEEM checkout this file is dropped into.  Per image the reference runs, on the host and per scale, an f32 interpolate of every
crop's CAM, a Python loop of float64 ``+=`` over the crops, a divide, a float64 interpolate, and a final float64 interpolate to
32 x 32 (``:96-137``); here the stride-8 CAMs never leave the GPU: ``pistoseg_b200.oeem.ensemble_32`` (f32 up-sampling, ordered
float64 overlap-add, normalise + resize + sum over scales in one pass, float64 resize to 32 x 32 -- bit-identical to the reference's
arithmetic) and one ``.npy`` per image written by a thread pool.
"""
import argparse
import importlib
import os
import sys

import numpy as np
import torch

_HERE = os.path.dirname(os.path.abspath(__file__))
for _p in (os.path.dirname(os.path.dirname(_HERE)), _HERE):   # the repo root (pistoseg_b200) and classification/ (dataset, network, utils)
    if _p not in sys.path:
        sys.path.insert(0, _p)

from pistoseg_b200 import io as pio   # noqa: E402
from pistoseg_b200 import oeem        # noqa: E402

for _k in ("OMP_NUM_THREADS", "OPENBLAS_NUM_THREADS", "MKL_NUM_THREADS", "VECLIB_MAXIMUM_THREADS", "NUMEXPR_NUM_THREADS"):
    os.environ.setdefault(_k, "2")    # the reference pins 2 host threads (prepare_seg_inputs.py:16-24)


def parse_args(argv=None):
    ap = argparse.ArgumentParser()
    ap.add_argument("-batch", default=20, type=int)
    ap.add_argument("-d", "--device", nargs="+", help="GPU id to use parallel", required=True, type=int)
    ap.add_argument("-ckpt", type=str, required=True, help="the checkpoint model name")
    ap.add_argument("-dataset", default="wsss4luad", help="the dataset name")
    return ap.parse_args(argv)


def data_path(dataset):
    """prepare_seg_inputs.py:54-59."""
    return {"glas": "classification/glas/1.training/img", "wsss4luad": "classification/WSSS4LUAD/1.training",
            "bcss": "classification/BCSS-WSSS/training"}[dataset]


def load_config(dataset, root="."):
    import yaml
    with open(os.path.join(root, f"classification/configuration_{dataset}.yml")) as f:
        return yaml.safe_load(f)


def build_dataset(cfg, dataset, root="."):
    """The reference's TrainingSetCAM with its transform (prepare_seg_inputs.py:61-66)."""
    from torchvision import transforms
    TrainingSetCAM = importlib.import_module("dataset").TrainingSetCAM
    size = cfg["network_image_size"]
    tf = transforms.Compose([transforms.Resize((size, size)), transforms.ToTensor(), transforms.Normalize(mean=cfg["mean"], std=cfg["std"])])
    return TrainingSetCAM(data_path_name=os.path.join(root, data_path(dataset)), transform=tf, patch_size=cfg["side_length"], stride=cfg["stride"],
                          scales=cfg["scales"], num_class=cfg["num_of_class"])


def build_classifier(cfg, dataset, ckpt, devices, root="."):
    """wideResNet with fc_cls copied into the 1 x 1 convolution fc_cam (prepare_seg_inputs.py:69-77)."""
    net_cam = getattr(importlib.import_module("network.wide_resnet"), "wideResNet")(num_class=cfg["num_of_class"])
    pretrained = torch.load(os.path.join(root, f"classification/weights/{dataset}/{ckpt}.pth"), map_location="cpu", weights_only=False)["model"]
    pretrained = {k[7:]: v for k, v in pretrained.items()}
    pretrained["fc_cam.weight"] = pretrained["fc_cls.weight"].unsqueeze(-1).unsqueeze(-1).to(torch.float64)
    pretrained["fc_cam.bias"] = pretrained["fc_cls.bias"]
    net_cam.load_state_dict(pretrained)
    net_cam.eval()
    return torch.nn.DataParallel(net_cam, device_ids=devices).cuda()


def image_cams(net_cam, scaled_im_list, batch_size, device):
    """forward_cam of every crop of every scale: list over scales of CUDA f32 [n_crops, C, h/8, w/8] (prepare_seg_inputs.py:107-117,
    without the interpolate and without the copy to the host)."""
    fwd = net_cam.module.forward_cam if hasattr(net_cam, "module") else net_cam.forward_cam
    out = []
    for im_list in scaled_im_list:
        ims = torch.vstack(list(im_list))
        out.append(torch.cat([fwd(b.to(device, non_blocking=True)).float() for b in torch.split(ims, batch_size)]))
    return out


def main(args, net_cam=None, dset=None, root=".", image_wh=None, progress=True):
    """``image_wh(name) -> (w, h)`` (rows, columns of the original image, as the reference names them) defaults to opening the file."""
    cfg = load_config(args.dataset, root)
    side, scales = cfg["side_length"], cfg["scales"]
    out_dir = os.path.join(root, f"classification/{args.dataset}-" + args.ckpt.replace(".pth", "") + "_train_pseudo_mask")
    os.makedirs(out_dir, exist_ok=True)
    device = torch.device("cuda", args.device[0])
    if dset is None:
        dset = build_dataset(cfg, args.dataset, root)
    if net_cam is None:
        net_cam = build_classifier(cfg, args.dataset, args.ckpt, args.device, root)
    if image_wh is None:
        from PIL import Image

        def image_wh(name):
            with Image.open(os.path.join(root, data_path(args.dataset), name)) as im:
                return im.size[1], im.size[0]          # np.asarray(Image).shape[:2]
    loader = torch.utils.data.DataLoader(dset, batch_size=1, drop_last=False)
    if progress:
        from tqdm import tqdm
        loader = tqdm(loader)
    writer = pio.AsyncWriter()
    n = 0
    with torch.no_grad():
        for im_name, scaled_im_list, scaled_position_list, _scales, _big_label in loader:
            name = im_name[0]
            w, h = image_wh(name)
            cams = image_cams(net_cam, scaled_im_list, args.batch, device)
            positions = [[(int(p[0]), int(p[1])) for p in plist] for plist in scaled_position_list]
            ens32 = oeem.ensemble_32(cams, positions, scales, (w, h), side=side)
            writer.submit(np.save, os.path.join(out_dir, ".".join(name.split(".")[:-1]) + ".npy"), ens32.cpu().numpy())
            n += 1
    writer.close()
    return n


if __name__ == "__main__":
    main(parse_args())
