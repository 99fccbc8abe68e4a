"""Script-level integration: the drop-in infer_pseudo_masks.py / segmentation_test.py driven with a stub backbone and a tiny
synthetic dataset, their file outputs and reports compared with the reference procedure restated in oracle/."""
import logging
import os
import types

import numpy as np
import pytest
import torch
from PIL import Image

from oracle import confusion as oconf
from oracle import fuse as ofuse
from oracle import stitch as ostitch
from oracle import tta as otta

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

pytestmark = pytest.mark.gpu


class StubBackbone(torch.nn.Module):
    """Per-pixel affine map of the image + a fixed position-dependent bias: NOT equivariant under flips / rotations
    (so the d4 merge matters) and built from single-rounding elementwise ops (bit-identical on CPU and GPU)."""

    def __init__(self, C, size, seed):
        super().__init__()
        g = torch.Generator().manual_seed(seed)
        self.register_buffer("a", torch.randn(C, generator=g))
        self.register_buffer("bias", torch.randn((C, size, size), generator=g) * 2)
        self.C = C

    def forward(self, x):
        ch = torch.stack([x[:, c % 3] for c in range(self.C)], 1)
        return ch * self.a.view(1, -1, 1, 1) + self.bias.unsqueeze(0)


class PseudoDataset(torch.utils.data.Dataset):
    def __init__(self, n, size, labels, seed):
        g = torch.Generator().manual_seed(seed)
        self.images = torch.randn((n, 3, size, size), generator=g)
        self.tissue = torch.where(torch.rand((n, size, size), generator=g) < 0.2, 0.0, 127.0).double()
        self.names = [f"{1000 + i}-{i}-{i}-{labels[i % len(labels)]}.png" for i in range(n)]

    def __len__(self):
        return len(self.names)

    def __getitem__(self, i):
        return {"image": self.images[i], "tissue": self.tissue[i], "name": self.names[i]}


def test_infer_pseudo_masks_outputs(cuda, tmp_path):
    import infer_pseudo_masks as script
    size, n, C = 64, 10, 3
    labels = ["[1, 1, 0]", "[0, 1, 0]", "[1, 1, 1]", "[1, 0, 1]"]
    ds = PseudoDataset(n, size, labels, seed=3)
    model = StubBackbone(C, size, seed=4)
    orig = {name: (50 + 3 * i, 40 + 2 * i) for i, name in enumerate(ds.names)}   # (w, h) of the "original" tiles
    args = types.SimpleNamespace(checkpoint=None, train_data=str(tmp_path), save_dir=str(tmp_path / "pmask"), gpus=0, dataset="wsss4luad",
                                 batch_size=4, num_workers=0, pin_memory=False, patch_size=size)
    script.main(args, model=model.to(cuda), dataset=ds, original_size=lambda name: orig[name])
    for d in ("mask", "logits_32x32", "background-img", "entropy"):
        assert (tmp_path / "pmask" / d).is_dir()
    # reference procedure on the CPU (infer_pseudo_masks.py:118-154)
    merged = otta.d4_merge_mean(model.cpu(), ds.images)
    agree, total = 0, 0
    for i, name in enumerate(ds.names):
        low = torch.load(tmp_path / "pmask" / "logits_32x32" / (name.split(".png")[0] + ".pt"), map_location="cpu")
        ref_low = ofuse.lowres_32(merged[i:i + 1], literal=True)[0]
        assert low.shape == (C, 32, 32) and low.dtype == torch.float32
        assert float(((low - ref_low).abs() / ref_low.abs().clamp_min(1)).max()) <= 1e-5
        lab = script.label_from_name(name, "wsss4luad")
        mask, _ = ofuse.get_mask_pred_and_entropy(merged[i].clone(), ds.tissue[i].numpy(), lab)
        ref_png = Image.fromarray(np.uint8(mask), mode="P").resize(orig[name], resample=Image.BILINEAR)
        got = Image.open(tmp_path / "pmask" / "mask" / name)
        assert got.mode == "P" and got.size == orig[name]
        assert got.getpalette()[:12] == [0, 64, 128, 64, 128, 0, 243, 152, 0, 255, 255, 255]
        a, b = np.array(got), np.array(ref_png)
        agree += (a == b).sum(); total += a.size
    assert agree / total >= 0.9999


class PatchDataset(torch.utils.data.Dataset):
    """Tiles of two images cut at 3 scales with stride P/2, reflect-free (tiles are fully inside), named like
    split_validation.ipynb does."""

    def __init__(self, P, C, seed):
        g = torch.Generator().manual_seed(seed)
        self.sizes = {"00": (90, 120), "01": (75, 70)}   # (h, w)
        self.items = []
        for idx, (h, w) in self.sizes.items():
            for scale in (1.0, 1.25, 1.5):
                hs, ws = int(h * scale), int(w * scale)
                ys = sorted(set(list(range(0, max(hs - P, 0) + 1, P // 2)) + [max(hs - P, 0)]))
                xs = sorted(set(list(range(0, max(ws - P, 0) + 1, P // 2)) + [max(ws - P, 0)]))
                for y in ys:
                    for x in xs:
                        oh, ow = min(P, hs - y), min(P, ws - x)
                        img = torch.randn((3, P, P), generator=g)
                        msk = torch.randint(0, 4, (P, P), generator=g)
                        self.items.append((img, msk, f"{idx}_{scale}_{y}_{x}-[1, 1, 1].png", oh, ow))
        self.gt = {idx: torch.randint(0, 4, hw, generator=g).numpy().astype(np.uint8) for idx, hw in self.sizes.items()}

    def __len__(self):
        return len(self.items)

    def __getitem__(self, i):
        return self.items[i]


def test_segmentation_test_report(cuda, tmp_path, capsys):
    import segmentation_test as script
    P, C = 48, 3
    ds = PatchDataset(P, C, seed=11)
    model = StubBackbone(C, P, seed=12)
    args = types.SimpleNamespace(dataset="wsss4luad", checkpoint=str(tmp_path), patch_size=P, test_data=str(tmp_path / "patches"), batch_size=7,
                                 gpus=[0], num_workers=0, pin_memory=False, save_dir=str(tmp_path / "test"))
    test_iou, big_iou = script.main(args, model=model.to(cuda), dataset=ds, image_size=lambda i: (ds.sizes[i][1], ds.sizes[i][0]),
                                    load_gt=lambda i: ds.gt[i])
    out = capsys.readouterr().out
    assert "mIoU(big mask):" in out and "tIoU, sIoU, nIoU:" in out and "0/" in out
    # reference procedure on the CPU
    model = model.cpu()
    cm_patch = np.zeros((3, 3))
    tiles = {idx: [] for idx in ds.sizes}
    for img, msk, name, oh, ow in ds.items:
        logit = model(img[None])[0]
        cm_patch += oconf.generate_matrix(ofuse.miou_pred(logit[None]).numpy()[0], msk.numpy(), 3)
        idx, scale, pos = script.parse_tile_name(name)
        tiles[idx].append((logit, scale, pos, (oh, ow)))
    assert np.abs(test_iou.confusion_matrix - cm_patch).sum() <= 2e-4 * cm_patch.sum()
    cm_big = np.zeros((3, 3))
    for idx, (h, w) in ds.sizes.items():
        probs = ostitch.big_mask_fuse(tiles[idx], (h, w))
        pred, lab = ostitch.big_mask_labels(probs, ds.gt[idx])
        cm_big += oconf.generate_matrix(pred, ds.gt[idx], 3)
        got = np.array(Image.open(tmp_path / "test" / "mask" / f"{idx}.png"))
        assert got.shape == (h, w) and (got == lab).mean() >= 0.9999
    assert np.abs(big_iou.confusion_matrix - cm_big).sum() <= 2e-4 * cm_big.sum()
    assert abs(big_iou.Mean_Intersection_over_Union() - oconf.mean_iou(cm_big)) < 1e-3


def test_miou_mask_dropin_matches_reference_fixture(cuda):
    """pistoseg_b200.metrics.mIoUMask against the fixture produced by the reference's own loss.py."""
    from loss import mIoUMask
    z = np.load(os.path.join(os.path.dirname(__file__), "golden", "miou.npz"))
    for tag, C in (("luad", 3), ("bcss", 4), ("empty", 3)):
        m = mIoUMask(num_classes=C)
        r1 = m(torch.from_numpy(z[f"{tag}_logits1"]).to(cuda), torch.from_numpy(z[f"{tag}_mask1"].astype(np.int64)).to(cuda))
        assert np.allclose(np.array(r1), z[f"{tag}_ret1"], rtol=0, atol=1e-12)
        m(torch.softmax(torch.from_numpy(z[f"{tag}_logits2"]), 1).to(cuda), torch.from_numpy(z[f"{tag}_mask2"].astype(np.int64)).to(cuda), probs=True)
        assert np.array_equal(m.confusion_matrix, z[f"{tag}_cm"])
        assert np.array_equal(m.Tissue_Intersection_over_Union(), z[f"{tag}_tissue"])
        assert m.Mean_Intersection_over_Union() == float(z[f"{tag}_miou"])
        assert m.Frequency_Weighted_Intersection_over_Union() == float(z[f"{tag}_fwiou"])
        # numpy entry points of the reference API
        m.reset()
        pred = ofuse.miou_pred(torch.from_numpy(z[f"{tag}_logits1"])).numpy()
        m.add_batch(pred, z[f"{tag}_mask1"])
        assert np.array_equal(m.confusion_matrix, oconf.generate_matrix(pred, z[f"{tag}_mask1"], C).astype(np.float64))


def test_function_level_dropins(cuda):
    from pistoseg_b200 import postproc
    z = np.load(os.path.join(os.path.dirname(__file__), "golden", "pmask.npz"))
    i = 0
    while f"logit{i}" in z:
        logit = torch.from_numpy(z[f"logit{i}"].copy()).to(cuda)
        lab = [int(v) for v in z[f"label{i}"]]
        low = postproc.interpolate_tensor(logit, (8, 8))
        assert np.array_equal(low.cpu().numpy(), z[f"low{i}"])
        mask, ent = postproc.get_mask_pred_and_entropy(logit, z[f"tissue{i}"], lab)
        assert mask.dtype == np.int64 and (mask == z[f"mask{i}"]).mean() >= 0.9999
        assert np.abs(np.asarray(ent, np.float32) - z[f"entropy{i}"]).max() <= 2e-5
        assert np.array_equal(logit.cpu().numpy(), z[f"mutated{i}"]), "the reference's in-place -1e10 fill must be reproduced"
        i += 1


def test_segmentation_test_loads_a_lightning_style_checkpoint(cuda, tmp_path, monkeypatch):
    """The script's DEFAULT path (no injected model / dataset, run.sh:64): torch.load of a checkpoint whose
    hyper_parameters['args'] is an argparse.Namespace -- refused by torch >= 2.6 unless the script opts out of weights_only --
    then SegmentationModule.load_from_checkpoint and TestDataset from the (here: stand-in) reference modules."""
    import argparse
    import sys
    import segmentation_test as script
    P, C = 48, 3
    ds = PatchDataset(P, C, seed=21)
    ckpt_dir = tmp_path / "logs"
    ckpt_dir.mkdir()
    ckpt = ckpt_dir / "epoch=03-val_iou=0.5.ckpt"
    hp = argparse.Namespace(dataset="wsss4luad", lr=0.1, arch="stub")
    torch.save({"hyper_parameters": {"args": hp}, "state_dict": StubBackbone(C, P, seed=22).state_dict()}, ckpt)

    class SegmentationModule:
        @staticmethod
        def load_from_checkpoint(path, args=None):
            blob = torch.load(path, map_location="cpu")      # what Lightning does internally: default weights_only
            m = StubBackbone(C, P, seed=0)
            m.load_state_dict(blob["state_dict"])
            return m

    monkeypatch.setitem(sys.modules, "models", types.ModuleType("models"))
    mod = types.ModuleType("models.segmentation_module"); mod.SegmentationModule = SegmentationModule
    monkeypatch.setitem(sys.modules, "models.segmentation_module", mod)
    dmod = types.ModuleType("dataset"); dmod.TestDataset = lambda args: ds
    monkeypatch.setitem(sys.modules, "dataset", dmod)
    data = tmp_path / "data"
    (data / "img").mkdir(parents=True); (data / "mask").mkdir()
    for idx, (h, w) in ds.sizes.items():
        Image.fromarray(np.zeros((h, w, 3), np.uint8)).save(data / "img" / f"{idx}.png")
        Image.fromarray(ds.gt[idx]).save(data / "mask" / f"{idx}.png")
    args = argparse.Namespace(dataset="wsss4luad", checkpoint=str(ckpt), patch_size=P, test_data=str(data / "patches"), batch_size=7, gpus=[0],
                              num_workers=0, pin_memory=False, save_dir=str(ckpt_dir / "test"))
    test_iou, big_iou = script.main(args)
    assert big_iou.confusion_matrix.sum() == sum(int((g < 3).sum()) for g in ds.gt.values())
    assert (ckpt_dir / "test" / "mask" / "00.png").exists()


def test_all_reduce_with_an_empty_shard_creates_the_accumulator(cuda):
    """A rank whose shard is empty must still own a matrix to put into the collective (metrics.mIoUMask.all_reduce)."""
    from pistoseg_b200.metrics import mIoUMask
    m = mIoUMask(num_classes=3, device=cuda)
    assert m._conf is None
    m.all_reduce()                      # not distributed here: must be a no-op that does not fail
    assert float(m.confusion_matrix.sum()) == 0.0


def test_stitch_rejects_crops_larger_than_the_tile(cuda):
    from pistoseg_b200 import ops
    from pistoseg_b200._lib import PistoError
    tiles = torch.zeros((2, 3, 8, 8), device=cuda)
    canvas = torch.zeros((3, 16, 16), dtype=torch.float64, device=cuda); count = torch.zeros((16, 16), dtype=torch.float64, device=cuda)
    for bad in ([[0, 0, 9, 8], [0, 0, 8, 8]], [[0, 0, 8, 12], [0, 0, 8, 8]], [[-1, 0, 8, 8], [0, 0, 8, 8]], [[0, 0, 8, 8]]):
        with pytest.raises(PistoError):
            ops.stitch_accumulate(tiles, bad, canvas, count)
    ops.stitch_accumulate(tiles, [[0, 0, 8, 8], [8, 8, 5, 3]], canvas, count)



def test_oeem_prepare_seg_inputs_script(cuda, tmp_path):
    """OEEM/classification/prepare_seg_inputs.py (drop-in for the reference script of the same path): configuration file, output directory
    and file names are the reference's; every ``.npy`` equals the oracle's restatement of prepare_seg_inputs.py:96-137 on the same CAMs
    bit for bit.  Dataset and classifier are stubs with the reference modules' interfaces (TrainingSetCAM item layout, forward_cam)."""
    import importlib.util
    spec = importlib.util.spec_from_file_location("oeem_prepare_seg_inputs", os.path.join(ROOT, "OEEM", "classification", "prepare_seg_inputs.py"))
    script = importlib.util.module_from_spec(spec); spec.loader.exec_module(script)
    from pistoseg_b200 import oeem
    root = tmp_path
    (root / "classification").mkdir()
    (root / "classification" / "configuration_wsss4luad.yml").write_text(
        "---\nside_length: 224\nstride: 74\nnum_of_class: 3\nmean: [0.485, 0.456, 0.406]\nstd: [0.229, 0.224, 0.225]\nscales: [1, 1.5, 2]\nnetwork_image_size: 224\n")
    scales, side, stride = [1, 1.5, 2], 224, 74
    sizes = {"437-[1, 0, 1].png": (300, 260), "12-[0, 1, 1].png": (150, 400)}   # (rows, cols)

    class Stub(torch.utils.data.Dataset):   # dataset.TrainingSetCAM.__getitem__ (dataset.py:76-87)
        def __init__(self):
            self.files = sorted(sizes)
        def __len__(self):
            return len(self.files)
        def __getitem__(self, i):
            name = self.files[i]
            w, h = sizes[name]
            g = torch.Generator().manual_seed(i)
            ims, poss = [], []
            for s in scales:
                pos = oeem.online_cut_positions(int(w * s), int(h * s), side, stride)
                ims.append([torch.randn((3, 224, 224), generator=g) for _ in pos])
                poss.append([(np.int64(y), np.int64(x)) for y, x in pos])
            return name, ims, poss, scales, np.array([1, 0, 1])

    class Net(torch.nn.Module):             # network.wide_resnet.wideResNet.forward_cam: [n,3,224,224] -> [n,C,28,28]
        def __init__(self):
            super().__init__()
            self.mix = torch.nn.Conv2d(3, 3, 1)
        def forward_cam(self, x):
            return self.mix(torch.nn.functional.avg_pool2d(x, 8))

    torch.manual_seed(3)
    net = Net().to(cuda).eval()
    args = types.SimpleNamespace(batch=5, device=[cuda.index or 0], ckpt="res38d.pth", dataset="wsss4luad")
    n = script.main(args, net_cam=net, dset=Stub(), root=str(root), image_wh=lambda name: sizes[name], progress=False)
    assert n == 2
    out_dir = root / "classification" / "wsss4luad-res38d_train_pseudo_mask"
    assert sorted(os.listdir(out_dir)) == ["12-[0, 1, 1].npy", "437-[1, 0, 1].npy"]
    ds = Stub()
    for i, name in enumerate(ds.files):
        _, ims, poss, _, _ = ds[i]
        with torch.no_grad():
            cams = [torch.cat([net.forward_cam(b.to(cuda)) for b in torch.split(torch.stack(l), 5)]).cpu() for l in ims]
        pos = [[(int(y), int(x)) for y, x in p] for p in poss]
        ref = ostitch.cam_to_32(ostitch.cam_ensemble(cams, pos, scales, sizes[name], side=side))
        got = np.load(out_dir / (name[:-4] + ".npy"))
        assert got.dtype == np.float64 and got.shape == (3, 32, 32)
        assert np.array_equal(got, ref)
