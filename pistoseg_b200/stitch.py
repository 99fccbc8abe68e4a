"""Big-mask fusion of ``segmentation_test.py:141-215`` and the OEEM CAM ensemble
(``OEEM/classification/prepare_seg_inputs.py:96-138``) on the GPU: float64 canvases stay in HBM, tiles never go to
the host, the overlap-add is deterministic (one owner thread per canvas pixel, tiles visited in index order)."""
import torch

from . import ops


class BigMaskFuser:
    """One test image: accumulate softmax(tile logits) per scale, then normalise, resize to the image size, average
    over scales, argmax, confusion, background."""

    def __init__(self, image_hw, num_classes=3, device="cuda"):
        self.h, self.w = int(image_hw[0]), int(image_hw[1])
        self.C = num_classes
        self.device = torch.device(device)
        self.canvas, self.count = {}, {}

    def add_tiles(self, logits, scale, positions, crops):
        """logits CUDA f32 [n,C,Hp,Wp]; positions [(y, x)]; crops [(orig_h, orig_w)]  (segmentation_test.py:145-174);
        or positions = CUDA int32 [n,4] rows (y, x, crop_h, crop_w) with crops=None."""
        key = float(scale)
        if key not in self.canvas:
            hs, ws = int(self.h * scale), int(self.w * scale)
            self.canvas[key] = torch.zeros((self.C, hs, ws), dtype=torch.float64, device=self.device)
            self.count[key] = torch.zeros((hs, ws), dtype=torch.float64, device=self.device)
        if torch.is_tensor(positions) and positions.is_cuda:
            pos = positions                       # prebuilt int32 [n,4] (y, x, crop_h, crop_w) on the device: a fixed tile grid can be cached
        else:
            pos = [[p[0], p[1], c[0], c[1]] for p, c in zip(positions, crops)]
        ops.stitch_accumulate(logits, pos, self.canvas[key], self.count[key], softmax=True)

    def fused(self):
        """[C,h,w] float64: mean over scales of the normalised, resized canvases (segmentation_test.py:187-204)."""
        total = torch.empty((self.C, self.h, self.w), dtype=torch.float64, device=self.device)
        for i, (key, canvas) in enumerate(self.canvas.items()):
            # canvas / count, float64 bilinear to the image size and the sum over scales in one pass; the canvases stay as they are
            ops.canvas_resize_accumulate(canvas, self.count[key], total, accumulate=i > 0)
        if not self.canvas:
            total.zero_()
        ops.canvas_normalize(total, None, float(max(len(self.canvas), 1)))
        return total

    def finish(self, gt=None, conf=None, bg_match=3, bg_label=3):
        """argmax + confusion (before the background overwrite) + mask_pred[gt == 3] = 3 (segmentation_test.py:207-211)."""
        return ops.argmax_f64(self.fused(), gt=gt, bg_match=bg_match, bg_label=bg_label, conf=conf)


def cam_ensemble(cams_per_scale, positions_per_scale, scales, image_wh, side=224):
    """prepare_seg_inputs.py:96-136 for one image (w = rows, h = cols as in the reference).  cams: CUDA f32 [n,C,28,28]."""
    w, h = image_wh
    device = cams_per_scale[0].device
    C = cams_per_scale[0].shape[1]
    ens = torch.zeros((C, w, h), dtype=torch.float64, device=device)
    for s, scale in enumerate(scales):
        w_, h_ = int(w * scale), int(h * scale)
        ix, iy = (side if w_ >= side else w_), (side if h_ >= side else h_)
        crops = ops.upsample_bilinear(cams_per_scale[s], (ix, iy))      # f32, :116
        canvas = torch.zeros((C, w_, h_), dtype=torch.float64, device=device)
        count = torch.zeros((w_, h_), dtype=torch.float64, device=device)
        ops.stitch_accumulate(crops, [[y, x, ix, iy] for y, x in positions_per_scale[s]], canvas, count, softmax=False)
        ops.canvas_resize_accumulate(canvas, count, ens, min_count=1.0)  # sum_counter[sum_counter < 1] = 1; /; resize; ensemble +=
    ops.canvas_normalize(ens, None, float(len(scales)))
    return ens
