"""Function-level drop-ins for the post-processing helpers of ``infer_pseudo_masks.py`` / ``segmentation_test.py`` /
``infer_revise_masks.py`` (SURVEY.md 8(b) level 2), each one call into libpistoseg_b200."""
import numpy as np
import torch

from . import _lib, ops


def interpolate_tensor(tensor, target_shape):
    """``F.interpolate(tensor.unsqueeze(0), target_shape, mode='bilinear')[0]`` (infer_pseudo_masks.py:89-90)."""
    return ops.upsample_bilinear(tensor.unsqueeze(0), target_shape)[0]


def check_tissue_region_is_too_small(patch_mask_pred, patch_label):
    """infer_pseudo_masks.py:62-67 (unused by the pipeline; host-side numpy as in the reference)."""
    for i in range(len(patch_label)):
        if patch_label[i] == 1:
            if np.sum(patch_mask_pred == i) / (patch_mask_pred.shape[-2] * patch_mask_pred.shape[-1]) < 0.1:
                return True
    return False


def get_mask_pred_and_entropy(patch_logit_pred, tissue, patch_label):
    """infer_pseudo_masks.py:69-87 for ONE tile: returns (np.int64 [H,W] mask, entropy [H,W]).

    Keeps the reference's side effect: for multi-label tiles the caller's logits are overwritten with -1e10 on the
    absent classes."""
    C_, H, W = patch_logit_pred.shape
    if sum(patch_label) == 1:
        mask_pred = np.full((H, W), patch_label.index(1))
        entropy = np.zeros_like(mask_pred)
        mask_pred[np.asarray(tissue) == 0] = len(patch_label)
        return mask_pred, entropy
    present = torch.tensor([patch_label], dtype=torch.uint8)
    tis = torch.as_tensor(np.asarray(tissue) == 0).to(torch.uint8)[None]
    out = ops.fuse_argmax_confusion([patch_logit_pred[None].contiguous()], [0], (H, W), mask_mode=_lib.MASK_FILL,
                                    decide=_lib.DECIDE_SOFTMAX, present=present, bg=tis, bg_match=1, bg_label=len(patch_label),
                                    want_entropy=True)
    for i, v in enumerate(patch_label):   # Warning: inplace operation! (infer_pseudo_masks.py:78)
        if v == 0:
            patch_logit_pred[i, :, :] = -1e10
    return out["labels"][0].cpu().numpy().astype(np.int64), out["entropy"][0].cpu().numpy()


def pseudo_mask_batch(views, xforms, size, present, tissue_is_bg, lowres=(32, 32), want_fused=False):
    """The whole per-batch post-processing of infer_pseudo_masks.py:121-137 in ONE kernel launch:
    TTA merge of the raw per-view model outputs, 32x32 logit export, label masking, softmax/argmax, background.

    views: list of CUDA f32 [B,C,h,w] model outputs (one per augmented input); xforms: de-augmentation codes;
    present: [B,C] 0/1; tissue_is_bg: [B,H,W] u8/bool (1 where tissue == 0) or None.
    Returns dict(labels u8 [B,H,W], lowres f32 [B,C,32,32][, fused])."""
    C_ = views[0].shape[1]
    return ops.fuse_argmax_confusion(views, xforms, size, mask_mode=_lib.MASK_FILL, decide=_lib.DECIDE_SOFTMAX, present=present,
                                     bg=tissue_is_bg, bg_match=1, bg_label=C_, lowres=lowres, want_fused=want_fused)


def revise_masks(x, label, background=None, bg_value=3):
    """infer_revise_masks.py:137-143,154-155: (x * label[B,C+1,1,1])[:, 1:] -> argmax -> mask[background > 0] = 3.
    x CUDA f32 [B,C+1,H,W]; label [B,C+1] (first entry = the constant bg score 1); background [B,H,W] u8 (0/255)."""
    B, C1, H, W = x.shape
    present = (label[:, 1:] != 0).to(torch.uint8)
    bg = None if background is None else (torch.as_tensor(background) > 0).to(torch.uint8)
    out = ops.fuse_argmax_confusion([x[:, 1:]], [0], (H, W), mask_mode=_lib.MASK_MULTIPLY, decide=_lib.DECIDE_RAW, present=present,
                                    bg=bg, bg_match=1, bg_label=bg_value)
    return out["labels"]
