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


def pil_nearest_index(n_in, n_out):
    """Source index of every output row / column of PIL's NEAREST resize (``Image.resize`` on a mode-'P' image; Pillow
    ``ImagingScaleAffine``): the double ``xo = 0.5 * a`` is ACCUMULATED (``xo += a``, ``a = n_in / n_out``), index = ``int(xo)``.
    ``np.add.accumulate`` performs exactly those sequential double additions."""
    a = n_in / n_out
    steps = np.full(n_out, a, dtype=np.float64)
    steps[0] = a * 0.5
    return np.minimum(np.add.accumulate(steps).astype(np.int64), n_in - 1).astype(np.int32)


def revise_masks_to_original(x, label, original_hw, backgrounds=None, bg_value=3):
    """infer_revise_masks.py:137-143,152-155 for one head: ``(x * label)[:, 1:]`` -> argmax -> PIL mode-'P' resize (NEAREST) to each
    tile's ORIGINAL ``(h, w)`` -> ``mask[background > 0] = 3`` at the original resolution (the reference applies the background
    AFTER the resize, on the full-size ``utils.get_background`` mask).

    x CUDA f32 [B,C+1,S,S]; label [B,C+1]; original_hw: list of (h, w); backgrounds: list of [h,w] uint8 arrays / tensors
    (0 / 255) or None.  Returns a list of CUDA uint8 [h,w] tensors (views of one pooled buffer)."""
    B = x.shape[0]
    small = revise_masks(x, label)                      # [B,S,S] u8, background not applied yet
    S_h, S_w = int(small.shape[1]), int(small.shape[2])
    idx, items, off_i, off_o = [], [], 0, 0
    for j, (h, w) in enumerate(original_hw):
        iy, ix = pil_nearest_index(S_h, int(h)), pil_nearest_index(S_w, int(w))
        idx += [iy, ix]
        items.append([j, int(h), int(w), off_i, off_i + int(h), (off_o if backgrounds is not None else -1), off_o])
        off_i += int(h) + int(w)
        off_o += int(h) * int(w)
    index_pool = torch.from_numpy(np.concatenate(idx)).to(x.device)
    bg_pool = None
    if backgrounds is not None:
        bg_pool = torch.cat([torch.as_tensor(np.asarray(b) if not torch.is_tensor(b) else b).reshape(-1).to(torch.uint8).to(x.device, non_blocking=True)
                             for b in backgrounds])
    out_pool = torch.empty(off_o, dtype=torch.uint8, device=x.device)
    ops.resize_nearest_bg(small, items, index_pool, bg_pool, out_pool, bg_value)
    return [out_pool[it[6]:it[6] + it[1] * it[2]].view(it[1], it[2]) for it in items]


WSSS4LUAD_PALETTE = [0, 64, 128, 64, 128, 0, 243, 152, 0, 255, 255, 255] + [0] * 252 * 3   # infer_revise_masks.py:150
BCSS_PALETTE = [255, 0, 0, 0, 255, 0, 0, 0, 255, 153, 0, 255, 255, 255, 255]                # infer_revise_masks.py:182-187


def revise_masks_to_png(heads, label, names, original_hw, save_dir, backgrounds=None, dataset="wsss4luad", pool=None):
    """The saving half of infer_revise_masks.py:145-206: for every head (``{'pmask': x, 'pcam': x, 'cam': x}``) writes
    ``<save_dir>/refine/<head>/<name>.png`` -- mode 'P', the dataset's palette, original size; wsss4luad also overwrites the
    background.  PNG encoding stays on the host (threaded through ``io.AsyncWriter`` when ``pool`` is given)."""
    import os
    from PIL import Image
    wsss = dataset == "wsss4luad"
    palette = WSSS4LUAD_PALETTE if wsss else BCSS_PALETTE
    written = []
    for head, x in heads.items():
        d = os.path.join(save_dir, "refine", head)
        os.makedirs(d, exist_ok=True)
        masks = revise_masks_to_original(x, label, original_hw, backgrounds if wsss else None)
        for name, m in zip(names, masks):
            arr = m.cpu().numpy()
            path = os.path.join(d, name + ".png")

            def save(arr=arr, path=path):
                im = Image.fromarray(np.uint8(arr), mode="P")
                im.putpalette(palette)
                im.save(path)
            if pool is not None:
                pool.submit(save)
            else:
                save()
            written.append(path)
    return written

