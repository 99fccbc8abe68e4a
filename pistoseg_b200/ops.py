"""Tensor-level wrappers over libpistoseg_b200 (C ABI in include/pistoseg_b200.h).

torch is used for device memory and streams only; every op below is one call into the shared library on
``torch.cuda.current_stream()``.
"""
import ctypes as C

import numpy as np
import torch

from . import _lib
from ._lib import (DECIDE_RAW, DECIDE_SOFTMAX, FUSE_LOGIT_MEAN, FUSE_PROB_MEAN, IMPL_AUTO, IMPL_GENERIC, IMPL_STREAM,
                   MASK_FILL, MASK_MULTIPLY, MASK_NEG_INF, MASK_NONE)


def _dev_index(t):
    if not t.is_cuda:
        raise _lib.PistoError("pistoseg_b200 ops need CUDA tensors (there is no CPU path); got a CPU tensor")
    return t.device.index if t.device.index is not None else torch.cuda.current_device()


def _stream(dev):
    return C.c_void_p(torch.cuda.current_stream(dev).cuda_stream)


def _ptr(t):
    return C.c_void_p(t.data_ptr()) if t is not None else C.c_void_p(None)


def _u8(t, name, device):
    if t is None:
        return None
    if isinstance(t, np.ndarray):
        t = torch.from_numpy(t)
    if t.dtype == torch.bool:
        t = t.to(torch.uint8)
    if t.dtype != torch.uint8:
        raise _lib.PistoError(f"{name} must be uint8/bool, got {t.dtype}")
    return t.to(device).contiguous()


def new_confusion(num_class, device):
    """Device-resident int64 [C,C] accumulator (rows = ground truth, cols = prediction)."""
    return torch.zeros((num_class, num_class), dtype=torch.int64, device=device)


def confusion_accumulate(pred, gt, conf):
    """conf[gt, pred] += 1 over pixels with gt < C   (replaces mIoUMask._generate_matrix, loss.py:17-24).
    Returns the number of counted pixels whose prediction was >= C (an error in the reference)."""
    dev = _dev_index(pred)
    pred = _u8(pred, "pred", pred.device)
    gt = _u8(gt, "gt", pred.device)
    if pred.numel() != gt.numel():
        raise _lib.PistoError("pred and gt differ in size")
    num_class = conf.shape[0]
    bad = torch.zeros(1, dtype=torch.int64, device=pred.device)
    lib = _lib.load()
    _lib.check(lib.pisto_confusion_accumulate(_lib.handle(dev), _ptr(pred), _ptr(gt), pred.numel(), num_class, _ptr(conf),
                                              _ptr(bad), _stream(dev)))
    return bad


def _make_views(views, xforms, C_):
    arr = (_lib.View * len(views))()
    keep = []
    for i, (v, code) in enumerate(zip(views, xforms)):
        if v.dtype != torch.float32:
            raise _lib.PistoError(f"view {i}: float32 expected, got {v.dtype}")
        if v.dim() != 4 or v.shape[1] < C_:
            raise _lib.PistoError(f"view {i}: expected [N,C,h,w], got {tuple(v.shape)}")
        stride = 0
        if not v.is_contiguous():
            # allow a channel slice of a contiguous tensor (e.g. [:, 1:] of [B, C+1, H, W], infer_revise_masks.py:139)
            if v.stride(3) == 1 and v.stride(2) == v.shape[3] and v.stride(1) == v.shape[2] * v.shape[3]:
                stride = v.stride(0)
            else:
                v = v.contiguous()
        keep.append(v)
        arr[i].logits = v.data_ptr()
        arr[i].tile_stride = stride
        arr[i].h, arr[i].w = int(v.shape[2]), int(v.shape[3])
        arr[i].xform = int(code)
    return arr, keep


def fuse_argmax_confusion(views, xforms=None, size=None, *, fuse_mode=FUSE_LOGIT_MEAN, mask_mode=MASK_NONE,
                          decide=DECIDE_SOFTMAX, present=None, bg=None, bg_match=0, bg_label=None, gt=None, conf=None,
                          want_labels=True, want_fused=False, want_entropy=False, lowres=None, impl=IMPL_AUTO, want_raw_labels=False):
    """The fused hot path (include/pistoseg_b200.h: pisto_fuse_argmax_confusion).

    views   list of CUDA float32 [N,C,h_v,w_v] logits as the backbone produced them for each augmented input
    xforms  de-augmentation code per view (k + 4*hflip), default all 0
    size    (T_h, T_w) output tile size, default = de-augmented size of view 0
    Returns a dict with the requested outputs: labels u8 [N,T_h,T_w], fused f32 [N,C,T_h,T_w],
    entropy f32 [N,T_h,T_w], lowres f32 [N,C,lh,lw], conf (the int64 [C,C] tensor passed in, updated in place).
    want_raw_labels: also ``labels_raw`` u8 [N,T_h,T_w] = the same scores decided with DECIDE_RAW, written in the same pass
    (one full-resolution view or impl=IMPL_GENERIC; segmentation_test.py:137-139,182).
    """
    v0 = views[0]
    dev = _dev_index(v0)
    N, C_ = int(v0.shape[0]), int(v0.shape[1])
    xforms = list(xforms) if xforms is not None else [0] * len(views)
    if size is None:
        size = (v0.shape[3], v0.shape[2]) if (xforms[0] & 1) else (v0.shape[2], v0.shape[3])
    T_h, T_w = int(size[0]), int(size[1])
    arr, keep = _make_views(views, xforms, C_)
    device = v0.device
    present_t = _u8(present, "present", device)
    bg_t = _u8(bg, "bg", device)
    gt_t = _u8(gt, "gt", device)
    if present_t is not None and tuple(present_t.shape) != (N, C_):
        raise _lib.PistoError(f"present must be [N,C]=({N},{C_}), got {tuple(present_t.shape)}")
    for name, t in (("bg", bg_t), ("gt", gt_t)):
        if t is not None and t.numel() != N * T_h * T_w:
            raise _lib.PistoError(f"{name} must have N*T_h*T_w elements")
    out = {}
    a = _lib.FuseArgs()
    a.N, a.C, a.T_h, a.T_w = N, C_, T_h, T_w
    a.fuse_mode, a.mask_mode, a.decide_mode = fuse_mode, (mask_mode if present_t is not None else MASK_NONE), decide
    a.bg_match = int(bg_match)
    a.bg_label = int(C_ if bg_label is None else bg_label)
    a.impl = impl
    a.present, a.bg, a.gt = _ptr(present_t).value, _ptr(bg_t).value, _ptr(gt_t).value
    if want_labels:
        out["labels"] = torch.empty((N, T_h, T_w), dtype=torch.uint8, device=device)
        a.label_out = out["labels"].data_ptr()
    if want_raw_labels:
        out["labels_raw"] = torch.empty((N, T_h, T_w), dtype=torch.uint8, device=device)
        a.label_raw_out = out["labels_raw"].data_ptr()
    if want_fused or (lowres is not None and not _is_gather((T_h, T_w), lowres)):
        out["fused"] = torch.empty((N, C_, T_h, T_w), dtype=torch.float32, device=device)
        a.fused_out = out["fused"].data_ptr()
    if want_entropy:
        out["entropy"] = torch.empty((N, T_h, T_w), dtype=torch.float32, device=device)
        a.entropy_out = out["entropy"].data_ptr()
    if lowres is not None:
        a.low_h, a.low_w = int(lowres[0]), int(lowres[1])
        out["lowres"] = torch.empty((N, C_, a.low_h, a.low_w), dtype=torch.float32, device=device)
        a.lowres_out = out["lowres"].data_ptr()
    if gt_t is not None:
        if conf is None:
            conf = new_confusion(C_, device)
        if conf.dtype != torch.int64 or tuple(conf.shape) != (C_, C_) or not conf.is_contiguous():
            raise _lib.PistoError("conf must be a contiguous int64 [C,C] CUDA tensor")
        a.conf = conf.data_ptr()
        out["conf"] = conf
    lib = _lib.load()
    _lib.check(lib.pisto_fuse_argmax_confusion(_lib.handle(dev), arr, len(views), C.byref(a), _stream(dev)))
    del keep
    return out


def _is_gather(size, low):
    return (size[0] % low[0] == 0 and size[1] % low[1] == 0 and (size[0] // low[0]) % 2 == 1 and (size[1] // low[1]) % 2 == 1)


def fuse_argmax_confusion_host(views, xforms=None, size=None, *, fuse_mode=FUSE_LOGIT_MEAN, mask_mode=MASK_NONE,
                               decide=DECIDE_SOFTMAX, present=None, bg=None, bg_match=0, bg_label=None, gt=None,
                               conf=None, want_labels=True, lowres=None, chunk=1024, device=0, out=None):
    """Same op with HOST buffers (CPU torch tensors, ideally pinned): H2D, kernel and D2H are pipelined inside the
    library (pisto_fuse_argmax_confusion_host).  This is the call bench.py's "e2e" number times.  The views may also be CUDA
    tensors (the reference's own dataflow, infer_pseudo_masks.py:119-137: the backbone output stays on the GPU, only ``tissue``
    comes from the host and labels / 32x32 logits go back): then only the byte masks are uploaded.
    ``out`` may carry preallocated pinned 'labels' / 'lowres' tensors.  conf is an int64 [C,C] CPU tensor (accumulated)."""
    v0 = views[0]
    N, C_ = int(v0.shape[0]), int(v0.shape[1])
    xforms = list(xforms) if xforms is not None else [0] * len(views)
    if size is None:
        size = (v0.shape[2], v0.shape[3])
    T_h, T_w = int(size[0]), int(size[1])
    arr = (_lib.View * len(views))()
    for i, (v, code) in enumerate(zip(views, xforms)):
        if v.dtype != torch.float32 or not v.is_contiguous():
            raise _lib.PistoError("views must be contiguous float32 tensors (CPU, ideally pinned; or already on the GPU)")
        if v.is_cuda and _dev_index(v) != int(device):
            raise _lib.PistoError("device-resident views must live on the device the call runs on")
        arr[i].logits = v.data_ptr(); arr[i].tile_stride = 0
        arr[i].h, arr[i].w, arr[i].xform = int(v.shape[2]), int(v.shape[3]), int(code)
    out = dict(out or {})
    a = _lib.FuseArgs()
    a.N, a.C, a.T_h, a.T_w = N, C_, T_h, T_w
    a.fuse_mode, a.mask_mode, a.decide_mode = fuse_mode, (mask_mode if present is not None else MASK_NONE), decide
    a.bg_match = int(bg_match)
    a.bg_label = int(C_ if bg_label is None else bg_label)
    keep = []
    for name, t in (("present", present), ("bg", bg), ("gt", gt)):
        if t is not None:
            if t.is_cuda or t.dtype != torch.uint8 or not t.is_contiguous():
                raise _lib.PistoError(f"host {name} must be a contiguous uint8 CPU tensor")
            setattr(a, name, t.data_ptr()); keep.append(t)
    if want_labels:
        if "labels" not in out:
            out["labels"] = torch.empty((N, T_h, T_w), dtype=torch.uint8).pin_memory()
        a.label_out = out["labels"].data_ptr()
    if lowres is not None:
        if not _is_gather((T_h, T_w), lowres):
            raise _lib.PistoError("host pipeline exports lowres only for odd-factor gathers (e.g. 224 -> 32)")
        a.low_h, a.low_w = int(lowres[0]), int(lowres[1])
        if "lowres" not in out:
            out["lowres"] = torch.empty((N, C_, a.low_h, a.low_w), dtype=torch.float32).pin_memory()
        a.lowres_out = out["lowres"].data_ptr()
    if gt is not None:
        if conf is None:
            conf = torch.zeros((C_, C_), dtype=torch.int64)
        a.conf = conf.data_ptr()
        out["conf"] = conf
    lib = _lib.load()
    _lib.check(lib.pisto_fuse_argmax_confusion_host(_lib.handle(device), arr, len(views), C.byref(a), int(chunk)))
    return out


def upsample_bilinear(x, size):
    """F.interpolate(x, size, mode='bilinear', align_corners=False) for CUDA float32 / float64 [..., h, w]
    (replaces interpolate_tensor: infer_pseudo_masks.py:89-90, segmentation_test.py:88-89,197)."""
    dev = _dev_index(x)
    if x.dtype not in (torch.float32, torch.float64):
        raise _lib.PistoError(f"float32/float64 expected, got {x.dtype}")
    x = x.contiguous()
    lead = x.shape[:-2]
    out = torch.empty((*lead, int(size[0]), int(size[1])), dtype=x.dtype, device=x.device)
    nc = int(np.prod(lead)) if len(lead) else 1
    lib = _lib.load()
    _lib.check(lib.pisto_upsample_bilinear(_lib.handle(dev), _ptr(x), _ptr(out), nc, x.shape[-2], x.shape[-1], int(size[0]),
                                           int(size[1]), 0 if x.dtype == torch.float32 else 1, _stream(dev)))
    return out


def stitch_accumulate(tiles, pos, canvas, count, softmax=True):
    """canvas[:, y:y+h, x:x+w] += softmax(tile)[:, :h, :w]; count += 1 for every tile, in tile order
    (segmentation_test.py:145-174).  tiles f32 [n,C,th,tw] CUDA; pos int32 [n,4] = (y, x, crop_h, crop_w);
    canvas f64 [C,H,W]; count f64 [H,W]."""
    dev = _dev_index(tiles)
    tiles = tiles.contiguous()
    pos = torch.as_tensor(pos, dtype=torch.int32).reshape(-1, 4).to(tiles.device).contiguous()
    n, C_, th, tw = tiles.shape
    if canvas.dtype != torch.float64 or count.dtype != torch.float64 or not canvas.is_contiguous() or not count.is_contiguous():
        raise _lib.PistoError("canvas / count must be contiguous float64 CUDA tensors")
    if pos.shape[0] != n:
        raise _lib.PistoError(f"stitch_accumulate: {pos.shape[0]} positions for {n} tiles")
    if n:
        # a crop larger than the tile would read the next tile (or past the buffer); the reference's numpy slicing cannot do that
        lo = pos.min(dim=0).values.tolist()
        hi = pos.max(dim=0).values.tolist()
        if lo[0] < 0 or lo[1] < 0 or lo[2] < 1 or lo[3] < 1 or hi[2] > th or hi[3] > tw:
            raise _lib.PistoError(f"stitch_accumulate: positions need y, x >= 0 and 1 <= crop_h <= {th}, 1 <= crop_w <= {tw}; got min {lo}, max {hi}")
    lib = _lib.load()
    _lib.check(lib.pisto_stitch_accumulate(_lib.handle(dev), _ptr(tiles), _ptr(pos), n, C_, th, tw, int(bool(softmax)), _ptr(canvas),
                                           _ptr(count), canvas.shape[-2], canvas.shape[-1], _stream(dev)))


def canvas_normalize(canvas, count=None, min_count=0.0):
    """canvas /= count (clamped below by min_count when > 0); count=None divides by the constant min_count."""
    dev = _dev_index(canvas)
    C_ = canvas.shape[0]
    hw = canvas[0].numel()
    lib = _lib.load()
    _lib.check(lib.pisto_canvas_normalize(_lib.handle(dev), _ptr(canvas), _ptr(count), C_, hw, float(min_count), _stream(dev)))


def canvas_resize_accumulate(canvas, count, out, min_count=0.0, accumulate=True):
    """out (+)= bilinear_f64(canvas / count) resized to out's size, in one pass (segmentation_test.py:187-199,
    prepare_seg_inputs.py:128-134).  canvas f64 [C,hi,wi], count f64 [hi,wi], out f64 [C,ho,wo]; canvas is not modified."""
    dev = _dev_index(canvas)
    canvas, count = canvas.contiguous(), count.contiguous()
    if canvas.dtype != torch.float64 or count.dtype != torch.float64 or out.dtype != torch.float64 or not out.is_contiguous():
        raise _lib.PistoError("canvas_resize_accumulate: contiguous float64 tensors expected")
    lib = _lib.load()
    _lib.check(lib.pisto_canvas_resize_accumulate(_lib.handle(dev), _ptr(canvas), _ptr(count), canvas.shape[0], canvas.shape[1], canvas.shape[2],
                                                  float(min_count), _ptr(out), out.shape[1], out.shape[2], int(bool(accumulate)), _stream(dev)))


def canvas_axpy(out, x, scale=1.0):
    """out += x * scale (float64)."""
    dev = _dev_index(out)
    lib = _lib.load()
    _lib.check(lib.pisto_canvas_axpy(_lib.handle(dev), _ptr(out), _ptr(x.contiguous()), out.numel(), float(scale), _stream(dev)))


def argmax_f64(scores, present=None, gt=None, bg_match=3, bg_label=3, conf=None, want_pred=True, want_labels=True):
    """np.argmax over the class axis of planar float64 scores [C,H,W] (+ confusion before the background overwrite,
    + label[gt == bg_match] = bg_label): segmentation_test.py:207-211, generate_CAM.py:91-99."""
    dev = _dev_index(scores)
    scores = scores.contiguous()
    C_ = scores.shape[0]
    hw = scores[0].numel()
    device = scores.device
    gt_t = _u8(gt, "gt", device)
    pres = None
    if present is not None:
        pres = (C.c_uint8 * C_)(*[int(v) for v in present])
    out = {}
    if want_pred:
        out["pred"] = torch.empty(scores.shape[1:], dtype=torch.uint8, device=device)
    if want_labels:
        out["labels"] = torch.empty(scores.shape[1:], dtype=torch.uint8, device=device)
    if gt_t is not None and conf is not None:
        out["conf"] = conf
    lib = _lib.load()
    _lib.check(lib.pisto_argmax_f64(_lib.handle(dev), _ptr(scores), C_, hw, C.cast(pres, C.c_void_p) if pres is not None else None,
                                    _ptr(gt_t), int(bg_match), int(bg_label), _ptr(out.get("pred")), _ptr(out.get("labels")),
                                    _ptr(conf) if gt_t is not None else None, _stream(dev)))
    return out


def mosaic_gather(pool, plans, cells, patch_num, patch_size, bg_label=3, packed=True):
    """pool: dict(img u8 flat HWC, bg u8 flat or None, off int64 [P], hw int32 [P,2], label u8 [P]) of CUDA tensors;
    plans: CUDA uint8 view of MosaicPlan[N]; cells: CUDA uint8 view of MosaicCell[N,4,pn*pn].
    packed=True gathers from the 32-bit packed copy of the pool (built on first use), packed=False from the planar buffers.
    Returns (img u8 [N,S,S,3], mask u8 [N,S,S])  (create_dataset.ipynb:273-372)."""
    dev = _dev_index(pool["img"])
    device = pool["img"].device
    N = plans.numel() // C.sizeof(_lib.MosaicPlan)
    S = patch_num * patch_size
    img = torch.empty((N, S, S, 3), dtype=torch.uint8, device=device)
    mask = torch.empty((N, S, S), dtype=torch.uint8, device=device)
    lib = _lib.load()
    if packed:
        if pool.get("rgba") is None:  # pack once per pool: r | g << 8 | b << 16 | (bg > 0) << 24
            n_px = pool["img"].numel() // 3
            pool["rgba"] = torch.empty(n_px, dtype=torch.int32, device=device)
            _lib.check(lib.pisto_mosaic_pack_pool(_lib.handle(dev), _ptr(pool["img"]), _ptr(pool.get("bg")), n_px, _ptr(pool["rgba"]), _stream(dev)))
        _lib.check(lib.pisto_mosaic_gather_packed(_lib.handle(dev), _ptr(pool["rgba"]), _ptr(pool["off"]), _ptr(pool["hw"]), _ptr(pool["label"]),
                                                  _ptr(plans), _ptr(cells), N, patch_num, patch_size, int(bg_label), _ptr(img), _ptr(mask), _stream(dev)))
    else:
        _lib.check(lib.pisto_mosaic_gather(_lib.handle(dev), _ptr(pool["img"]), _ptr(pool.get("bg")), _ptr(pool["off"]), _ptr(pool["hw"]),
                                           _ptr(pool["label"]), _ptr(plans), _ptr(cells), N, patch_num, patch_size, int(bg_label),
                                           _ptr(img), _ptr(mask), _stream(dev)))
    return img, mask


def get_background(rgb, thresh=200, min_size=50):
    """rgb: CUDA uint8 [N,H,W,3] (or [H,W,3]) -> uint8 [N,H,W] (or [H,W]) in {0, 255}: gray > thresh with 4-connected
    components smaller than min_size removed (utils.py:155-163, dataset.py:100-109)."""
    if not rgb.is_cuda or rgb.dtype != torch.uint8 or rgb.shape[-1] != 3 or rgb.dim() not in (3, 4):
        raise _lib.PistoError("get_background: expected a CUDA uint8 tensor [N,H,W,3] or [H,W,3]")
    single = rgb.dim() == 3
    x = (rgb[None] if single else rgb).contiguous()
    N, H, W = x.shape[:3]
    dev = _dev_index(x)
    out = torch.empty((N, H, W), dtype=torch.uint8, device=x.device)
    scratch = torch.empty(2 * N * H * W, dtype=torch.int32, device=x.device)
    _lib.check(_lib.load().pisto_get_background(_lib.handle(dev), _ptr(x), N, H, W, int(thresh), int(min_size), _ptr(scratch), _ptr(out), _stream(dev)))
    return out[0] if single else out


def resize_nearest_bg(labels, items, index_pool, bg_pool, out_pool, bg_value=3):
    """``pisto_resize_nearest_bg``: labels CUDA u8 [n,S_h,S_w]; items = list of (tile, h, w, iy_off, ix_off, bg_off, out_off);
    index_pool CUDA int32, bg_pool CUDA u8 or None, out_pool CUDA u8 (written).  See postproc.revise_masks_to_original."""
    dev = _dev_index(labels)
    arr = (_lib.ResizeDesc * len(items))()
    for d, it in zip(arr, items):
        d.tile, d.h, d.w, d.iy_off, d.ix_off, d.bg_off, d.out_off = (int(v) for v in it)
    host = torch.frombuffer(bytearray(bytes(arr)), dtype=torch.uint8) if len(items) else torch.empty(0, dtype=torch.uint8)
    ddesc = host.to(labels.device)
    lib = _lib.load()
    _lib.check(lib.pisto_resize_nearest_bg(_lib.handle(dev), labels.data_ptr(), int(labels.shape[0]), int(labels.shape[1]), int(labels.shape[2]),
                                           ddesc.data_ptr(), len(items), index_pool.data_ptr(), bg_pool.data_ptr() if bg_pool is not None else None,
                                           out_pool.data_ptr(), int(bg_value), _stream(dev)))
    return out_pool

