"""``get_background`` -- drop-in for ``utils.get_background`` (``utils.py:155-163``) and ``BaseDataset._get_background``
(``dataset.py:100-109``): numpy RGB region in, numpy uint8 mask ({0, 255}) out, computed on the GPU
(``pisto_get_background``: fixed-point gray, threshold, 4-connected component sizes by union-find)."""
import numpy as np
import torch

from . import ops


def get_background(region, device="cuda"):
    """region: uint8 [H,W,3] (RGB) numpy array or tensor.  Same result as the reference, bit for bit."""
    t = torch.as_tensor(np.ascontiguousarray(region) if isinstance(region, np.ndarray) else region)
    on_host = not t.is_cuda
    mask = ops.get_background(t.to(device) if on_host else t)
    return mask.cpu().numpy() if on_host else mask


def get_background_batch(regions):
    """regions: CUDA uint8 [N,H,W,3] -> CUDA uint8 [N,H,W]; nothing leaves the device."""
    return ops.get_background(regions)
