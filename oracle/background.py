"""Background mask -- restates ``utils.get_background`` (``utils.py:155-163``) / ``BaseDataset._get_background``
(``dataset.py:100-109``).  TEST INFRASTRUCTURE ONLY.

``cv2.cvtColor`` / ``cv2.threshold`` are the reference's own calls (cv2 is installed).  ``skimage`` is not installed, so
``morphology.remove_small_objects(ar, min_size, connectivity=1)`` (scikit-image 0.19.3, ``environment.yaml``) is restated from
its published algorithm: label the boolean array with ``scipy.ndimage.label`` and the connectivity-1 (4-neighbour)
footprint, ``np.bincount`` the component sizes, clear the components with ``size < min_size`` -- parity unpinned for that
one call (no skimage here to run), pinned for the rest.
"""
import cv2
import numpy as np
from scipy import ndimage as ndi


def remove_small_objects(ar, min_size=64, connectivity=1):
    ar = np.asarray(ar, bool)
    out = ar.copy()
    if min_size == 0:
        return out
    footprint = ndi.generate_binary_structure(ar.ndim, connectivity)
    ccs, _ = ndi.label(ar, footprint)
    sizes = np.bincount(ccs.ravel())
    too_small = sizes < min_size
    out[too_small[ccs]] = 0
    return out


def get_background(region):
    gray = cv2.cvtColor(region, cv2.COLOR_RGB2GRAY)
    _, binary = cv2.threshold(gray, 200, 255, cv2.THRESH_BINARY)
    binary = np.uint8(binary)
    dst = remove_small_objects(binary == 255, min_size=50, connectivity=1)
    return np.asarray(dst, dtype=np.uint8) * 255
