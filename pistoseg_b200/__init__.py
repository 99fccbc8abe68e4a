"""pistoseg_b200 -- B200 (sm_100a) implementation of PistoSeg's dense-prediction post-processing hot path.

Host-side mirror of the reference's Python surface for that path:
  ops       tensor-level wrappers over the C ABI (include/pistoseg_b200.h)
  postproc  interpolate_tensor / get_mask_pred_and_entropy / pseudo-mask batch driver (infer_pseudo_masks.py)
  metrics   mIoUMask (loss.py)
  tta       SegmentationTTAWrapper + d4_transform stand-ins (ttach)
  stitch    big-mask canvas fusion (segmentation_test.py:141-215), OEEM CAM ensemble
  mosaic    mosaic plan generator + synthesis (create_dataset.ipynb CropAndConcatDataset)
  dist      tile sharding + confusion-matrix all-reduce
  validation  Lightning validation hooks (models/mosaic_module.py, models/segmentation_module.py) as a mixin
  background  get_background (utils.py:155-163, dataset.py:100-109)
  oeem      OEEM tiling positions, CAM ensemble exports (OEEM/classification/prepare_seg_inputs.py, generate_CAM.py)
  io        palette PNG / .pt writers on a thread pool
The compute lives in libpistoseg_b200.so; there is no CPU or torch fallback.
"""
__version__ = "0.1.0"
