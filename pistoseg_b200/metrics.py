"""``mIoUMask`` -- drop-in for the reference's metric module (``loss.py:8-67``), computed on the GPU.

Same constructor, attributes and methods.  The confusion matrix lives on the device as int64 (exact counts) and is
materialised as the reference's float64 numpy ``[C, C]`` array whenever ``confusion_matrix`` is read; the IoU formulas
(``loss.py:33-53``) are the reference's float64 numpy expressions, unchanged.
"""
import numpy as np
import torch

from . import _lib, ops


class mIoUMask(torch.nn.Module):

    def __init__(self, num_classes=3, ignore_class=None, eps=1e-7, device=None):
        super().__init__()
        self.eps = eps
        self.num_class = num_classes + (1 if ignore_class is not None else 0)
        self.ignore_class = ignore_class
        self._device = torch.device(device) if device is not None else None
        self._conf = None            # device int64 [C, C]
        self._host = np.zeros((self.num_class,) * 2)   # counts added through the numpy path / assignment

    # ---- device accumulator -------------------------------------------------------------------------------------
    def _acc(self, device):
        if self._conf is None:
            self._device = torch.device(device)
            self._conf = ops.new_confusion(self.num_class, self._device)
        return self._conf

    @property
    def confusion_matrix(self):
        m = self._host.copy()
        if self._conf is not None:
            m = m + self._conf.cpu().numpy().astype(np.float64)
        if self.ignore_class is not None:
            m[self.ignore_class, :] = 0   # loss.py:19-20: pixels whose gt is the ignore class are not counted
        return m

    @confusion_matrix.setter
    def confusion_matrix(self, value):
        self._host = np.array(value, dtype=np.float64)
        if self._conf is not None:
            self._conf.zero_()

    def all_reduce(self, group=None):
        """Sum the device accumulator over the ranks of a torch.distributed process group (one NCCL all-reduce of
        C*C int64 values); exact, independent of the number of ranks."""
        import torch.distributed as dist
        if dist.is_available() and dist.is_initialized():
            # every rank takes part, also one whose shard was empty and never created its accumulator (else the others would
            # block in the collective for ever)
            conf = self._conf if self._conf is not None else self._acc(self._device or torch.device("cuda", torch.cuda.current_device()))
            dist.all_reduce(conf, op=dist.ReduceOp.SUM, group=group)

    # ---- reference API ----------------------------------------------------------------------------------------------
    def _generate_matrix(self, pre_image, gt_image):
        """numpy in, numpy out (``loss.py:17-24``), counted by pisto_confusion_accumulate."""
        dev = self._device or torch.device("cuda", torch.cuda.current_device())
        pred = torch.as_tensor(np.ascontiguousarray(pre_image)).to(torch.uint8).to(dev)
        gt_np = np.asarray(gt_image)
        gt = torch.as_tensor(np.ascontiguousarray(np.where((gt_np >= 0) & (gt_np < 255), gt_np, 255))).to(torch.uint8).to(dev)
        conf = ops.new_confusion(self.num_class, dev)
        ops.confusion_accumulate(pred, gt, conf)
        m = conf.cpu().numpy()
        if self.ignore_class is not None:
            m[self.ignore_class, :] = 0
        return m

    def reset(self):
        self._host = np.zeros((self.num_class,) * 2)
        if self._conf is not None:
            self._conf.zero_()

    def add_batch(self, gt_image, pre_image):
        # the reference's argument names are swapped twice and cancel (loss.py:29-31, 65): first = prediction
        assert gt_image.shape == pre_image.shape
        self._host += self._generate_matrix(gt_image, pre_image)

    def Tissue_Intersection_over_Union(self):
        cm = self.confusion_matrix
        with np.errstate(invalid="ignore", divide="ignore"):
            MIoU = np.diag(cm) / (np.sum(cm, axis=1) + np.sum(cm, axis=0) - np.diag(cm))
        MIoU[np.isnan(MIoU)] = 0
        return MIoU

    def Mean_Intersection_over_Union(self):
        return np.mean(self.Tissue_Intersection_over_Union())

    def Frequency_Weighted_Intersection_over_Union(self):
        cm = self.confusion_matrix
        with np.errstate(invalid="ignore", divide="ignore"):
            freq = np.sum(cm, axis=1) / np.sum(cm)
            iu = np.diag(cm) / (np.sum(cm, axis=1) + np.sum(cm, axis=0) - np.diag(cm))
        return (freq[freq > 0] * iu[freq > 0]).sum()

    def update(self, logits, mask, probs=False):
        """forward() without the host read-back: softmax / argmax / confusion in one kernel, nothing leaves the GPU."""
        if not logits.is_cuda:
            raise _lib.PistoError("mIoUMask needs CUDA logits (pistoseg_b200 has no CPU path)")
        gt = mask.to(logits.device).byte()          # loss.py:62
        conf = self._acc(logits.device)
        if logits.dtype == torch.float64:
            if not probs:
                raise _lib.PistoError("float64 input is only supported with probs=True (segmentation_test.py:207)")
            for n in range(logits.shape[0]):
                ops.argmax_f64(logits[n], gt=gt[n], bg_match=255, bg_label=0, conf=conf, want_pred=False, want_labels=False)
            return
        ops.fuse_argmax_confusion([logits.float()], [0], logits.shape[-2:], decide=_lib.DECIDE_RAW if probs else _lib.DECIDE_SOFTMAX,
                                  gt=gt, conf=conf, want_labels=False)

    def forward(self, logits, mask, probs=False):
        self.update(logits, mask, probs)
        return self.Mean_Intersection_over_Union(), self.Frequency_Weighted_Intersection_over_Union()
