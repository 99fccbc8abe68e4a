"""Drop-in for the metric half of the reference's ``loss.py``: ``from loss import mIoUMask`` keeps working, the
confusion matrix / softmax / argmax now run in libpistoseg_b200 (see pistoseg_b200/metrics.py).

``DiceLoss`` (``loss.py:70-118`` of the reference) is a training loss and is outside this repository's scope
(stage-1 / stage-5 training stays in the reference)."""
from pistoseg_b200.metrics import mIoUMask  # noqa: F401

__all__ = ["mIoUMask"]
