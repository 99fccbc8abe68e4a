"""CPU oracle for the PistoSeg post-processing hot path.

TEST INFRASTRUCTURE ONLY.  This package restates, on the CPU (torch-CPU / numpy), the arithmetic that
the reference executes after the backbone logits.  Only ``tests/``, ``__graft_entry__.smoke()`` and the
``cpu_baseline`` / ``--impl reference`` legs of ``bench.py`` may import it, and only as the checker or the
timed CPU baseline -- never from ``pistoseg_b200`` (the product has exactly one backend: the sm_100a
library, and fails loudly if that library is missing).

Pinning status (see DESIGN.md "Oracle"): the reference ships no tests / golden vectors for this path
(SURVEY.md section 4), so the oracle is pinned against *outputs of the reference's own code and of the
libraries it calls, run in the build container*:
  * ``loss.py`` of the reference is imported directly by ``tests/golden/make_golden.py`` (it needs only
    torch + numpy) and its confusion matrices / IoU reports are committed as fixtures;
  * ``torch.nn.functional.interpolate`` / ``softmax`` / ``argmax`` (what ``interpolate_tensor`` and
    ``get_mask_pred_and_entropy`` call) pin the restated bilinear / fusion arithmetic bit-for-bit;
  * ``cv2.warpAffine`` / ``cv2.flip`` / ``cv2.getRotationMatrix2D`` pin the mosaic warp bit-for-bit.
  * ttach 0.0.3 and albumentations 1.2.1 are NOT installed: the d4 merge and the mosaic parameter
    sampling are restated from their published behaviour; those two pieces are "parity unpinned".
"""
