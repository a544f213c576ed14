"""paste_masks_in_image -- same surface as detectron2/layers/mask_ops.py:74-147, one fused kernel."""
from typing import Tuple

import torch

from .. import ops  # noqa: F401  (registers torch.ops.d2b200.*)

__all__ = ["paste_masks_in_image"]


@torch.jit.script_if_tracing
def paste_masks_in_image(masks: torch.Tensor, boxes: torch.Tensor, image_shape: Tuple[int, int], threshold: float = 0.5):
    """masks (N, M, M) soft masks in [0,1]; boxes Boxes or (N, 4) tensor; returns (N, H, W) bool masks
    (uint8 = value*255 when threshold < 0).  No chunking / 1 GB budget needed (mask_ops.py:14,123): nothing but the
    output is materialised.  Scriptable like the reference (tests/layers/test_mask_ops.py:156-165)."""
    assert masks.shape[-1] == masks.shape[-2], "Only square mask predictions are supported"
    n = len(masks)
    if n == 0:
        return masks.new_empty((0,) + image_shape, dtype=torch.uint8)
    if not isinstance(boxes, torch.Tensor):
        boxes = boxes.tensor
    assert len(boxes) == n, boxes.shape
    img_h, img_w = int(image_shape[0]), int(image_shape[1])
    if masks.dim() == 4:  # (N, 1, M, M) as produced upstream
        masks = masks[:, 0]
    return torch.ops.d2b200.paste_masks(masks, boxes, img_h, img_w, float(threshold))
