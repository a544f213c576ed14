"""paste_masks_in_image -- same surface as detectron2/layers/mask_ops.py:74-147, one fused kernel."""
from typing import Tuple

import torch

from .. import ops  # noqa: F401  (registers torch.ops.d2b200.*)

__all__ = ["paste_masks_in_image", "paste_masks_in_image_packed", "unpack_mask_bits"]


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


def paste_masks_in_image_packed(masks: torch.Tensor, boxes: torch.Tensor, image_shape: Tuple[int, int], threshold: float = 0.5):
    """`paste_masks_in_image` with the boolean result bit-packed on the device: int32 (N, H, ceil(W / 32)), bit b of word w of
    row y = pixel (y, 32 w + b).  Same decisions as the byte form, 1 / 8 of the bytes to copy to the host (the consumer of the
    pasted masks -- RLE encoding for COCO evaluation, postprocessing.py:61-66 -- runs there); `unpack_mask_bits` restores the
    (N, H, W) bool tensor on either side."""
    assert masks.shape[-1] == masks.shape[-2], "Only square mask predictions are supported"
    n = len(masks)
    img_h, img_w = int(image_shape[0]), int(image_shape[1])
    if n == 0:
        return masks.new_empty((0, img_h, (img_w + 31) // 32), dtype=torch.int32)
    if not isinstance(boxes, torch.Tensor):
        boxes = boxes.tensor
    assert len(boxes) == n, boxes.shape
    if masks.dim() == 4:
        masks = masks[:, 0]
    return torch.ops.d2b200.paste_masks_packed(masks, boxes, img_h, img_w, float(threshold))


def unpack_mask_bits(packed: torch.Tensor, width: int) -> torch.Tensor:
    """(N, H, ceil(W / 32)) int32 words -> (N, H, W) bool (plain torch ops: works on CPU tensors after the copy)."""
    shifts = torch.arange(32, device=packed.device, dtype=torch.int32)
    bits = (packed.unsqueeze(-1) >> shifts) & 1
    return bits.reshape(packed.shape[0], packed.shape[1], -1)[..., :width].to(torch.bool)
