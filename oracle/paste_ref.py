"""CPU port of the reference's paste_masks_in_image CPU branch (TEST / BASELINE INFRASTRUCTURE).

Restates detectron2/layers/mask_ops.py:17-69 (_do_paste_mask, skip_empty=True) and :110-147 (one mask per chunk on
CPU) with torch ops, because the reference's Python file cannot travel to the GPU box.  Pinned against the real
reference function in tests/test_oracle_pins.py::test_paste_port_matches_reference (authoring container only) and
against the golden fixture everywhere.
"""
import torch
import torch.nn.functional as F


def _paste_one(mask, box, img_h, img_w):
    # region that tightly bounds the box (mask_ops.py:38-44)
    x0i = int(torch.clamp(box[0].floor() - 1, min=0).item())
    y0i = int(torch.clamp(box[1].floor() - 1, min=0).item())
    x1i = int(torch.clamp(box[2].ceil() + 1, max=img_w).item())
    y1i = int(torch.clamp(box[3].ceil() + 1, max=img_h).item())
    x0, y0, x1, y1 = box[0], box[1], box[2], box[3]
    ys = torch.arange(y0i, y1i, dtype=torch.float32) + 0.5
    xs = torch.arange(x0i, x1i, dtype=torch.float32) + 0.5
    ys = (ys - y0) / (y1 - y0) * 2 - 1
    xs = (xs - x0) / (x1 - x0) * 2 - 1
    gx = xs[None, :].expand(ys.numel(), xs.numel())
    gy = ys[:, None].expand(ys.numel(), xs.numel())
    grid = torch.stack([gx, gy], dim=2)[None]
    out = F.grid_sample(mask[None, None].float(), grid, align_corners=False)
    return out[0, 0], (slice(y0i, y1i), slice(x0i, x1i))


def paste_masks_in_image_cpu(masks, boxes, image_shape, threshold=0.5):
    n = len(masks)
    img_h, img_w = image_shape
    res = torch.zeros(n, img_h, img_w, dtype=torch.bool if threshold >= 0 else torch.uint8)
    for i in range(n):
        soft, (sy, sx) = _paste_one(masks[i], boxes[i], img_h, img_w)
        res[i, sy, sx] = (soft >= threshold) if threshold >= 0 else (soft * 255).to(torch.uint8)
    return res
