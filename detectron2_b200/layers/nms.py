"""nms / batched_nms / nms_rotated / batched_nms_rotated -- same surface as detectron2/layers/nms.py.

Differences in mechanism (not results): the reference's batched variants add per-category coordinate offsets in
Python and call a single-class NMS; here the offsets are applied inside the gather kernel (same fp32 arithmetic),
the IoU bitmask is reduced on the device, and `*_fixed` variants return (padded keep, count) without a host sync.
"""
import torch

from .. import ops

# torchvision switches batched_nms to a per-class Python loop above this size on GPU (ops/boxes.py); kept for parity
_TRICK_MAX_NUMEL = 100_000


def nms(boxes: torch.Tensor, scores: torch.Tensor, iou_threshold: float) -> torch.Tensor:
    """torchvision.ops.nms semantics (layers/nms.py:6): keep indices sorted by decreasing score, suppress IoU > thr."""
    return torch.ops.d2b200.nms(boxes, scores, None, float(iou_threshold), False, True)


def batched_nms(boxes: torch.Tensor, scores: torch.Tensor, idxs: torch.Tensor, iou_threshold: float) -> torch.Tensor:
    """Per-category NMS (layers/nms.py:11-22 -> torchvision batched_nms, always on boxes.float()).  Scriptable like the
    reference (tests/layers/test_nms.py:16-29): the body is dispatcher ops only."""
    assert boxes.shape[-1] == 4
    boxes = boxes.float()
    if boxes.numel() == 0:
        return torch.empty((0,), dtype=torch.int64, device=boxes.device)
    # torchvision leaves the coordinate-offset trick for a per-class Python loop above 25 000 boxes (ops/boxes.py
    # _batched_nms_vanilla: raw coordinates, one nms + host sync per class).  Same results in ONE call here: the
    # categories only segment the boxes (D2B_NMS_NO_OFFSET).
    return torch.ops.d2b200.nms(boxes, scores, idxs, float(iou_threshold), False, boxes.numel() <= 100000)


def batched_nms_fixed(boxes, scores, idxs, iou_threshold):
    """Sync-free variant: (keep[M] padded, num_keep[1]) device tensors."""
    return ops.nms_fixed(boxes.float(), scores, idxs, float(iou_threshold), False)


def batched_nms_images_fixed(boxes, scores, idxs, iou_threshold, num_categories: int, max_segment: int = 0):
    """`batched_nms` of several images in ONE NMS call, sync-free: boxes [N, M, 4] (or a list of N [M, 4] tensors), scores
    [N, M], idxs [M] or [N, M] category ids in [0, num_categories).  Same kept sets, in the same order per image, as N calls of
    `batched_nms` (the reference's per-image loop, proposal_utils.py:96-133): every image keeps torchvision's own coordinate
    offsets -- category * (max coordinate of THAT image + 1), fp32 -- and its categories are made disjoint from the other
    images' (image * num_categories + category).  Returns (keep [N*M] padded, num_keep [1]): flat indices image * M + box in
    descending score order over all images; `keep // M` is the image.  max_segment: bound on the boxes per (image, category)."""
    if isinstance(boxes, (list, tuple)):
        boxes = torch.stack([b.float() for b in boxes])
    if isinstance(scores, (list, tuple)):
        scores = torch.stack(list(scores))
    boxes = boxes.float()
    n, m = boxes.shape[0], boxes.shape[1]
    idxs = idxs if idxs.dim() == 2 else idxs[None, :].expand(n, m)
    if boxes.numel() == 0:
        return (torch.zeros((n * m,), dtype=torch.int64, device=boxes.device),
                torch.zeros((1,), dtype=torch.int64, device=boxes.device))
    if m * 4 > _TRICK_MAX_NUMEL:  # torchvision's per-class loop on raw coordinates above 25 000 boxes per image
        nms_boxes = boxes
    else:  # torchvision _batched_nms_coordinate_trick, per image
        mx = boxes.reshape(n, -1).max(dim=1).values
        nms_boxes = boxes + (idxs.to(boxes.dtype) * (mx[:, None] + 1.0))[..., None]
    cat = idxs + torch.arange(n, device=boxes.device, dtype=idxs.dtype)[:, None] * int(num_categories)
    return ops.nms_fixed(nms_boxes.reshape(-1, 4), scores.reshape(-1), cat.reshape(-1), float(iou_threshold), False,
                         apply_offsets=False, max_segment=int(max_segment))


def nms_rotated(boxes: torch.Tensor, scores: torch.Tensor, iou_threshold: float) -> torch.Tensor:
    """Rotated NMS over (cx, cy, w, h, angle_deg) boxes (layers/nms.py:28-89); suppress IoU >= thr like the
    reference CPU kernel (nms_rotated_cpu.cpp:54)."""
    return torch.ops.detectron2.nms_rotated(boxes, scores, iou_threshold)


def batched_nms_rotated(boxes: torch.Tensor, scores: torch.Tensor, idxs: torch.Tensor, iou_threshold: float) -> torch.Tensor:
    """Per-category rotated NMS (layers/nms.py:97-147); the min/max-coordinate offsets of :137-146 are computed and
    applied inside the kernel pipeline in fp32."""
    assert boxes.shape[-1] == 5
    if boxes.numel() == 0:
        return torch.empty((0,), dtype=torch.int64, device=boxes.device)
    return torch.ops.d2b200.nms(boxes.float(), scores, idxs, float(iou_threshold), True, True)
