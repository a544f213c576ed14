"""pairwise_iou_rotated -- same surface as detectron2/layers/rotated_boxes.py:6-21."""
import torch


def pairwise_iou_rotated(boxes1, boxes2):
    """IoU matrix [N, M] of two sets of (cx, cy, w, h, angle_deg) boxes."""
    return torch.ops.detectron2.box_iou_rotated(boxes1, boxes2)
