"""ROIAlignRotated -- same surface as detectron2/layers/roi_align_rotated.py:11-103."""
import torch
from torch import nn
from torch.nn.modules.utils import _pair

from .. import ops


def roi_align_rotated(input, roi, output_size, spatial_scale, sampling_ratio):
    """Differentiable (w.r.t. input) rotated RoIAlign; `roi` is K x 6 (batch idx, cx, cy, w, h, angle degrees).
    Mirrors `_ROIAlignRotated.apply` (roi_align_rotated.py:11-48)."""
    ph, pw = _pair(output_size)
    return ops.roi_align_rotated_op(input, roi, float(spatial_scale), int(ph), int(pw), int(sampling_ratio))


class ROIAlignRotated(nn.Module):
    def __init__(self, output_size, spatial_scale, sampling_ratio):
        super().__init__()
        self.output_size = output_size
        self.spatial_scale = spatial_scale
        self.sampling_ratio = sampling_ratio

    def forward(self, input, rois):
        assert rois.dim() == 2 and rois.size(1) == 6
        orig_dtype = input.dtype
        if orig_dtype == torch.float16:  # roi_align_rotated.py:81-83
            input = input.float()
            rois = rois.float()
        return roi_align_rotated(input, rois, self.output_size, self.spatial_scale, self.sampling_ratio).to(
            dtype=orig_dtype)

    def __repr__(self):
        return (f"{self.__class__.__name__}(output_size={self.output_size}, spatial_scale={self.spatial_scale}, "
                f"sampling_ratio={self.sampling_ratio})")
