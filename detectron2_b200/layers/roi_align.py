"""ROIAlign -- same surface as detectron2/layers/roi_align.py:7-74, backed by d2b200::roi_align."""
from typing import Tuple, Union

import torch
from torch import nn
from torch.nn.modules.utils import _pair

from .. import ops


def roi_align(input: torch.Tensor, boxes: torch.Tensor, output_size: Union[int, Tuple[int, int]],
              spatial_scale: float = 1.0, sampling_ratio: int = -1, aligned: bool = False) -> torch.Tensor:
    """Functional form with torchvision.ops.roi_align's signature (the name detectron2.layers re-exports,
    layers/__init__.py:6).  `boxes` is a K x 5 tensor (batch_idx, x1, y1, x2, y2) or a list of L_i x 4 tensors."""
    if not isinstance(boxes, torch.Tensor):  # list[Tensor[L,4]] -> K x 5, like torchvision convert_boxes_to_roi_format
        ids = torch.cat([torch.full((len(b), 1), i, dtype=b.dtype, device=b.device) for i, b in enumerate(boxes)])
        boxes = torch.cat([ids, torch.cat(list(boxes), dim=0)], dim=1)
    ph, pw = _pair(output_size)
    return ops.roi_align_op(input, boxes.to(dtype=input.dtype), float(spatial_scale), int(ph), int(pw),
                            int(sampling_ratio), bool(aligned))


class ROIAlign(nn.Module):
    def __init__(self, output_size, spatial_scale, sampling_ratio, aligned=True):
        """output_size (h, w); spatial_scale; sampling_ratio (0 = adaptive); aligned: pixel-centre convention
        (shift by -0.5 after scaling), see the reference docstring (roi_align.py:9-35)."""
        super().__init__()
        self.output_size = output_size
        self.spatial_scale = spatial_scale
        self.sampling_ratio = sampling_ratio
        self.aligned = aligned

    def forward(self, input, rois):
        """input: NCHW feature map; rois: K x 5 (batch index, x1, y1, x2, y2)."""
        assert rois.dim() == 2 and rois.size(1) == 5
        if input.is_quantized:
            input = input.dequantize()
        return roi_align(input, rois.to(dtype=input.dtype), self.output_size, self.spatial_scale,
                         self.sampling_ratio, self.aligned)

    def __repr__(self):
        return (f"{self.__class__.__name__}(output_size={self.output_size}, spatial_scale={self.spatial_scale}, "
                f"sampling_ratio={self.sampling_ratio}, aligned={self.aligned})")
