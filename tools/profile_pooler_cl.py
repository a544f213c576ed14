"""ncu driver: the box pooler (1000 RoIs, 7x7) on channels_last features -> roi_align_nhwc_kernel alone, and once on
NCHW features with the layout change forced (nchw_to_nhwc_kernel + roi_align_nhwc_kernel)."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench  # noqa: E402
from detectron2_b200 import ops  # noqa: E402
from detectron2_b200.poolers import ROIPooler  # noqa: E402

dev = torch.device("cuda", 0)
scales = [s for (_, _, s) in bench.LEVELS]
pooler = ROIPooler(7, scales, 0, "ROIAlignV2")
for i in range(3):
    h = bench.make_image_inputs(100 + i)
    nchw = [t.to(dev) for t in h["feats"]]
    cl = [t.contiguous(memory_format=torch.channels_last) for t in nchw]
    boxes = h["proposals"].to(dev)
    ops.POOLER_LAYOUT = "auto"
    pooler(cl, [boxes])
    ops.POOLER_LAYOUT = "nhwc"
    pooler(nchw, [boxes])
torch.cuda.synchronize()
print("done")
