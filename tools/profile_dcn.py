"""ncu driver: deformable-conv forward on the tensor-core path at the R50 res3 shape."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from detectron2_b200 import ops  # noqa: E402

prec = int(sys.argv[1]) if len(sys.argv) > 1 else 1
cin, hh, ww = 128, 100, 168
xx = torch.randn(2, cin, hh, ww, device="cuda")
off = torch.randn(2, 18, hh, ww, device="cuda") * 2
wt = torch.randn(cin, cin, 3, 3, device="cuda") * 0.05
for _ in range(3):
    ops.deform_conv_op(xx, off, None, wt, None, [1, 1], [1, 1], [1, 1], 1, 1, prec)
torch.cuda.synchronize()
print("done")
