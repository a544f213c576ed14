"""Tiny driver for ncu: a few eager hot-path steps (no CUDA graphs) so that -k / -s / -c select kernels simply."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench  # noqa: E402

dev = torch.device("cuda", 0)
runner = bench.OursRunner(dev)
ins = [runner.to_device(bench.make_image_inputs(i)) for i in range(3)]
n = int(sys.argv[1]) if len(sys.argv) > 1 else 4
for i in range(n):
    runner.step(ins[i % 3])
torch.cuda.synchronize()
print("done")
