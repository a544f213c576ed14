"""A/B timing of the channels-last ROIPooler forward (box head 1024 RoIs 7x7, mask head 256 RoIs 14x14, cfg1 single level) in
CUDA graphs with rotating inputs.  (profiles/r2_pooler_fwd_ab.md was taken with a build that selected the kernel variant
through an environment knob; the shipped kernel chooses per RoI.)
    python tools/bench_pooler_fwd.py
"""
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402
from detectron2_b200 import ops  # noqa: E402


def main():
    dev = torch.device("cuda", 0)
    runner = bench.TrainRunner(dev)
    host = [bench.make_train_inputs(s) for s in (0, 1)]
    devin = [runner.to_device(h) for h in host]
    cls = [ops.pyramid_to_channels_last(d["feats"]) for d in devin]
    side = torch.cuda.Stream()
    res = {}
    # cfg1: 512 boxes over one 1x256x200x304 map (channels_last), sr = 0 and 2
    g = torch.Generator().manual_seed(0)
    x = [torch.rand(1, 256, 200, 304, generator=g).to(dev).contiguous(memory_format=torch.channels_last) for _ in range(2)]
    cx, cy = torch.rand(512, generator=g) * 1216, torch.rand(512, generator=g) * 800
    w, h = 16 + torch.rand(512, generator=g) * 300, 16 + torch.rand(512, generator=g) * 300
    rois = torch.stack([torch.zeros(512), (cx - w / 2).clamp(0, 1216), (cy - h / 2).clamp(0, 800), (cx + w / 2).clamp(0, 1216),
                        (cy + h / 2).clamp(0, 800)], 1).to(dev)
    stages = {
        "box_pool_fwd": lambda b: runner.pool_fwd(devin[b], "box", cls[b]),
        "mask_pool_fwd": lambda b: runner.pool_fwd(devin[b], "mask", cls[b]),
        "cfg1_sr0": lambda b: ops.roi_align_op(x[b], rois, 0.25, 7, 7, 0, True),
        "cfg1_sr2": lambda b: ops.roi_align_op(x[b], rois, 0.25, 7, 7, 2, True),
    }
    for name, fn in stages.items():
        for b in range(2):
            fn(b)
        torch.cuda.synchronize()
        sg = [bench.graph_of(lambda b=b: fn(b), side) for b in range(2)]
        torch.cuda.synchronize()
        res[name + "_us"] = round(bench.time_graphs([g_ for g_, _ in sg], 20) * 1e3, 2)
    print(json.dumps(res))


if __name__ == "__main__":
    main()
