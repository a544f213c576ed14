"""Driver for ncu captures of the training hot path's kernels: a few eager passes over one input set.

    ncu --set full --clock-control none --import-source on -k regex:'dcn_|roi_align_bwd_nhwc|roi_align_nhwc' -s <skip> -c <n> \
        -o gpurun_out/prof python tools/profile_train.py [--iters 2] [--what pool,dconv,nms]
"""
import argparse
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--iters", type=int, default=2)
    ap.add_argument("--what", default="pool,dconv,nms")
    a = ap.parse_args()
    dev = torch.device("cuda:0")
    r = bench.TrainRunner(dev)
    d = r.to_device(bench.make_train_inputs(0))
    torch.cuda.synchronize()
    what = a.what.split(",")
    for _ in range(a.iters):
        if "nms" in what:
            r.rpn_nms(d)
        if "pool" in what:
            cl = r.ops.pyramid_to_channels_last(d["feats"])
            _, rb = r.pool_fwd(d, "box", cl)
            _, rm = r.pool_fwd(d, "mask", cl)
            r.pool_bwd(d, "box", rb, True)
            r.pool_bwd(d, "mask", rm, True)
        if "dconv" in what:
            for si in range(3):
                r.dconv_bwd(d, si, 0, r.dconv_fwd(d, si, 0))
        torch.cuda.synchronize()
    print("done")


if __name__ == "__main__":
    main()
