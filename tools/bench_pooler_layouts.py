"""Device time of the ROIPooler forward per feature-map layout on bench.py's synthetic image (3 rotating inputs,
one CUDA graph per input, like bench.py's per-stage timing):
   nchw   NCHW features -> roi_align_v3_kernel
   cl     channels_last features consumed in place -> roi_align_nhwc_kernel
   xpose  NCHW features -> nchw_to_nhwc_kernel + roi_align_nhwc_kernel
Usage: python tools/bench_pooler_layouts.py   (prints one JSON line)"""
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench  # noqa: E402
from detectron2_b200 import ops  # noqa: E402
from detectron2_b200.poolers import ROIPooler  # noqa: E402


def timed(fn_per_buf, rep=20):
    side = torch.cuda.Stream()
    graphs, keep = [], []
    with torch.cuda.stream(side):
        for fn in fn_per_buf:
            fn()
        torch.cuda.synchronize()
        for fn in fn_per_buf:
            g = torch.cuda.CUDAGraph()
            with torch.cuda.graph(g, stream=side):
                keep.append(fn())
            graphs.append(g)
    torch.cuda.synchronize()
    for i in range(3):
        graphs[i % len(graphs)].replay()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize()
    a.record()
    for i in range(rep):
        graphs[i % len(graphs)].replay()
    b.record()
    torch.cuda.synchronize()
    return a.elapsed_time(b) / rep * 1e3  # us


def main():
    dev = torch.device("cuda:0")
    scales = [s for (_, _, s) in bench.LEVELS]
    host = [bench.make_image_inputs(100 + i) for i in range(3)]
    res = {}
    for name, k, out in (("box_1000x7x7", 1000, 7), ("mask_100x14x14", 100, 14), ("box_512x7x7", 512, 7),
                         ("box_2000x7x7", 2000, 7)):
        pooler = ROIPooler(out, scales, 0, "ROIAlignV2")
        g = torch.Generator().manual_seed(k)
        boxes = [bench.synth_boxes(g, k).to(dev) for _ in host]
        nchw = [[t.to(dev) for t in h["feats"]] for h in host]
        cl = [[t.contiguous(memory_format=torch.channels_last) for t in f] for f in nchw]
        r = {}
        for mode in ("nchw", "cl", "xpose"):
            ops.POOLER_LAYOUT = {"nchw": "nchw", "cl": "auto", "xpose": "nhwc"}[mode]
            feats = cl if mode == "cl" else nchw
            r[mode + "_us"] = round(timed([(lambda f=f, b=b: pooler(f, [b])) for f, b in zip(feats, boxes)]), 1)
        ops.POOLER_LAYOUT = "auto"
        r["auto_picks"] = ops._pick_layout(nchw[0], k * bench.C * out * out)
        res[name] = r
        del nchw, cl
    print(json.dumps(res))


if __name__ == "__main__":
    main()
