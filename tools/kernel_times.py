"""Per-kernel device time of one eager training hot-path step at real clocks (torch.profiler / CUPTI), averaged over a few
iterations: where the step's time goes kernel by kernel (ncu's own durations are taken at reduced clocks).

    python tools/kernel_times.py [--iters 5] > gpurun_out/kernel_times.txt
"""
import argparse
import collections
import os
import sys

import torch
from torch.profiler import ProfilerActivity, profile

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--iters", type=int, default=5)
    ap.add_argument("--autograd", action="store_true", help="the public-API + torch.autograd step of the end-to-end measurement")
    ap.add_argument("--bf16", action="store_true", help="activations / gradients in bf16 (the end-to-end transport)")
    a = ap.parse_args()
    dev = torch.device("cuda:0")
    r = bench.TrainRunner(dev)
    h = bench.make_train_inputs(0)
    if a.bf16:
        to_half = lambda t: t.to(torch.bfloat16) if t.is_floating_point() else t  # noqa: E731
        h = {k: bench.map_tensors({k: v}, to_half if k in bench.E2E_HALF_KEYS else (lambda t: t))[k] for k, v in h.items()}
    d = r.to_device(h)
    step = (lambda d: r.step_autograd(d, True)) if a.autograd else r.step
    for _ in range(2):
        step(d)
    torch.cuda.synchronize()
    with profile(activities=[ProfilerActivity.CUDA]) as prof:
        for _ in range(a.iters):
            step(d)
        torch.cuda.synchronize()
    tot = collections.OrderedDict()
    for e in prof.events():
        if e.device_type == torch.autograd.DeviceType.CUDA:
            n = e.name
            t, c = tot.get(n, (0.0, 0))
            tot[n] = (t + e.device_time, c + 1)
    total = sum(t for t, _ in tot.values()) / a.iters
    print("kernel time per step: %.1f us over %d kernels" % (total, sum(c for _, c in tot.values()) // a.iters))
    for n, (t, c) in sorted(tot.items(), key=lambda kv: -kv[1][0]):
        print("%9.1f us/step  %5.1f %%  x%-4d avg %8.1f us  %s" % (t / a.iters, 100 * t / a.iters / total, c // a.iters, t / c, n[:110]))


if __name__ == "__main__":
    main()
