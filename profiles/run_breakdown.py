"""Per-shape kernel breakdown of one SwinUNETR window batch (C3: feature_size 48, 96^3 windows, sw_batch_size 4).

    python profiles/run_breakdown.py [--batch 4] [--reps 5] > profiles/r02_breakdown.txt

CUDA-event pair around every C-ABI launch (monai_b200._kernels.profile_start/stop); launches are grouped by entry point and by
their (flops, bytes) signature, i.e. by layer shape.  ms = per window batch, averaged over --reps forward passes.
"""
import argparse
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from monai_b200 import _kernels as K  # noqa: E402
from monai_b200.networks.nets import SwinUNETR  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--batch", type=int, default=4)
    ap.add_argument("--reps", type=int, default=5)
    a = ap.parse_args()
    torch.manual_seed(0)
    net = SwinUNETR(in_channels=1, out_channels=14, feature_size=48).cuda().half().eval()
    x = torch.randn(a.batch, 1, 96, 96, 96, device="cuda").half()
    with torch.no_grad():
        for _ in range(2):
            net(x)
        K.profile_start()
        for _ in range(a.reps):
            net(x)
        prof = K.profile_stop(by_shape=True)
    rows = sorted(prof.items(), key=lambda kv: -kv[1]["ms"])
    total = sum(v["ms"] for _, v in rows)
    print(f"# SwinUNETR fs48, batch {a.batch} x 96^3, {a.reps} reps; total {total / a.reps:.3f} ms per batch")
    print(f"{'ms/batch':>9} {'share':>6} {'n':>4} {'GB/s':>7} {'TF/s':>7}  launch")
    for name, v in rows:
        ms = v["ms"] / a.reps
        n = v["n"] // a.reps
        gbs = v["bytes"] / a.reps / (ms * 1e-3) / 1e9 if ms > 0 else 0
        tfs = v["flops"] / a.reps / (ms * 1e-3) / 1e12 if ms > 0 else 0
        print(f"{ms:9.3f} {100 * v['ms'] / total:5.1f}% {n:4d} {gbs:7.0f} {tfs:7.1f}  {name}")


if __name__ == "__main__":
    main()
