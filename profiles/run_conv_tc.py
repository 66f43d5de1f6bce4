"""Time the tcgen05 3x3x3 convolution on the SwinUNETR hot shapes (CUDA events, L2 flushed between runs).

    python profiles/run_conv_tc.py [--iters 10] [--batch 1]
Prints one JSON line per shape: ms, TFLOP/s and the fraction of the measured bf16 tensor peak (MEASURED_PEAKS.json).
"""
import argparse
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from monai_b200 import _kernels as K  # noqa: E402

SHAPES = [  # (name, Cin, Cout, spatial)
    ("decoder1.conv1 96->48 @96^3", 96, 48, (96, 96, 96)),
    ("encoder1.conv2 48->48 @96^3", 48, 48, (96, 96, 96)),
    ("decoder2.conv1 96->48 @48^3", 96, 48, (48, 48, 48)),
    ("encoder3 96->96 @24^3", 96, 96, (24, 24, 24)),
    ("encoder4 192->192 @12^3", 192, 192, (12, 12, 12)),
    ("decoder5.conv1 768->384 @6^3", 768, 384, (6, 6, 6)),
]


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--iters", type=int, default=10)
    ap.add_argument("--batch", type=int, default=1)
    ap.add_argument("--only", type=int, default=-1)
    a = ap.parse_args()
    peaks = {}
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        peaks = json.load(open(p))
    peak = peaks.get("bf16_tflops", 1590.0)
    dev = torch.device("cuda")
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
    for i, (name, cin, cout, sp) in enumerate(SHAPES):
        if a.only >= 0 and i != a.only:
            continue
        x = K.NC8(a.batch, cin, sp, dev)
        x.buf.normal_()
        w = K.conv3x3x3_tc_pack_weight(torch.randn(cout, cin, 3, 3, 3, device=dev) / (27 * cin) ** 0.5)
        out = K.NC8(a.batch, cout, sp, dev)
        for _ in range(3):
            K.conv3x3x3_tc(x, w, cin, cout, out=out, want_stats=True)
        torch.cuda.synchronize()
        ms = []
        for _ in range(a.iters):
            flush.fill_(0)
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            K.conv3x3x3_tc(x, w, cin, cout, out=out, want_stats=True)
            e1.record()
            e1.synchronize()
            ms.append(e0.elapsed_time(e1))
        ms.sort()
        t = ms[len(ms) // 2]
        flops = 2.0 * a.batch * sp[0] * sp[1] * sp[2] * cin * cout * 27
        tf = flops / (t * 1e-3) / 1e12
        print(json.dumps({"shape": name, "batch": a.batch, "ms_median": round(t, 4), "ms_min": round(ms[0], 4), "tflops": round(tf, 1),
                          "frac_of_measured_bf16_peak": round(tf / peak, 4), "peak_tflops": peak}))


if __name__ == "__main__":
    main()
