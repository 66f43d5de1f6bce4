"""Time b200_resample_affine on the C4 shapes (Spacing 256^3 -> 320^3 and a rotated/scaled 320^3 -> 320^3 resample)."""
import json
import math
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from monai_b200 import _kernels as K  # noqa: E402


def rot(ax, ay, az):
    cx, sx, cy, sy, cz, sz = math.cos(ax), math.sin(ax), math.cos(ay), math.sin(ay), math.cos(az), math.sin(az)
    rx = np.array([[1, 0, 0], [0, cx, -sx], [0, sx, cx]])
    ry = np.array([[cy, 0, sy], [0, 1, 0], [-sy, 0, cy]])
    rz = np.array([[cz, -sz, 0], [sz, cz, 0], [0, 0, 1]])
    return rx @ ry @ rz


def main():
    dev = torch.device("cuda")
    peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json"))) if os.path.exists(os.path.join(ROOT, "MEASURED_PEAKS.json")) else {}
    hbm = peaks.get("hbm_gbs", 6650.0)
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
    cases = []
    m = np.zeros((3, 4)); m[:, :3] = np.eye(3) * 0.8; m[:, 3] = -0.1
    cases.append(("spacing 256^3 -> 320^3 (scale 0.8)", (256,) * 3, (320,) * 3, m))
    r = rot(0.15, -0.1, 0.2) * 1.05
    c = np.array([159.5] * 3)
    m2 = np.zeros((3, 4)); m2[:, :3] = r; m2[:, 3] = c - r @ c + np.array([3.0, -2.0, 1.5])
    cases.append(("rand-affine 320^3 -> 320^3 (rotated)", (320,) * 3, (320,) * 3, m2))
    for name, si, so, mat in cases:
        src = torch.rand((1, *si), device=dev)
        for _ in range(2):
            K.resample_affine(src, so, mat.reshape(-1).tolist(), 1, 1, False)
        torch.cuda.synchronize()
        ms = []
        for _ in range(5):
            flush.fill_(0)
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            K.resample_affine(src, so, mat.reshape(-1).tolist(), 1, 1, False)
            e1.record()
            e1.synchronize()
            ms.append(e0.elapsed_time(e1))
        ms.sort()
        nbytes = (np.prod(si) + np.prod(so)) * 4
        t = ms[len(ms) // 2]
        print(json.dumps({"case": name, "ms": round(t, 4), "algorithmic_MB": round(nbytes / 1e6, 1), "GB/s": round(nbytes / t / 1e6, 1),
                          "frac_of_measured_hbm": round(nbytes / t / 1e6 / hbm, 3)}))


if __name__ == "__main__":
    main()
