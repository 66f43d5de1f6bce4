"""Time the fused gather-form blend (mode 0) on the C2 / C3 window tables (CUDA events, L2 flushed)."""
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from monai_b200 import _kernels as K  # noqa: E402
from monai_b200.data.utils import dense_patch_starts, importance_factors  # noqa: E402


def main():
    dev = torch.device("cuda")
    peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json"))) if os.path.exists(os.path.join(ROOT, "MEASURED_PEAKS.json")) else {}
    hbm = peaks.get("hbm_gbs", 6650.0)
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
    for name, vol, roi, C in (("C2 256^3", (256, 256, 256), (96, 96, 96), 2), ("C3 512^3", (512, 512, 512), (96, 96, 96), 2)):
        starts = dense_patch_starts(vol, roi, (48, 48, 48))
        n = len(starts[0]) * len(starts[1]) * len(starts[2])
        preds = torch.randn((n, C, *roi), device=dev, dtype=torch.float16)
        st = [torch.tensor(s, dtype=torch.int32, device=dev) for s in starts]
        st[2]._align = 8
        st[2]._max_cover = 3   # roi 96 / interval 48 with a snapped last window: at most three windows cover a voxel per axis
        f, clamp = importance_factors(roi, "gaussian", 0.125)
        f = [t.to(dev) for t in f]
        out = torch.empty((1, C, *vol), device=dev, dtype=torch.float16)
        for _ in range(2):
            K.sw_blend(0, preds, 0, n, (1, C, *vol), roi, st, f, clamp, None, out)
        torch.cuda.synchronize()
        ms = []
        for _ in range(5):
            flush.fill_(0)
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            K.sw_blend(0, preds, 0, n, (1, C, *vol), roi, st, f, clamp, None, out)
            e1.record()
            e1.synchronize()
            ms.append(e0.elapsed_time(e1))
        ms.sort()
        nbytes = preds.numel() * 2 + out.numel() * 2
        t = ms[len(ms) // 2]
        print(json.dumps({"case": name, "windows": n, "ms": round(t, 4), "algorithmic_MB": round(nbytes / 1e6, 1), "GB/s": round(nbytes / t / 1e6, 1),
                          "frac_of_measured_hbm": round(nbytes / t / 1e6 / hbm, 3)}))


if __name__ == "__main__":
    main()
