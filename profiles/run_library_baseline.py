"""The "honest bar" of SURVEY.md section 8(d): the same SwinUNETR forward on the SAME B200 through PyTorch's library
kernels (cuDNN / cuBLAS, fp16, channels-first eager) -- the torch functional restatement in oracle/networks.py moved to
the GPU.  Measurement infrastructure only (never imported by the product).  Prints one JSON line:
ms per window batch, windows/s, and the voxels/s a 512^3 volume (1000 windows of 96^3) would run at on that path.

    python profiles/run_library_baseline.py [--batch 4] [--iters 5]
"""
import argparse
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests", "golden"))
from oracle import networks as onet  # noqa: E402
from weights import fill_state_dict  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--batch", type=int, default=4)
    ap.add_argument("--iters", type=int, default=5)
    a = ap.parse_args()
    from monai_b200.networks.nets import SwinUNETR

    dev = torch.device("cuda")
    sd = fill_state_dict(SwinUNETR(in_channels=1, out_channels=2, feature_size=48).state_dict(), 1)
    sd = {k: (v.to(dev).half() if v.is_floating_point() else v.to(dev)) for k, v in sd.items()}
    x = torch.randn((a.batch, 1, 96, 96, 96), device=dev, dtype=torch.float16)
    torch.backends.cudnn.benchmark = True
    with torch.no_grad():
        for _ in range(3):
            y = onet.swin_unetr_forward(sd, x)
        torch.cuda.synchronize()
        ms = []
        for _ in range(a.iters):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            y = onet.swin_unetr_forward(sd, x)
            e1.record()
            e1.synchronize()
            ms.append(e0.elapsed_time(e1))
    ms.sort()
    t = ms[len(ms) // 2]
    per_win = t / a.batch
    print(json.dumps({"what": "SwinUNETR fs48 forward, torch eager fp16 (cuDNN/cuBLAS) on this GPU", "batch": a.batch, "ms_per_batch": round(t, 3),
                      "ms_per_window": round(per_win, 3), "voxels_per_s_512cube_1000_windows": 512.0**3 / (per_win * 1000 * 1e-3),
                      "finite": bool(torch.isfinite(y.float()).all())}))


if __name__ == "__main__":
    main()
