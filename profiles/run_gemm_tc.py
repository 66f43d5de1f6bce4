"""Time the tcgen05 GEMM on the SwinUNETR linear-layer shapes (CUDA events, L2 flushed between runs)."""
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from monai_b200 import _kernels as K  # noqa: E402
from monai_b200 import _lib as L  # noqa: E402

SHAPES = [  # name, rows per batch item, K, N, act, residual
    ("stage1.qkv 48->144", 343 * 343, 48, 144, 0, False),
    ("stage1.proj 48->48", 343 * 343, 48, 48, 0, True),
    ("stage1.fc1 48->192 gelu", 48**3, 48, 192, L.ACT_GELU, False),
    ("stage1.fc1 48->192 (no gelu)", 48**3, 48, 192, 0, False),
    ("stage1.fc2 192->48 (no res)", 48**3, 192, 48, 0, False),
    ("stage1.fc2 192->48 +res", 48**3, 192, 48, 0, True),
    ("stage1.merge 384->96", 24**3, 384, 96, 0, False),
    ("stage2.fc1 96->384 gelu", 24**3, 96, 384, L.ACT_GELU, False),
    ("stage4.fc1 384->1536 gelu", 6**3, 384, 1536, L.ACT_GELU, False),
    ("decoder1.up 48->8x48", 48**3, 48, 8 * 48, 0, False),
    ("decoder1.conv3 1x1 96->48 @96^3", 96**3, 96, 48, 0, False),
    ("decoder1.conv3 1x1 96->48 @96^3 + InstanceNorm statistics (as the network runs it)", 96**3, 96, 48, 0, False, True),
]


def main():
    dev = torch.device("cuda")
    batch = int(sys.argv[1]) if len(sys.argv) > 1 else 4
    only = int(sys.argv[2]) if len(sys.argv) > 2 else -1
    peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json"))) if os.path.exists(os.path.join(ROOT, "MEASURED_PEAKS.json")) else {}
    hbm = peaks.get("hbm_gbs", 6650.0)
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
    for idx, (name, S, Kd, N, act, res, *rest) in enumerate(SHAPES):
        stats = bool(rest and rest[0])
        if only >= 0 and idx != only:
            continue
        x = K.NC8(batch, Kd, (1, 1, S), dev)
        x.buf.normal_()
        w = K.gemm_tc_pack_weight(torch.randn(N, Kd, device=dev) / Kd**0.5)
        bias = torch.randn(N, device=dev)
        out = K.NC8(batch, N, (1, 1, S), dev)
        r = K.NC8(batch, N, (1, 1, S), dev) if res else None
        if r is not None:
            r.buf.normal_()
        for _ in range(3):
            K.gemm_tc(x, w, Kd, N, bias=None if stats else bias, out=out, res=r, act=act, want_stats=stats)
        torch.cuda.synchronize()
        ms = []
        for _ in range(8):
            flush.fill_(0)
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            K.gemm_tc(x, w, Kd, N, bias=None if stats else bias, out=out, res=r, act=act, want_stats=stats)
            e1.record()
            e1.synchronize()
            ms.append(e0.elapsed_time(e1))
        ms.sort()
        t = ms[len(ms) // 2]
        nbytes = batch * S * (Kd + N * (2 if res else 1)) * 2
        print(json.dumps({"shape": name, "batch": batch, "ms": round(t, 4), "GB/s": round(nbytes / t / 1e6, 1), "frac_hbm": round(nbytes / t / 1e6 / hbm, 3),
                          "tflops": round(2.0 * batch * S * Kd * N / t / 1e9, 1)}))


if __name__ == "__main__":
    main()
