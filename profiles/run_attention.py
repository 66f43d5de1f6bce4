"""Time the window-attention kernels on the SwinUNETR stage shapes (CUDA events, L2 flushed between runs).

    python profiles/run_attention.py [--batch 8] [--iters 5]
One JSON line per (stage, shifted?, kernel): ms, elements of the score matrix per second, share of the MUFU floor."""
import argparse
import json
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from monai_b200 import _kernels as K  # noqa: E402
from monai_b200.networks.nets.swin_unetr import window_plan  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--batch", type=int, default=8)
    ap.add_argument("--iters", type=int, default=5)
    ap.add_argument("--only-tc", action="store_true")
    ap.add_argument("--stages", type=int, default=4, help="time the first k stages only")
    a = ap.parse_args()
    dev = torch.device("cuda")
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
    for name, dims, heads in (("stage1 48^3", (48, 48, 48), 3), ("stage2 24^3", (24, 24, 24), 6), ("stage3 12^3", (12, 12, 12), 12), ("stage4 6^3", (6, 6, 6), 24))[:a.stages]:
        C = heads * 16
        for shifted in (False, True):
            ss = (3, 3, 3) if shifted else (0, 0, 0)
            src, region, nW, n = window_plan(dims, (7, 7, 7), ss)
            if shifted and region is None:
                continue
            qkv = K.NC8(a.batch, 3 * C, (1, nW, n), dev)
            qkv.buf.normal_()
            table = torch.randn(((2 * 7 - 1) ** 3, heads), device=dev) * 0.2
            sched, reps, ntypes = K.window_attention_tc_plan(region, nW, n)
            pb = K.window_attention_tc_pack_bias(table, heads, n, (7, 7, 7), None if reps is None else torch.from_numpy(reps).to(dev), ntypes)
            sched_t = torch.from_numpy(sched).to(dev)
            reg_t = None if region is None else torch.from_numpy(region).to(dev)
            kernels = {"tcgen05": lambda: K.window_attention_tc(qkv, C, heads, nW, n, pb, sched_t, ntypes)}
            if not a.only_tc:
                kernels["mma.sync"] = lambda: K.window_attention_nc8(qkv, C, heads, nW, n, 0.25, table, (7, 7, 7), reg_t)
            for kn, fn in kernels.items():
                for _ in range(2):
                    fn()
                torch.cuda.synchronize()
                ms = []
                for _ in range(a.iters):
                    flush.fill_(0)
                    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                    e0.record(); fn(); e1.record(); e1.synchronize()
                    ms.append(e0.elapsed_time(e1))
                t = sorted(ms)[len(ms) // 2]
                elems = float(a.batch) * nW * heads * n * n
                mufu_floor_ms = elems / (148 * 16 * 1.9e9) * 1e3
                print(json.dumps({"shape": name, "shifted": shifted, "kernel": kn, "batch": a.batch, "windows": nW, "tokens": n, "ms": round(t, 4),
                                  "Gelem/s": round(elems / t / 1e6, 1), "x_mufu_floor": round(t / mufu_floor_ms, 2)}))


if __name__ == "__main__":
    main()
