"""Two SwinUNETR (feature_size 48) forward passes of a 4 x 96^3 window batch, for ncu captures of the non-conv kernels:

    ncu --set full --clock-control none --import-source on \
        -k regex:"conv_cin1_tc|mlp_fused|window_attention_tc|norm_act|head_conv_norm" --launch-skip 33 --launch-count 33 \
        -o gpurun_out/r02_swin_kernels python profiles/run_ncu_forward.py
    python profiles/summarize_ncu.py gpurun_out/r02_swin_kernels.ncu-rep > profiles/r02_swin_kernels_ncu_summary.txt

(the first pass -- 33 matching launches -- warms up and is skipped).  An optional argument sets the batch (default 4; bench.py runs 25).
"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from monai_b200.networks.nets import SwinUNETR  # noqa: E402

torch.manual_seed(0)
net = SwinUNETR(in_channels=1, out_channels=14, feature_size=48).cuda().half().eval()
BATCH = int(sys.argv[1]) if len(sys.argv) > 1 else 4
x = torch.randn(BATCH, 1, 96, 96, 96, device="cuda").half()
with torch.no_grad():
    for _ in range(2):
        y = net(x)
torch.cuda.synchronize()
print(float(y.float().abs().mean()))
