"""cProfile of the C4 transform pipeline's HOST side (32 volumes of 256^3, kernels asynchronous): where the Python time goes."""
import cProfile
import os
import pstats
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench  # noqa: E402
from monai_b200.data import MetaTensor  # noqa: E402

lazy = len(sys.argv) > 1 and sys.argv[1] == "lazy"
pipe = bench._transform_pipeline(lazy=lazy)
aff = torch.diag(torch.tensor([1.25, 1.25, 1.25, 1.0], dtype=torch.float64))
vols = [torch.rand((1, 256, 256, 256), device="cuda") for _ in range(4)]
for v in vols:
    pipe({"image": MetaTensor(v, affine=aff)})
torch.cuda.synchronize()
pr = cProfile.Profile()
pr.enable()
for _ in range(8):
    for v in vols:
        y = pipe({"image": MetaTensor(v, affine=aff)})["image"]
pr.disable()
torch.cuda.synchronize()
st = pstats.Stats(pr)
st.sort_stats("cumulative").print_stats(45)
