"""GPU parity of the dense-grid resampler (b200_grid_pull): the reference's monai._C.grid_pull goldens and outputs of its own
C++ sources (oracle/_ref), and the dense-grid form of Resample against fixtures of the real reference's torch path."""
import os

import numpy as np
import pytest
import torch

from monai_b200.networks.layers import grid_pull
from monai_b200.transforms import Resample
from oracle import resample as ors

pytestmark = pytest.mark.gpu
DEV = "cuda"
INTERP = ["nearest", "linear", "quadratic", "cubic", "fourth", "fifth", "sixth", "seventh"]


def test_grid_pull_reproduces_1d_bp_fwd_rows(golden_dir):
    """tests/networks/layers/test_grid_pull.py with tests/testing_data/1D_BP_fwd.txt: 7 bounds x 8 spline orders, 1-D."""
    g = np.load(os.path.join(golden_dir, "grid_pull.npz"))
    x = torch.arange(10, dtype=torch.float32, device=DEV).reshape(1, 1, 10)
    grid = (torch.arange(20, dtype=torch.float32, device=DEV) + 0.5).reshape(1, 20, 1)
    for row, lab in zip(g["bp1d.rows"], g["bp1d.labels"]):
        it, bt = str(lab).split()
        got = grid_pull(x, grid, interpolation=it.split(".")[1], bound=bt.split(".")[1])
        assert tuple(got.shape) == (1, 1, 20)
        np.testing.assert_allclose(got.cpu().numpy().reshape(-1), row, rtol=1e-4, atol=1e-4, err_msg=str(lab))


def test_grid_pull_matches_the_compiled_reference_3d(golden_dir):
    g = np.load(os.path.join(golden_dir, "grid_pull.npz"))
    x, grid = torch.from_numpy(g["x"]).to(DEV), torch.from_numpy(g["grid"]).to(DEV)
    for bn in ors.BOUNDS:
        for o in range(8):
            got = grid_pull(x, grid, interpolation=o, bound=bn).cpu().numpy()
            np.testing.assert_allclose(got, g[f"y.{bn}.{o}"], rtol=2e-4, atol=3e-5, err_msg=f"{bn} order {o}")
    got = grid_pull(x, grid, interpolation=[3, 1, 2], bound=["dct2", "dft", "dst1"]).cpu().numpy()
    np.testing.assert_allclose(got, g["y.mixed"], rtol=2e-4, atol=3e-5)
    got = grid_pull(x, grid, interpolation="linear", bound="replicate", extrapolate=False).cpu().numpy()
    np.testing.assert_allclose(got, g["y.noextrap"], rtol=2e-4, atol=3e-5)
    # float64 grid (the reference's default coordinate dtype) and an fp16 source
    got64 = grid_pull(x, grid.double(), interpolation="cubic", bound="dct1").cpu().numpy()
    np.testing.assert_allclose(got64, g["y.dct1.3"], rtol=2e-4, atol=3e-5)
    got16 = grid_pull(x.half(), grid, interpolation="linear", bound="zero").float().cpu().numpy()
    np.testing.assert_allclose(got16, g["y.zero.1"], rtol=5e-3, atol=5e-3)
    with pytest.raises(ValueError):
        grid_pull(x, grid, interpolation="linear", bound="nonsense")


def test_dense_grid_resample_matches_the_real_reference(golden_dir):
    """Resample.__call__ with a dense grid tensor (tests/transforms/test_resampler.py cases + random deformation grids for every
    mode / padding / align_corners / norm_coords), fixtures produced by the real reference's torch path."""
    g = np.load(os.path.join(golden_dir, "resampler.npz"))
    for i in range(int(g["n"])):
        mode, pad, align, norm = (str(v) for v in g[f"c{i}.cfg"])
        img, grid = torch.from_numpy(g[f"c{i}.img"]).to(DEV), torch.from_numpy(g[f"c{i}.grid"])
        y = Resample(mode=mode, padding_mode=pad, norm_coords=bool(int(norm)), align_corners=bool(int(align)))(img, grid.to(DEV))
        assert y.dtype == torch.float32 and tuple(y.shape) == tuple(g[f"c{i}.y"].shape)
        diff = np.abs(y.cpu().numpy() - g[f"c{i}.y"])
        if mode == "nearest":   # a coordinate within round-off of .5 may round the other way
            assert (diff > 1e-5).mean() < 5e-3, (i, mode, pad, align, norm)
        else:
            assert diff.max() < 1e-5, (i, mode, pad, align, norm, diff.max())
