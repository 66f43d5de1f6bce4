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


def test_grid_push_and_count_match_the_compiled_reference(golden_dir):
    """b200_grid_push against monai._C.grid_push / grid_count of the reference's C++ (fixtures of make_golden.py grid_push_ref).
    The kernel scatters with float atomics, so the comparison is to rounding (1e-5), not bit for bit."""
    from monai_b200.networks.layers import grid_count, grid_push

    g = np.load(os.path.join(golden_dir, "grid_push.npz"))
    names = {0: "replicate", 1: "dct1", 2: "dct2", 3: "dst1", 4: "dst2", 5: "dft", 7: "zero"}
    for i in range(int(g["n"])):
        bound, order, extrap, *shape = (int(v) for v in g[f"c{i}.cfg"])
        x, grid = torch.from_numpy(g[f"c{i}.x"]).cuda(), torch.from_numpy(g[f"c{i}.grid"]).cuda()
        got = grid_push(x, grid, shape, interpolation=order, bound=names[bound], extrapolate=bool(extrap)).cpu().numpy()
        np.testing.assert_allclose(got, g[f"c{i}.y"], rtol=1e-5, atol=4e-6, err_msg=f"case {i}: bound {bound} order {order} extrapolate {extrap}")
    cg = torch.from_numpy(g["count.grid"]).cuda()
    np.testing.assert_allclose(grid_count(cg, (5, 6, 7), interpolation="linear", bound="dct2").cpu().numpy(), g["count.y"], rtol=1e-5, atol=4e-6)
    np.testing.assert_allclose(grid_count(cg.double(), (5, 6, 7), interpolation="linear", bound="dct2").cpu().numpy(), g["count.y"], rtol=1e-5, atol=4e-6)
    # adjointness, at a size the fixture does not cover: <push(x), y> == <x, pull(y)>
    from monai_b200.networks.layers import grid_pull
    gen = torch.Generator(device="cuda").manual_seed(5)
    x = torch.randn((2, 3, 40, 36, 44), device="cuda", generator=gen)
    y = torch.randn((2, 3, 48, 40, 32), device="cuda", generator=gen)
    grid = torch.rand((2, 40, 36, 44, 3), device="cuda", generator=gen) * torch.tensor([52.0, 44.0, 36.0], device="cuda") - 2.0
    for order, bound in ((1, "zero"), (3, "dct2"), (2, "dst1"), (0, "dft")):
        lhs = (grid_push(x, grid, y.shape[2:], interpolation=order, bound=bound).double() * y.double()).sum().item()
        rhs = (x.double() * grid_pull(y, grid, interpolation=order, bound=bound).double()).sum().item()
        assert abs(lhs - rhs) <= 1e-4 * max(1.0, abs(rhs)) + 0.05, (order, bound, lhs, rhs)
    with pytest.raises(ValueError):
        grid_push(x, grid[:, :, :, :5], y.shape[2:])


def test_grid_grad_matches_the_compiled_reference(golden_dir):
    """b200_grid_grad against monai._C.grid_grad of the reference's C++, plus a finite-difference check of grid_pull at a larger size."""
    from monai_b200.networks.layers import grid_grad

    g = np.load(os.path.join(golden_dir, "grid_push.npz"))
    names = {0: "replicate", 1: "dct1", 2: "dct2", 3: "dst1", 4: "dst2", 5: "dft", 7: "zero"}
    for i in range(int(g["n_grad"])):
        bound, order, extrap = (int(v) for v in g[f"g{i}.cfg"])
        x, grid = torch.from_numpy(g[f"g{i}.x"]).cuda(), torch.from_numpy(g[f"g{i}.grid"]).cuda()
        got = grid_grad(x, grid, interpolation=order, bound=names[bound], extrapolate=bool(extrap)).cpu().numpy()
        assert got.shape == g[f"g{i}.y"].shape
        np.testing.assert_allclose(got, g[f"g{i}.y"], rtol=1e-4, atol=2e-5, err_msg=f"case {i}: bound {bound} order {order} extrapolate {extrap}")
    gen = torch.Generator(device="cuda").manual_seed(9)
    x = torch.randn((1, 2, 24, 20, 28), device="cuda", generator=gen)
    grid = (torch.rand((1, 16, 18, 14, 3), device="cuda", generator=gen) * torch.tensor([20.0, 16.0, 24.0], device="cuda") + 1.5).double()
    got = grid_grad(x, grid, interpolation="cubic", bound="dct2")
    h = 1e-3
    for d in range(3):
        e = torch.zeros(3, device="cuda", dtype=torch.float64)
        e[d] = h
        fd = (grid_pull(x, grid + e, interpolation="cubic", bound="dct2") - grid_pull(x, grid - e, interpolation="cubic", bound="dct2")) / (2 * h)
        assert (fd - got[..., d]).abs().max().item() < 5e-2 * max(1.0, got[..., d].abs().max().item())
    # 2-D: the lifted axis is dropped from the result
    x2, g2 = x[:, :, :, :, 0], grid[:, :, :, 0, :2].float()
    assert grid_grad(x2, g2, interpolation="linear", bound="zero").shape == (1, 2, 16, 18, 2)


def test_grid_count_and_grad_reproduce_1d_bp_bwd_rows(golden_dir):
    """tests/testing_data/1D_BP_bwd.txt (the gradient rows of tests/networks/layers/test_grid_pull.py): d/d input of
    grid_pull(...).sum() is grid_count into the input's shape, d/d grid is grid_grad -- 7 bounds x 8 spline orders, 1-D."""
    from monai_b200.networks.layers import grid_count, grid_grad, grid_push

    g = np.load(os.path.join(golden_dir, "grid_push.npz"))
    x = torch.arange(10, dtype=torch.float32, device="cuda").reshape(1, 1, 10)
    grid = (torch.arange(20, dtype=torch.float32, device="cuda") + 0.5).reshape(1, 20, 1)
    assert len(g["bp1d_bwd.labels"]) == 56
    for row, lab in zip(g["bp1d_bwd.rows"], g["bp1d_bwd.labels"]):
        it, bt = str(lab).split()
        kw = dict(interpolation=it.split(".")[1], bound=bt.split(".")[1])
        cnt = grid_count(grid, (10,), **kw)
        assert cnt.shape == (1, 1, 10)
        np.testing.assert_allclose(cnt.cpu().numpy().reshape(-1), row[:10], rtol=1e-4, atol=1e-4, err_msg=f"count {lab}")
        psh = grid_push(torch.ones((1, 1, 20), device="cuda"), grid, (10,), **kw)
        np.testing.assert_allclose(psh.cpu().numpy().reshape(-1), row[:10], rtol=1e-4, atol=1e-4, err_msg=f"push {lab}")
        grd = grid_grad(x, grid, **kw)
        assert grd.shape == (1, 1, 20, 1)
        np.testing.assert_allclose(grd.cpu().numpy().reshape(-1), row[10:], rtol=1e-4, atol=1e-4, err_msg=f"grad {lab}")
