"""GPU: gradients of grid_pull / grid_push / grid_count (autograd Functions composed from the four forward kernels) against the
reference's own gradient goldens (tests/testing_data/1D_BP_bwd.txt, all 224 rows of tests/networks/layers/test_grid_pull.py) and against
monai._C.grid_{pull,push,count}_backward of the compiled reference (fixtures of tests/golden/make_golden.py grid_push_ref)."""
import os

import numpy as np
import pytest
import torch

from monai_b200.networks.layers import grid_count, grid_grad, grid_pull, grid_push

pytestmark = pytest.mark.gpu
NAMES = {0: "replicate", 1: "dct1", 2: "dct2", 3: "dst1", 4: "dst2", 5: "dft", 7: "zero"}


def test_grid_pull_gradients_reproduce_all_1d_bp_bwd_rows(golden_dir):
    """The reference's test_grid_pull, verbatim in structure: per (bound, interpolation) four rows -- input and grid require grad
    (10 + 20 values), input only, grid only, neither (a single 0)."""
    g = np.load(os.path.join(golden_dir, "grid_push.npz"))
    rows, labels = g["bp1d_bwd.all_rows"], g["bp1d_bwd.all_labels"]
    assert len(rows) == 224
    for i in range(0, 224, 4):
        it, bt = str(labels[i]).split()
        for j, (input_g, grid_g) in enumerate(((True, True), (True, False), (False, True), (False, False))):
            assert str(labels[i + j]) == str(labels[i])
            want = rows[i + j][~np.isnan(rows[i + j])]
            x = torch.arange(10, dtype=torch.float32, device="cuda").reshape(1, 1, 10).requires_grad_(input_g)
            base = torch.arange(20, dtype=torch.float32, device="cuda").reshape(1, 20, 1).requires_grad_(grid_g)
            result = grid_pull(x, base + 0.5, interpolation=it.split(".")[1], bound=bt.split(".")[1])
            grads = []
            if input_g or grid_g:
                result.sum().backward()
            if input_g:
                grads.append(x.grad.view(-1))
            if grid_g:
                grads.append(base.grad.view(-1))
            got = torch.cat(grads).cpu().numpy() if grads else np.zeros(1)
            np.testing.assert_allclose(got, want, rtol=1e-4, atol=1e-4, err_msg=f"{labels[i]} input_g={input_g} grid_g={grid_g}")


def test_backward_of_pull_push_count_matches_the_compiled_reference(golden_dir):
    g = np.load(os.path.join(golden_dir, "grid_push.npz"))
    for i in range(int(g["n_bwd"])):
        bound, order, extrap = (int(v) for v in g[f"b{i}.cfg"])
        kw = dict(interpolation=order, bound=NAMES[bound], extrapolate=bool(extrap))
        tol = dict(rtol=1e-4, atol=2e-5, err_msg=f"case {i}: bound {bound} order {order} extrapolate {extrap}")
        t = lambda k: torch.from_numpy(g[f"b{i}.{k}"]).cuda()   # noqa: E731
        x, grid = t("x").requires_grad_(), t("grid").requires_grad_()
        grid_pull(x, grid, **kw).backward(t("gout"))
        np.testing.assert_allclose(x.grad.cpu().numpy(), g[f"b{i}.pull_dx"], **tol)
        np.testing.assert_allclose(grid.grad.cpu().numpy(), g[f"b{i}.pull_dg"], **tol)
        xin, grid = t("xin").requires_grad_(), t("grid").requires_grad_()
        grid_push(xin, grid, (6, 5, 7), **kw).backward(t("gvol"))
        np.testing.assert_allclose(xin.grad.cpu().numpy(), g[f"b{i}.push_dx"], **tol)
        np.testing.assert_allclose(grid.grad.cpu().numpy(), g[f"b{i}.push_dg"], **tol)
        grid = t("grid").requires_grad_()
        grid_count(grid, (6, 5, 7), **kw).backward(t("gcnt"))
        np.testing.assert_allclose(grid.grad.cpu().numpy(), g[f"b{i}.count_dg"], **tol)


def test_grid_grad_is_forward_only_and_says_so():
    x = torch.randn((1, 1, 6, 5, 4), device="cuda", requires_grad=True)
    grid = torch.rand((1, 3, 3, 3, 3), device="cuda") * 3
    with pytest.raises(NotImplementedError):
        grid_grad(x, grid)
    with torch.no_grad():
        assert grid_grad(x, grid).shape == (1, 1, 3, 3, 3, 3)
    assert not grid_pull(x.detach(), grid).requires_grad
