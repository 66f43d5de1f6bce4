"""`grid_pull` with the reference's call contract (monai/networks/layers/spatial_transforms.py:35-132), backed by the CUDA
kernel b200_grid_pull (no monai._C needed): spline orders 0-7, the seven boundary conditions of the discrete transforms,
per-axis settings given in the reference's [W, H, D] order (= the order of the grid's last dimension).  grid_pull / grid_push /
grid_count are differentiable (their backward passes are compositions of the same four forward kernels); grid_grad is forward only."""
from __future__ import annotations

from typing import Sequence

import numpy as np
import torch

from ... import _kernels as K

__all__ = ["AffineTransform", "grid_pull", "grid_push", "grid_count", "grid_grad"]


def _codes(v, table: dict, what: str, n: int) -> list[int]:
    vals = list(v) if isinstance(v, (list, tuple)) else [v]
    out = []
    for x in vals:
        x = getattr(x, "name", x)   # enum members of the reference (BoundType.dct2, InterpolationType.linear) by name
        if isinstance(x, str):
            key = x.lower()
            if key not in table:
                raise ValueError(f"unknown {what} {x!r}; options: {sorted(table)}")
            out.append(table[key])
        else:
            out.append(int(x))
    if len(out) == 1:
        out = out * n
    if len(out) < n:
        out = out + [out[-1]] * (n - len(out))
    return out[:n]


class _GridPull(torch.autograd.Function):
    """grid_pull with the gradients of monai._C.grid_pull_backward, composed from the forward kernels of the adjoint operators:
    d/d input = grid_push(grad) into the input's shape; d/d grid = sum_c grad * grid_grad(input) (pushpull_cpu.cpp, do_push / do_grad
    branches; checked against the compiled reference's backward entry points and its 1D_BP_bwd.txt rows)."""

    @staticmethod
    def forward(ctx, x3, g3, bnd, order, extrapolate):
        ctx.opt = (bnd, order, extrapolate)
        ctx.save_for_backward(x3, g3)
        return K.grid_pull(x3, g3, bnd, order, extrapolate=extrapolate, channel_last=True)

    @staticmethod
    def backward(ctx, grad):
        x3, g3 = ctx.saved_tensors
        bnd, order, extrapolate = ctx.opt
        grad = grad.contiguous()
        gi = gg = None
        if ctx.needs_input_grad[0]:
            gi = K.grid_push(grad, g3, x3.shape[2:], bnd, order, extrapolate=extrapolate).to(x3.dtype)
        if ctx.needs_input_grad[1]:
            gg = (K.grid_grad(x3, g3, bnd, order, extrapolate=extrapolate).float() * grad.float().unsqueeze(-1)).sum(1).to(g3.dtype)
        return gi, gg, None, None, None


class _GridPush(torch.autograd.Function):
    """grid_push; backward (monai._C.grid_push_backward): d/d input = grid_pull(grad), d/d grid = sum_c input * grid_grad(grad)."""

    @staticmethod
    def forward(ctx, x3, g3, shape3, bnd, order, extrapolate):
        ctx.opt = (bnd, order, extrapolate)
        ctx.save_for_backward(x3, g3)
        return K.grid_push(x3, g3, shape3, bnd, order, extrapolate=extrapolate)

    @staticmethod
    def backward(ctx, grad):
        x3, g3 = ctx.saved_tensors
        bnd, order, extrapolate = ctx.opt
        grad = grad.contiguous()
        gi = gg = None
        if ctx.needs_input_grad[0]:
            gi = K.grid_pull(grad, g3, bnd, order, extrapolate=extrapolate, channel_last=True).to(x3.dtype)
        if ctx.needs_input_grad[1]:
            gg = (K.grid_grad(grad, g3, bnd, order, extrapolate=extrapolate).float() * x3.float().unsqueeze(-1)).sum(1).to(g3.dtype)
        return gi, gg, None, None, None, None


class _GridCount(torch.autograd.Function):
    """grid_count; backward (monai._C.grid_count_backward): d/d grid = grid_grad(grad) of the single channel."""

    @staticmethod
    def forward(ctx, g3, shape3, bnd, order, extrapolate):
        ctx.opt = (bnd, order, extrapolate)
        ctx.save_for_backward(g3)
        return K.grid_push(None, g3, shape3, bnd, order, extrapolate=extrapolate)

    @staticmethod
    def backward(ctx, grad):
        (g3,) = ctx.saved_tensors
        bnd, order, extrapolate = ctx.opt
        gg = None
        if ctx.needs_input_grad[0]:
            gg = K.grid_grad(grad.contiguous(), g3, bnd, order, extrapolate=extrapolate)[:, 0].to(g3.dtype)
        return gg, None, None, None, None


def _wants_grad(*tensors) -> bool:
    return torch.is_grad_enabled() and any(t is not None and t.requires_grad for t in tensors)


def grid_pull(input: torch.Tensor, grid: torch.Tensor, interpolation="linear", bound="zero", extrapolate: bool = True) -> torch.Tensor:  # noqa: A002
    """Sample `input` (B, C, Wi[, Hi[, Di]]) at the voxel coordinates `grid` (B, Wo[, Ho[, Do]], 1|2|3).

    interpolation: 0-7 or 'nearest' | 'linear' | 'quadratic' | 'cubic' | 'fourth' | 'fifth' | 'sixth' | 'seventh' (or a list, one per
    dimension); bound: 0 'replicate'/'nearest'/'border', 1 'dct1'/'mirror', 2 'dct2'/'reflect', 3 'dst1'/'antimirror',
    4 'dst2'/'antireflect', 5 'dft'/'wrap', 7 'zero'/'zeros' (or a list); extrapolate=False zeroes samples outside the field of view.
    'sliding' (flow fields only) is not implemented.  Differentiable in `input` and `grid` (the reference's _GridPull)."""
    if not input.is_cuda:
        raise RuntimeError("monai_b200.grid_pull runs on CUDA tensors only (there is no CPU fallback)")
    nd = grid.shape[-1]
    if nd not in (1, 2, 3) or input.dim() != nd + 2 or grid.dim() != nd + 2:
        raise ValueError(f"grid_pull expects input (B, C, *{nd} spatial) and grid (B, *{nd} spatial, {nd}); got {tuple(input.shape)}, {tuple(grid.shape)}")
    order = _codes(interpolation, K.ORDERS, "interpolation", nd)
    if nd == 3:
        # drop-in fidelity: the reference's implementation object copies interpolation1 into interpolation2
        # (monai/csrc/resample/pushpull_cpu.cpp:499-501, same in pushpull_cuda.cu), so its third axis runs with the second
        # axis' order.  b200_grid_pull itself is per-axis; the quirk lives here, where the reference's callers would see it.
        order = [order[0], order[1], order[1]]
    bnd = _codes(bound, K.BOUNDS, "bound", nd)
    if any(b == 6 for b in bnd):
        raise NotImplementedError("bound 'sliding' (deformation fields) is not implemented")
    like = input
    x = input.as_subclass(torch.Tensor) if type(input) is not torch.Tensor else input
    g = grid.as_subclass(torch.Tensor) if type(grid) is not torch.Tensor else grid
    lift = 3 - nd
    x3 = x.reshape(*x.shape, *([1] * lift))                     # trailing singleton axes: coordinate 0 there
    if g.dtype not in (torch.float32, torch.float64):
        g = g.float()
    g3 = g.reshape(g.shape[0], *g.shape[1:-1], *([1] * lift), nd)
    if lift:
        g3 = torch.cat([g3, torch.zeros((*g3.shape[:-1], lift), dtype=g3.dtype, device=g3.device)], dim=-1)
    if _wants_grad(x3, g3):   # x3 / g3 are differentiable views of the caller's tensors: the gradients flow back through them
        out = _GridPull.apply(x3.contiguous(), g3.contiguous(), bnd + [0] * lift, order + [0] * lift, extrapolate)
    else:
        out = K.grid_pull(x3.detach(), g3.detach(), bnd + [0] * lift, order + [0] * lift, extrapolate=extrapolate, channel_last=True)
    out = out.reshape(x.shape[0], x.shape[1], *g.shape[1:-1])
    if type(like) is not torch.Tensor and hasattr(like, "copy_meta_from"):
        wrapped = type(like)(out)
        wrapped.copy_meta_from(like, copy_attr=False)
        return wrapped
    return out


def _push_args(grid: torch.Tensor, interpolation, bound, what: str):
    if not grid.is_cuda:
        raise RuntimeError(f"monai_b200.{what} runs on CUDA tensors only (there is no CPU fallback)")
    nd = grid.shape[-1]
    if nd not in (1, 2, 3) or grid.dim() != nd + 2:
        raise ValueError(f"{what} expects a grid (B, *{nd} spatial, {nd}); got {tuple(grid.shape)}")
    order = _codes(interpolation, K.ORDERS, "interpolation", nd)
    if nd == 3:
        order = [order[0], order[1], order[1]]      # the reference's interpolation2 = interpolation1 quirk, as in grid_pull
    bnd = _codes(bound, K.BOUNDS, "bound", nd)
    if any(b == 6 for b in bnd):
        raise NotImplementedError("bound 'sliding' (deformation fields) is not implemented")
    g = grid.as_subclass(torch.Tensor) if type(grid) is not torch.Tensor else grid
    if g.dtype not in (torch.float32, torch.float64):
        g = g.float()
    lift = 3 - nd
    g3 = g.reshape(g.shape[0], *g.shape[1:-1], *([1] * lift), nd)
    if lift:
        g3 = torch.cat([g3, torch.zeros((*g3.shape[:-1], lift), dtype=g3.dtype, device=g3.device)], dim=-1)
    return nd, lift, order + [0] * lift, bnd + [0] * lift, g3


def grid_push(input: torch.Tensor, grid: torch.Tensor, shape=None, interpolation="linear", bound="zero", extrapolate: bool = True) -> torch.Tensor:  # noqa: A002
    """Splat `input` (B, C, Wi[, Hi[, Di]]) at the voxel coordinates `grid` (B, Wi[, Hi[, Di]], 1|2|3) into a volume of spatial
    `shape` (default: the input's): the adjoint of grid_pull (monai/networks/layers/spatial_transforms.py:160-235 -> monai._C.grid_push).
    Same `interpolation` / `bound` / `extrapolate` vocabulary as grid_pull.  Float32 accumulation; differentiable in `input` and `grid`."""
    nd, lift, order, bnd, g3 = _push_args(grid, interpolation, bound, "grid_push")
    if input.dim() != nd + 2:
        raise ValueError(f"grid_push expects input (B, C, *{nd} spatial); got {tuple(input.shape)}")
    x = input.as_subclass(torch.Tensor) if type(input) is not torch.Tensor else input
    if shape is None:
        shape = tuple(x.shape[2:])
    shape = [int(v) for v in shape]
    x3 = x.reshape(*x.shape, *([1] * lift))
    if _wants_grad(x3, g3):
        out = _GridPush.apply(x3.contiguous(), g3.contiguous(), shape + [1] * lift, bnd, order, extrapolate)
    else:
        out = K.grid_push(x3.detach(), g3.detach(), shape + [1] * lift, bnd, order, extrapolate=extrapolate)
    out = out.reshape(x.shape[0], x.shape[1], *shape).to(x.dtype if x.dtype.is_floating_point else torch.float32)
    if type(input) is not torch.Tensor and hasattr(input, "copy_meta_from"):
        wrapped = type(input)(out)
        wrapped.copy_meta_from(input, copy_attr=False)
        return wrapped
    return out


def grid_count(grid: torch.Tensor, shape=None, interpolation="linear", bound="zero", extrapolate: bool = True) -> torch.Tensor:
    """grid_push of an image of ones: how much every voxel of the `shape` volume receives (B, 1, *shape)
    (monai/networks/layers/spatial_transforms.py:262-311 -> monai._C.grid_count).  `shape` defaults to the grid's spatial shape.
    Differentiable in `grid`."""
    nd, lift, order, bnd, g3 = _push_args(grid, interpolation, bound, "grid_count")
    if shape is None:
        shape = tuple(grid.shape[1:-1])
    shape = [int(v) for v in shape]
    if _wants_grad(g3):
        out = _GridCount.apply(g3.contiguous(), shape + [1] * lift, bnd, order, extrapolate)
    else:
        out = K.grid_push(None, g3.detach(), shape + [1] * lift, bnd, order, extrapolate=extrapolate)
    out = out.reshape(grid.shape[0], 1, *shape).to(grid.dtype if grid.dtype.is_floating_point else torch.float32)
    if type(grid) is not torch.Tensor and hasattr(grid, "copy_meta_from"):
        wrapped = type(grid)(out)
        wrapped.copy_meta_from(grid, copy_attr=False)
        return wrapped
    return out


def grid_grad(input: torch.Tensor, grid: torch.Tensor, interpolation="linear", bound="zero", extrapolate: bool = True) -> torch.Tensor:  # noqa: A002
    """Spatial gradients of `input` (B, C, Wi[, Hi[, Di]]) sampled at the voxel coordinates `grid` (B, Wo[, Ho[, Do]], 1|2|3):
    (B, C, Wo[, Ho[, Do]], 1|2|3) (monai/networks/layers/spatial_transforms.py:345-408 -> monai._C.grid_grad).  Forward only."""
    nd, lift, order, bnd, g3 = _push_args(grid, interpolation, bound, "grid_grad")
    if input.dim() != nd + 2:
        raise ValueError(f"grid_grad expects input (B, C, *{nd} spatial); got {tuple(input.shape)}")
    x = input.as_subclass(torch.Tensor) if type(input) is not torch.Tensor else input
    if _wants_grad(x, g3):
        raise NotImplementedError("monai_b200.grid_grad is forward only (its backward needs the spline Hessians); call it under torch.no_grad() "
                                  "or on tensors that do not require grad")
    out = K.grid_grad(x.detach().reshape(*x.shape, *([1] * lift)), g3.detach(), bnd, order, extrapolate=extrapolate)
    out = out.reshape(x.shape[0], x.shape[1], *grid.shape[1:-1], 3)[..., :nd].to(x.dtype if x.dtype.is_floating_point else torch.float32)
    if type(input) is not torch.Tensor and hasattr(input, "copy_meta_from"):
        wrapped = type(input)(out)
        wrapped.copy_meta_from(input, copy_attr=False)
        return wrapped
    return out


class AffineTransform(torch.nn.Module):
    """Apply a batch of affine matrices to a batch of images (monai/networks/layers/spatial_transforms.py:439-592): the same constructor,
    `forward(src, theta, spatial_size=None)`, argument checks and conventions (`normalized`, `reverse_indexing`, `zero_centered`,
    `align_corners`, bilinear / nearest, zeros / border / reflection).  The reference builds a dense F.affine_grid and calls
    F.grid_sample; here the whole chain is folded into ONE output-index -> source-index matrix per batch item on the host
    (monai_b200.transforms.utils.affine_transform_matrix, float64) and b200_resample_affine evaluates it per voxel, no grid in memory.
    Forward only: `theta` is read on the host (one small device-to-host copy when it lives on the GPU)."""

    def __init__(self, spatial_size=None, normalized: bool = False, mode: str = "bilinear", padding_mode: str = "zeros",
                 align_corners: bool = True, reverse_indexing: bool = True, zero_centered: bool | None = None) -> None:
        super().__init__()
        if spatial_size is not None and not isinstance(spatial_size, (list, tuple)):
            spatial_size = (spatial_size,)
        self.spatial_size = tuple(int(s) for s in spatial_size) if spatial_size is not None else None
        self.normalized = normalized
        mode = str(getattr(mode, "value", mode)).lower()
        padding_mode = str(getattr(padding_mode, "value", padding_mode)).lower()
        if mode not in ("bilinear", "nearest", "bicubic"):
            raise ValueError(f"unsupported mode {mode!r}; options: bilinear, nearest, bicubic")
        if mode == "bicubic":
            raise NotImplementedError("monai_b200 AffineTransform implements bilinear and nearest interpolation")
        if padding_mode not in ("zeros", "border", "reflection"):
            raise ValueError(f"unsupported padding_mode {padding_mode!r}; options: zeros, border, reflection")
        self.mode, self.padding_mode = mode, padding_mode
        self.align_corners = align_corners
        self.reverse_indexing = reverse_indexing
        if zero_centered is not None and self.normalized:
            raise ValueError("`normalized=True` is not compatible with the `zero_centered` option.")
        self.zero_centered = zero_centered if zero_centered is not None else False

    def forward(self, src: torch.Tensor, theta: torch.Tensor, spatial_size=None) -> torch.Tensor:
        from ...transforms import utils as U
        from ...transforms.spatial import _resample

        if not isinstance(theta, torch.Tensor):
            raise TypeError(f"theta must be torch.Tensor but is {type(theta).__name__}.")
        if theta.dim() not in (2, 3):
            raise ValueError(f"theta must be Nxdxd or dxd, got {theta.shape}.")
        if theta.dim() == 2:
            theta = theta[None]
        tshape = tuple(theta.shape[1:])
        if tshape not in ((2, 3), (3, 4), (3, 3), (4, 4)):
            raise ValueError(f"theta must be Nx3x3 or Nx4x4, got {theta.shape}.")
        if not torch.is_floating_point(theta):
            raise ValueError(f"theta must be floating point data, got {theta.dtype}")
        if not isinstance(src, torch.Tensor):
            raise TypeError(f"src must be torch.Tensor but is {type(src).__name__}.")
        sr = src.dim() - 2
        if sr not in (2, 3):
            raise ValueError(f"Unsupported src dimension: {sr}, available options are [2, 3].")
        if tshape[1] != sr + 1:
            raise ValueError(f"theta {tuple(theta.shape)} does not match the {sr} spatial dims of src {tuple(src.shape)}")
        if not torch.is_floating_point(src):
            raise RuntimeError(f"src must be floating point data (as F.grid_sample requires), got {src.dtype}")
        if theta.dtype != src.dtype:
            raise RuntimeError(f"src and theta must share a dtype (as F.grid_sample requires), got {src.dtype} and {theta.dtype}")
        if not src.is_cuda:
            raise RuntimeError("monai_b200.AffineTransform runs on CUDA tensors only (there is no CPU fallback)")
        dst_spatial = tuple(src.shape[2:])
        if self.spatial_size is not None:
            dst_spatial = self.spatial_size
        if spatial_size is not None:
            dst_spatial = tuple(int(s) for s in (spatial_size if isinstance(spatial_size, (list, tuple)) else (spatial_size,)))
        if len(dst_spatial) != sr:
            raise ValueError(f"spatial_size {dst_spatial} does not match the {sr} spatial dims of src")
        th = theta.detach().to(device="cpu", dtype=torch.float64).numpy()
        if tshape[0] == sr:      # pad to homogeneous form
            bottom = np.zeros((th.shape[0], 1, sr + 1))
            bottom[:, 0, -1] = 1.0
            th = np.concatenate([th, bottom], axis=1)
        if th.shape[0] == 1 and src.shape[0] > 1:
            th = np.repeat(th, src.shape[0], axis=0)
        if th.shape[0] != src.shape[0]:
            raise ValueError(f"affine and image batch dimension must match, got affine={th.shape[0]} image={src.shape[0]}.")
        x = src.as_subclass(torch.Tensor) if type(src) is not torch.Tensor else src
        out = []
        for b in range(x.shape[0]):
            m = U.affine_transform_matrix(th[b], tuple(x.shape[2:]), dst_spatial, self.normalized, self.reverse_indexing,
                                          bool(self.align_corners), self.zero_centered)
            out.append(_resample(x[b].detach(), m, sr, dst_spatial, self.mode, self.padding_mode, bool(self.align_corners)))
        return torch.stack(out, dim=0).to(src.dtype)
