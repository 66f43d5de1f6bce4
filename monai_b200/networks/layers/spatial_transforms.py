"""`grid_pull` with the reference's call contract (monai/networks/layers/spatial_transforms.py:35-132), backed by the CUDA
kernel b200_grid_pull (no monai._C needed): spline orders 0-7, the seven boundary conditions of the discrete transforms,
per-axis settings given in the reference's [W, H, D] order (= the order of the grid's last dimension).  Inference only."""
from __future__ import annotations

from typing import Sequence

import torch

from ... import _kernels as K

__all__ = ["grid_pull"]


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


def grid_pull(input: torch.Tensor, grid: torch.Tensor, interpolation="linear", bound="zero", extrapolate: bool = True) -> torch.Tensor:  # noqa: A002
    """Sample `input` (B, C, Wi[, Hi[, Di]]) at the voxel coordinates `grid` (B, Wo[, Ho[, Do]], 1|2|3).

    interpolation: 0-7 or 'nearest' | 'linear' | 'quadratic' | 'cubic' | 'fourth' | 'fifth' | 'sixth' | 'seventh' (or a list, one per
    dimension); bound: 0 'replicate'/'nearest'/'border', 1 'dct1'/'mirror', 2 'dct2'/'reflect', 3 'dst1'/'antimirror',
    4 'dst2'/'antireflect', 5 'dft'/'wrap', 7 'zero'/'zeros' (or a list); extrapolate=False zeroes samples outside the field of view.
    'sliding' (flow fields only) is not implemented."""
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
    out = K.grid_pull(x3.detach(), g3.detach(), bnd + [0] * lift, order + [0] * lift, extrapolate=extrapolate, channel_last=True)
    out = out.reshape(x.shape[0], x.shape[1], *g.shape[1:-1])
    if type(like) is not torch.Tensor and hasattr(like, "copy_meta_from"):
        wrapped = type(like)(out)
        wrapped.copy_meta_from(like, copy_attr=False)
        return wrapped
    return out
