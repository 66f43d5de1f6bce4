"""Host-side window planner and importance-map factors.

Pure integer / tiny-vector arithmetic that must be bit-exact with the reference:
  * `get_valid_patch_size`  -- monai/data/utils.py:343-354
  * `dense_patch_slices`    -- monai/data/utils.py:166-206
  * `compute_importance_map`-- monai/data/utils.py:1084-1134
The reference materialises the N-D importance map; the CUDA blend only needs its separable factors, so
`importance_factors` returns the per-axis vectors and the clamp value (exactly what the dense map would hold).
"""
from __future__ import annotations

import math
from typing import Sequence

import numpy as np
import torch

__all__ = [
    "get_valid_patch_size",
    "dense_patch_starts",
    "dense_patch_slices",
    "compute_importance_map",
    "importance_factors",
]


def _ensure_tuple_size(vals, dim: int, pad_val=0) -> tuple:
    if isinstance(vals, (int, float)) or vals is None:
        vals = (vals,)
    vals = tuple(vals) + (pad_val,) * dim
    return vals[:dim]


def get_valid_patch_size(image_size: Sequence[int], patch_size) -> tuple[int, ...]:
    """Patch size clipped to the image; 0/None entries (or missing trailing entries) take the image dimension."""
    ndim = len(image_size)
    if isinstance(patch_size, np.ndarray):
        patch_size = patch_size.tolist()
    # NB: as in the reference *code* (not its docstring) a scalar fills only the first axis; the rest fall back
    patch = _ensure_tuple_size(patch_size, ndim)
    return tuple(min(int(ms), int(ps) if ps else int(ms)) for ms, ps in zip(image_size, patch))


def dense_patch_starts(image_size: Sequence[int], patch_size: Sequence[int], scan_interval: Sequence[int]) -> list[list[int]]:
    """Per-axis window start positions; the window set is their Cartesian product in "ij" order."""
    ndim = len(image_size)
    patch_size = get_valid_patch_size(image_size, patch_size)
    scan_interval = _ensure_tuple_size(scan_interval, ndim)
    starts: list[list[int]] = []
    for i in range(ndim):
        if scan_interval[i] == 0:
            num = 1
        else:
            upper = int(math.ceil(float(image_size[i]) / scan_interval[i]))
            hit = next((d for d in range(upper) if d * scan_interval[i] + patch_size[i] >= image_size[i]), None)
            num = hit + 1 if hit is not None else 1
        axis = []
        for idx in range(num):
            s = idx * scan_interval[i]
            s -= max(s + patch_size[i] - image_size[i], 0)  # the last window is snapped back inside the image
            axis.append(int(s))
        starts.append(axis)
    return starts


def dense_patch_slices(image_size, patch_size, scan_interval, return_slice: bool = True):
    """All N-D windows of `patch_size` over `image_size` (same ordering as the reference: first axis slowest)."""
    patch_size = get_valid_patch_size(image_size, patch_size)
    starts = dense_patch_starts(image_size, patch_size, scan_interval)
    grid = np.asarray([g.flatten() for g in np.meshgrid(*starts, indexing="ij")]).T
    if return_slice:
        return [tuple(slice(int(s), int(s) + patch_size[d]) for d, s in enumerate(row)) for row in grid]
    return [tuple((int(s), int(s) + patch_size[d]) for d, s in enumerate(row)) for row in grid]


def _rep(v, n):
    if isinstance(v, (list, tuple)):
        if len(v) != n:
            raise ValueError(f"Sequence must have length {n}, got {len(v)}.")
        return tuple(v)
    return (v,) * n


def importance_factors(patch_size: Sequence[int], mode="constant", sigma_scale=0.125) -> tuple[list[torch.Tensor], float]:
    """Separable fp32 factors g_i and the clamp value such that
    compute_importance_map(...)[i,j,k] == max((g_0[i]*g_1[j])*g_2[k], clamp)   (fp32, same multiplication order)."""
    mode = str(getattr(mode, "value", mode)).lower()
    if mode == "constant":
        return [torch.ones(int(p), dtype=torch.float32) for p in patch_size], 1.0
    if mode != "gaussian":
        raise ValueError(f"Unsupported mode: {mode}, available options are [constant, gaussian].")
    sig = _rep(sigma_scale, len(patch_size))
    factors = []
    for p, s in zip(patch_size, sig):
        sigma = p * s
        x = torch.arange(start=-(p - 1) / 2.0, end=(p - 1) / 2.0 + 1, dtype=torch.float)
        factors.append(torch.exp(x**2 / (-2 * sigma**2)))
    # min of an outer product of positive vectors, evaluated in the reference's fp32 multiplication order
    m = factors[0].min()
    for f in factors[1:]:
        m = m * f.min()
    return factors, max(float(m.item()), 1e-3)


def compute_importance_map(patch_size, mode="constant", sigma_scale=0.125, device="cpu", dtype=torch.float32) -> torch.Tensor:
    """Dense importance map with the reference's values (used for API parity: `roi_weight_map`, `process_fn`)."""
    factors, clamp = importance_factors(tuple(patch_size), mode, sigma_scale)
    m = factors[0]
    for i, f in enumerate(factors[1:], start=1):
        m = m.unsqueeze(-1) * f[(None,) * i]
    m = torch.clamp_(m.to(torch.float), min=clamp)
    return m.to(device=torch.device(device), dtype=dtype)
