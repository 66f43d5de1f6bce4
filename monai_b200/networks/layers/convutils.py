"""Padding arithmetic and 1-D Gaussian taps (monai/networks/layers/convutils.py:22-53, 78-131).  Host-side, exact."""
from __future__ import annotations

from typing import Sequence

import numpy as np
import torch

__all__ = ["same_padding", "stride_minus_kernel_padding", "gaussian_1d"]


def _squeeze(vals):
    vals = tuple(int(v) for v in vals)
    return vals if len(vals) > 1 else vals[0]


def same_padding(kernel_size: Sequence[int] | int, dilation: Sequence[int] | int = 1):
    """Padding that keeps the shape at stride 1: (k - 1) / 2 * dilation; odd (k-1)*dilation is rejected."""
    k = np.atleast_1d(kernel_size)
    dil = np.atleast_1d(dilation)
    if np.any((k - 1) * dilation % 2 == 1):
        raise NotImplementedError(f"Same padding not available for kernel_size={k} and dilation={dil}.")
    return _squeeze((k - 1) / 2 * dil)


def stride_minus_kernel_padding(kernel_size: Sequence[int] | int, stride: Sequence[int] | int):
    return _squeeze(np.atleast_1d(stride) - np.atleast_1d(kernel_size))


def gaussian_1d(sigma, truncated: float = 4.0, approx: str = "erf", normalize: bool = False) -> torch.Tensor:
    """Discrete 1-D Gaussian taps.  "erf": difference of the error function over each unit cell, clamped at 0 and
    NOT normalised by default; tail = int(max(sigma*truncated, 0.5) + 0.5) taps on each side."""
    sigma = torch.as_tensor(sigma, dtype=torch.float, device=sigma.device if isinstance(sigma, torch.Tensor) else None)
    if truncated <= 0.0:
        raise ValueError(f"truncated must be positive, got {truncated}.")
    tail = int(max(float(sigma) * truncated, 0.5) + 0.5)
    x = torch.arange(-tail, tail + 1, dtype=torch.float, device=sigma.device)
    kind = approx.lower()
    if kind == "erf":
        t = 0.70710678 / torch.abs(sigma)
        out = (0.5 * ((t * (x + 0.5)).erf() - (t * (x - 0.5)).erf())).clamp(min=0)
    elif kind == "sampled":
        out = torch.exp(-0.5 / (sigma * sigma) * x**2)
        if not normalize:
            out = out / (2.5066282 * sigma)
    else:
        raise NotImplementedError(f"Unsupported option: approx='{approx}'.")
    return out / out.sum() if normalize else out
