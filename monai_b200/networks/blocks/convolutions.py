"""`Convolution` and `ResidualUnit` (monai/networks/blocks/convolutions.py:25-171, 174-318) on CUDA kernels.

`self.conv` is an nn.Conv{1,2,3}d / nn.ConvTranspose{1,2,3}d used ONLY as a parameter container (same names, shapes
and default initialisation as the reference, so `load_state_dict(reference.state_dict())` works).  The arithmetic is
`b200_conv3d_direct` (fp32 accumulate, exact-parity path) followed by the fused ADN kernels.  1-D / 2-D convolutions
run through the same 3-D kernel with leading singleton axes.
"""
from __future__ import annotations

from typing import Sequence

import numpy as np
import torch
import torch.nn as nn

from ... import _kernels as K
from ..layers.convutils import same_padding, stride_minus_kernel_padding
from .acti_norm import ADN, norm_act_from_modules

__all__ = ["Convolution", "ResidualUnit", "run_conv_module"]

_CONV = {1: nn.Conv1d, 2: nn.Conv2d, 3: nn.Conv3d}
_CONVT = {1: nn.ConvTranspose1d, 2: nn.ConvTranspose2d, 3: nn.ConvTranspose3d}


def _lift3(v, fill) -> tuple[int, int, int]:
    v = tuple(int(i) for i in v)
    return (fill,) * (3 - len(v)) + v  # type: ignore[return-value]


def run_conv_module(conv: nn.Module, x: torch.Tensor, out: torch.Tensor | None = None) -> torch.Tensor:
    """Execute a torch conv *container* (Conv / ConvTranspose, groups=1, dilation=1, zero padding) with the CUDA kernel."""
    transposed = isinstance(conv, (nn.ConvTranspose1d, nn.ConvTranspose2d, nn.ConvTranspose3d))
    if conv.groups != 1 or any(d != 1 for d in conv.dilation):
        raise NotImplementedError("monai_b200 convolutions support groups=1, dilation=1")
    if getattr(conv, "padding_mode", "zeros") != "zeros":
        raise NotImplementedError("monai_b200 convolutions support zero padding only")
    nd = x.dim() - 2
    lift = 3 - nd
    x3 = x.reshape(x.shape[0], x.shape[1], *([1] * lift), *x.shape[2:]) if lift else x
    w = conv.weight
    w3 = w.reshape(w.shape[0], w.shape[1], *([1] * lift), *w.shape[2:]) if lift else w
    out3 = None
    if out is not None:
        out3 = out.reshape(out.shape[0], out.shape[1], *([1] * lift), *out.shape[2:]) if lift else out
    y3 = K.conv3d_direct(
        x3, w3, conv.bias, stride=_lift3(conv.stride, 1), padding=_lift3(conv.padding, 0), transposed=transposed,
        output_padding=_lift3(conv.output_padding, 0) if transposed else 0, out=out3,
    )
    return y3.reshape(y3.shape[0], y3.shape[1], *y3.shape[2 + lift :]) if lift else y3


class Convolution(nn.Sequential):
    def __init__(
        self,
        spatial_dims: int,
        in_channels: int,
        out_channels: int,
        strides: Sequence[int] | int = 1,
        kernel_size: Sequence[int] | int = 3,
        adn_ordering: str = "NDA",
        act="PRELU",
        norm="INSTANCE",
        dropout=None,
        dropout_dim: int | None = 1,
        dilation: Sequence[int] | int = 1,
        groups: int = 1,
        bias: bool = True,
        conv_only: bool = False,
        is_transposed: bool = False,
        padding: Sequence[int] | int | None = None,
        output_padding: Sequence[int] | int | None = None,
    ) -> None:
        super().__init__()
        self.spatial_dims = spatial_dims
        self.in_channels = in_channels
        self.out_channels = out_channels
        self.is_transposed = is_transposed
        if padding is None:
            padding = same_padding(kernel_size, dilation)
        if is_transposed:
            if output_padding is None:
                output_padding = stride_minus_kernel_padding(1, strides)
            conv = _CONVT[spatial_dims](
                in_channels, out_channels, kernel_size=kernel_size, stride=strides, padding=padding,
                output_padding=output_padding, groups=groups, bias=bias, dilation=dilation,
            )
        else:
            conv = _CONV[spatial_dims](
                in_channels, out_channels, kernel_size=kernel_size, stride=strides, padding=padding, dilation=dilation,
                groups=groups, bias=bias,
            )
        self.add_module("conv", conv)
        if conv_only or (act is None and norm is None and dropout is None):
            return
        self.add_module(
            "adn",
            ADN(ordering=adn_ordering, in_channels=out_channels, act=act, norm=norm, norm_dim=spatial_dims, dropout=dropout, dropout_dim=dropout_dim),
        )

    def forward(self, x: torch.Tensor, out: torch.Tensor | None = None) -> torch.Tensor:  # type: ignore[override]
        has_adn = hasattr(self, "adn")
        y = run_conv_module(self.conv, x, out=None if has_adn else out)
        if has_adn:
            y = self.adn(y)
            if out is not None:
                out.copy_(y)
                y = out
        return y


class ResidualUnit(nn.Module):
    def __init__(
        self,
        spatial_dims: int,
        in_channels: int,
        out_channels: int,
        strides: Sequence[int] | int = 1,
        kernel_size: Sequence[int] | int = 3,
        subunits: int = 2,
        adn_ordering: str = "NDA",
        act="PRELU",
        norm="INSTANCE",
        dropout=None,
        dropout_dim: int | None = 1,
        dilation: Sequence[int] | int = 1,
        bias: bool = True,
        last_conv_only: bool = False,
        padding: Sequence[int] | int | None = None,
    ) -> None:
        super().__init__()
        self.spatial_dims = spatial_dims
        self.in_channels = in_channels
        self.out_channels = out_channels
        self.conv = nn.Sequential()
        self.residual: nn.Module = nn.Identity()
        if not padding:
            padding = same_padding(kernel_size, dilation)
        cin, stride = in_channels, strides
        for su in range(max(1, subunits)):
            self.conv.add_module(
                f"unit{su:d}",
                Convolution(
                    spatial_dims, cin, out_channels, strides=stride, kernel_size=kernel_size, adn_ordering=adn_ordering,
                    act=act, norm=norm, dropout=dropout, dropout_dim=dropout_dim, dilation=dilation, bias=bias,
                    conv_only=last_conv_only and su == (max(1, subunits) - 1), padding=padding,
                ),
            )
            cin, stride = out_channels, 1  # later sub-units keep channels and resolution
        if np.prod(strides) != 1 or in_channels != out_channels:
            rk, rp = kernel_size, padding
            if np.prod(strides) == 1:  # channel adaptation only: 1x1 kernel, no padding
                rk, rp = 1, 0
            self.residual = _CONV[spatial_dims](in_channels, out_channels, rk, strides, rp, bias=bias)

    def forward(self, x: torch.Tensor) -> torch.Tensor:
        res = x if isinstance(self.residual, nn.Identity) else run_conv_module(self.residual, x)
        cx = x
        for unit in self.conv:
            cx = unit(cx)
        return norm_act_from_modules(cx, None, None, res=res)  # cx + res
