"""DynUNet building blocks (monai/networks/blocks/dynunet_block.py:25-327) on CUDA kernels.

Same module tree, parameter names and defaults as the reference (`load_state_dict(reference.state_dict())` works); the arithmetic
is `b200_conv3d_direct` (any kernel / stride per axis, transposed included; fp32 accumulation) followed by ONE
`b200_instnorm_stats` + `b200_norm_act` pass per normalisation (affine instance norm, LeakyReLU and the residual add fused).
SwinUNETR keeps its own tensor-core specialisation of `UnetResBlock` (kernel 3, stride 1) in nets/swin_unetr.py.
"""
from __future__ import annotations

from typing import Sequence

import numpy as np
import torch
import torch.nn as nn

from ... import _kernels as K
from ..layers.factories import get_act_layer, get_norm_layer
from .acti_norm import norm_act_from_modules
from .convolutions import Convolution

__all__ = ["UnetBasicBlock", "UnetResBlock", "UnetUpBlock", "UnetOutBlock", "get_conv_layer", "get_padding", "get_output_padding"]

_LRELU = ("leakyrelu", {"inplace": True, "negative_slope": 0.01})


def get_padding(kernel_size, stride):
    """dynunet_block.py:304-312: (k - s + 1) / 2 per axis, truncated; negative values are an error."""
    p = (np.atleast_1d(kernel_size) - np.atleast_1d(stride) + 1) / 2
    if p.min() < 0:
        raise AssertionError("padding value should not be negative, please change the kernel size and/or stride.")
    out = tuple(int(v) for v in p)
    return out if len(out) > 1 else out[0]


def get_output_padding(kernel_size, stride, padding):
    """dynunet_block.py:315-327: 2p + s - k per axis."""
    op = 2 * np.atleast_1d(padding) + np.atleast_1d(stride) - np.atleast_1d(kernel_size)
    if op.min() < 0:
        raise AssertionError("out_padding value should not be negative, please change the kernel size and/or stride.")
    out = tuple(int(v) for v in op)
    return out if len(out) > 1 else out[0]


def get_conv_layer(spatial_dims: int, in_channels: int, out_channels: int, kernel_size=3, stride=1, act="PRELU", norm="INSTANCE",
                   dropout=None, bias: bool = False, conv_only: bool = True, is_transposed: bool = False) -> Convolution:
    """dynunet_block.py:270-301."""
    padding = get_padding(kernel_size, stride)
    output_padding = get_output_padding(kernel_size, stride, padding) if is_transposed else None
    return Convolution(spatial_dims, in_channels, out_channels, strides=stride, kernel_size=kernel_size, act=act, norm=norm, dropout=dropout,
                       bias=bias, conv_only=conv_only, is_transposed=is_transposed, padding=padding, output_padding=output_padding)


def _plain(spatial_dims, cin, cout, kernel_size, stride, dropout, **kw) -> Convolution:
    return get_conv_layer(spatial_dims, cin, cout, kernel_size=kernel_size, stride=stride, dropout=dropout, act=None, norm=None, conv_only=False, **kw)


class UnetBasicBlock(nn.Module):
    """conv - norm - lrelu - conv - norm - lrelu (dynunet_block.py:114-177)."""

    def __init__(self, spatial_dims: int, in_channels: int, out_channels: int, kernel_size, stride, norm_name, act_name=_LRELU, dropout=None):
        super().__init__()
        self.conv1 = _plain(spatial_dims, in_channels, out_channels, kernel_size, stride, dropout)
        self.conv2 = _plain(spatial_dims, out_channels, out_channels, kernel_size, 1, dropout)
        self.lrelu = get_act_layer(name=act_name)
        self.norm1 = get_norm_layer(name=norm_name, spatial_dims=spatial_dims, channels=out_channels)
        self.norm2 = get_norm_layer(name=norm_name, spatial_dims=spatial_dims, channels=out_channels)

    def forward(self, inp: torch.Tensor) -> torch.Tensor:
        out = norm_act_from_modules(self.conv1(inp), self.norm1, self.lrelu)
        return norm_act_from_modules(self.conv2(out), self.norm2, self.lrelu)


class UnetResBlock(nn.Module):
    """Residual form (dynunet_block.py:25-111): a 1x1 (strided) convolution + norm on the skip when shape or channels change."""

    def __init__(self, spatial_dims: int, in_channels: int, out_channels: int, kernel_size, stride, norm_name, act_name=_LRELU, dropout=None):
        super().__init__()
        self.conv1 = _plain(spatial_dims, in_channels, out_channels, kernel_size, stride, dropout)
        self.conv2 = _plain(spatial_dims, out_channels, out_channels, kernel_size, 1, dropout)
        self.lrelu = get_act_layer(name=act_name)
        self.norm1 = get_norm_layer(name=norm_name, spatial_dims=spatial_dims, channels=out_channels)
        self.norm2 = get_norm_layer(name=norm_name, spatial_dims=spatial_dims, channels=out_channels)
        self.downsample = in_channels != out_channels
        if not np.all(np.atleast_1d(stride) == 1):
            self.downsample = True
        if self.downsample:
            self.conv3 = _plain(spatial_dims, in_channels, out_channels, 1, stride, dropout)
            self.norm3 = get_norm_layer(name=norm_name, spatial_dims=spatial_dims, channels=out_channels)

    def forward(self, inp: torch.Tensor) -> torch.Tensor:
        out = norm_act_from_modules(self.conv1(inp), self.norm1, self.lrelu)
        out = self.conv2(out)
        residual = inp
        if hasattr(self, "conv3"):
            residual = norm_act_from_modules(self.conv3(inp), self.norm3, None)   # (affine) norm3 of the skip branch
        return norm_act_from_modules(out, self.norm2, self.lrelu, res=residual)     # lrelu(norm2(out) + residual)


class UnetUpBlock(nn.Module):
    """Transposed convolution (kernel = stride = upsample_kernel_size), concat with the skip, UnetBasicBlock (dynunet_block.py:180-244)."""

    def __init__(self, spatial_dims: int, in_channels: int, out_channels: int, kernel_size, stride, upsample_kernel_size, norm_name,
                 act_name=_LRELU, dropout=None, trans_bias: bool = False):
        super().__init__()
        self.transp_conv = _plain(spatial_dims, in_channels, out_channels, upsample_kernel_size, upsample_kernel_size, dropout, bias=trans_bias,
                                  is_transposed=True)
        self.conv_block = UnetBasicBlock(spatial_dims, out_channels + out_channels, out_channels, kernel_size=kernel_size, stride=1,
                                         dropout=dropout, norm_name=norm_name, act_name=act_name)

    def forward(self, inp: torch.Tensor, skip: torch.Tensor) -> torch.Tensor:
        return self.conv_block(K.cat_channels([self.transp_conv(inp), skip]))


class UnetOutBlock(nn.Module):
    """1x1 convolution with bias to the class channels (dynunet_block.py:247-267)."""

    def __init__(self, spatial_dims: int, in_channels: int, out_channels: int, dropout=None):
        super().__init__()
        self.conv = _plain(spatial_dims, in_channels, out_channels, 1, 1, dropout, bias=True)

    def forward(self, inp: torch.Tensor) -> torch.Tensor:
        return self.conv(inp)
