"""`UNet` (monai/networks/nets/unet.py:27-301): recursive Sequential(down, SkipConnection(sub), up).

Constructor arguments, module tree and state_dict keys are those of the reference (`model.0.conv.weight`,
`model.1.submodule...`, `model.2.conv.weight`, `...adn.A.weight`), so reference checkpoints load unchanged.
Every layer executes on the monai_b200 CUDA kernels; activations follow the input dtype (fp16 or fp32) and all
accumulation is fp32.
"""
from __future__ import annotations

import warnings
from collections.abc import Sequence

import torch
import torch.nn as nn

from ..blocks.convolutions import Convolution, ResidualUnit
from ..layers.simplelayers import SkipConnection

__all__ = ["UNet", "Unet"]


class UNet(nn.Module):
    def __init__(
        self,
        spatial_dims: int,
        in_channels: int,
        out_channels: int,
        channels: Sequence[int],
        strides: Sequence[int],
        kernel_size: Sequence[int] | int = 3,
        up_kernel_size: Sequence[int] | int = 3,
        num_res_units: int = 0,
        act="PRELU",
        norm="INSTANCE",
        dropout: float = 0.0,
        bias: bool = True,
        adn_ordering: str = "NDA",
    ) -> None:
        super().__init__()
        if len(channels) < 2:
            raise ValueError("the length of `channels` should be no less than 2.")
        delta = len(strides) - (len(channels) - 1)
        if delta < 0:
            raise ValueError("the length of `strides` should equal to `len(channels) - 1`.")
        if delta > 0:
            warnings.warn(f"`len(strides) > len(channels) - 1`, the last {delta} values of strides will not be used.")
        if isinstance(kernel_size, Sequence) and len(kernel_size) != spatial_dims:
            raise ValueError("the length of `kernel_size` should equal to `dimensions`.")
        if isinstance(up_kernel_size, Sequence) and len(up_kernel_size) != spatial_dims:
            raise ValueError("the length of `up_kernel_size` should equal to `dimensions`.")
        self.dimensions = spatial_dims
        self.in_channels = in_channels
        self.out_channels = out_channels
        self.channels = channels
        self.strides = strides
        self.kernel_size = kernel_size
        self.up_kernel_size = up_kernel_size
        self.num_res_units = num_res_units
        self.act = act
        self.norm = norm
        self.dropout = dropout
        self.bias = bias
        self.adn_ordering = adn_ordering
        self.model = self._level(in_channels, out_channels, list(channels), list(strides), True)

    # -- builders (same construction order as the reference so that seeded initialisation matches) -------------
    def _level(self, inc: int, outc: int, channels: list[int], strides: list[int], is_top: bool) -> nn.Module:
        c, s = channels[0], strides[0]
        if len(channels) > 2:
            sub = self._level(c, c, channels[1:], strides[1:], False)
            upc = c * 2
        else:
            sub = self._get_bottom_layer(c, channels[1])
            upc = c + channels[1]
        down = self._get_down_layer(inc, c, s, is_top)
        up = self._get_up_layer(upc, outc, s, is_top)
        return nn.Sequential(down, SkipConnection(sub), up)

    def _common(self) -> dict:
        return dict(act=self.act, norm=self.norm, dropout=self.dropout, bias=self.bias, adn_ordering=self.adn_ordering)

    def _get_down_layer(self, in_channels: int, out_channels: int, strides: int, is_top: bool) -> nn.Module:
        if self.num_res_units > 0:
            return ResidualUnit(
                self.dimensions, in_channels, out_channels, strides=strides, kernel_size=self.kernel_size,
                subunits=self.num_res_units, **self._common(),
            )
        return Convolution(self.dimensions, in_channels, out_channels, strides=strides, kernel_size=self.kernel_size, **self._common())

    def _get_bottom_layer(self, in_channels: int, out_channels: int) -> nn.Module:
        return self._get_down_layer(in_channels, out_channels, 1, False)

    def _get_up_layer(self, in_channels: int, out_channels: int, strides: int, is_top: bool) -> nn.Module:
        conv: nn.Module = Convolution(
            self.dimensions, in_channels, out_channels, strides=strides, kernel_size=self.up_kernel_size,
            conv_only=is_top and self.num_res_units == 0, is_transposed=True, **self._common(),
        )
        if self.num_res_units > 0:
            ru = ResidualUnit(
                self.dimensions, out_channels, out_channels, strides=1, kernel_size=self.kernel_size, subunits=1,
                last_conv_only=is_top, **self._common(),
            )
            conv = nn.Sequential(conv, ru)
        return conv

    def forward(self, x: torch.Tensor) -> torch.Tensor:
        if not x.is_cuda:
            raise RuntimeError("monai_b200.UNet runs on CUDA tensors only (there is no CPU fallback)")
        with torch.no_grad():
            return self.model(x.contiguous())


Unet = UNet
