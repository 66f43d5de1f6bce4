"""UNETR decoder blocks (monai/networks/blocks/unetr_block.py:22-259) for any kernel size / stride, built from the generic DynUNet
blocks of this package (`dynunet_block.py`: every convolution / normalisation / activation on the CUDA kernels)."""
from __future__ import annotations

import torch
import torch.nn as nn

from ... import _kernels as K
from .dynunet_block import UnetBasicBlock, UnetResBlock, get_conv_layer

__all__ = ["UnetrUpBlock", "UnetrPrUpBlock", "UnetrBasicBlock"]


class UnetrUpBlock(nn.Module):
    """transposed convolution (kernel = stride), concat with the skip, residual / basic block (unetr_block.py:22-86)."""

    def __init__(self, spatial_dims: int, in_channels: int, out_channels: int, kernel_size, upsample_kernel_size, norm_name, res_block: bool = False) -> None:
        super().__init__()
        self.transp_conv = get_conv_layer(spatial_dims, in_channels, out_channels, kernel_size=upsample_kernel_size, stride=upsample_kernel_size,
                                          conv_only=True, is_transposed=True)
        blk = UnetResBlock if res_block else UnetBasicBlock
        self.conv_block = blk(spatial_dims, out_channels + out_channels, out_channels, kernel_size=kernel_size, stride=1, norm_name=norm_name)

    def forward(self, inp: torch.Tensor, skip: torch.Tensor) -> torch.Tensor:
        return self.conv_block(K.cat_channels([self.transp_conv(inp), skip]))


class UnetrPrUpBlock(nn.Module):
    """projection upsampling: an initial transposed convolution, then `num_layer` x (transposed convolution [+ conv block]) (unetr_block.py:89-205)."""

    def __init__(self, spatial_dims: int, in_channels: int, out_channels: int, num_layer: int, kernel_size, stride, upsample_kernel_size,
                 norm_name, conv_block: bool = False, res_block: bool = False) -> None:
        super().__init__()
        up = dict(kernel_size=upsample_kernel_size, stride=upsample_kernel_size, conv_only=True, is_transposed=True)
        self.transp_conv_init = get_conv_layer(spatial_dims, in_channels, out_channels, **up)
        if conv_block:
            blk = UnetResBlock if res_block else UnetBasicBlock
            self.blocks = nn.ModuleList([
                nn.Sequential(get_conv_layer(spatial_dims, out_channels, out_channels, **up),
                              blk(spatial_dims=spatial_dims, in_channels=out_channels, out_channels=out_channels, kernel_size=kernel_size,
                                  stride=stride, norm_name=norm_name))
                for _ in range(num_layer)])
        else:
            self.blocks = nn.ModuleList([get_conv_layer(spatial_dims, out_channels, out_channels, **up) for _ in range(num_layer)])

    def forward(self, x: torch.Tensor) -> torch.Tensor:
        x = self.transp_conv_init(x)
        for blk in self.blocks:
            x = blk(x)
        return x


class UnetrBasicBlock(nn.Module):
    def __init__(self, spatial_dims: int, in_channels: int, out_channels: int, kernel_size, stride, norm_name, res_block: bool = False) -> None:
        super().__init__()
        blk = UnetResBlock if res_block else UnetBasicBlock
        self.layer = blk(spatial_dims=spatial_dims, in_channels=in_channels, out_channels=out_channels, kernel_size=kernel_size, stride=stride,
                         norm_name=norm_name)

    def forward(self, inp: torch.Tensor) -> torch.Tensor:
        return self.layer(inp)
