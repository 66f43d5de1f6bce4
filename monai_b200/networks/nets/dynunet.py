"""DynUNet (monai/networks/nets/dynunet.py:24-380; nnU-Net style) behind the reference's constructor, module tree and state_dict
keys, with every convolution / normalisation / activation on the CUDA kernels of this package ("other
predictors used with sliding window").

Topology: input block, n down blocks, bottleneck, n + 1 up blocks (transposed convolution + skip concat + conv block), 1x1 output
block; anisotropic kernels / strides per level; optional residual blocks.  As in the reference the blocks are registered twice -- in
their flat containers (`input_block`, `downsamples`, `bottleneck`, `upsamples`) and in the recursive `skip_layers` chain -- so a
reference checkpoint loads key for key.  Inference only: deep-supervision heads are constructed (their parameters load) but only the
full-resolution output is produced, which is what the reference returns in eval mode.
"""
from __future__ import annotations

from typing import Sequence

import torch
import torch.nn as nn

from ..blocks.dynunet_block import UnetBasicBlock, UnetOutBlock, UnetResBlock, UnetUpBlock

__all__ = ["DynUNet", "DynUnet", "Dynunet"]


class DynUNetSkipLayer(nn.Module):
    """One level of the U (dynunet.py:24-53): down block, everything below, up block fed with the level's skip."""

    def __init__(self, index: int, downsample: nn.Module, upsample: nn.Module, next_layer: nn.Module, heads=None, super_head: nn.Module | None = None):
        super().__init__()
        self.downsample = downsample
        self.next_layer = next_layer
        self.upsample = upsample
        self.super_head = super_head
        self.heads = heads
        self.index = index

    def forward(self, x: torch.Tensor) -> torch.Tensor:
        down = self.downsample(x)
        return self.upsample(self.next_layer(down), down)   # supervision heads only matter in training mode


class DynUNet(nn.Module):
    def __init__(
        self,
        spatial_dims: int,
        in_channels: int,
        out_channels: int,
        kernel_size: Sequence[Sequence[int] | int],
        strides: Sequence[Sequence[int] | int],
        upsample_kernel_size: Sequence[Sequence[int] | int],
        filters: Sequence[int] | None = None,
        dropout=None,
        norm_name=("INSTANCE", {"affine": True}),
        act_name=("leakyrelu", {"inplace": True, "negative_slope": 0.01}),
        deep_supervision: bool = False,
        deep_supr_num: int = 1,
        res_block: bool = False,
        trans_bias: bool = False,
    ) -> None:
        super().__init__()
        self.spatial_dims, self.in_channels, self.out_channels = spatial_dims, in_channels, out_channels
        self.kernel_size, self.strides, self.upsample_kernel_size = kernel_size, strides, upsample_kernel_size
        self.norm_name, self.act_name, self.dropout = norm_name, act_name, dropout
        self.conv_block = UnetResBlock if res_block else UnetBasicBlock
        self.trans_bias = trans_bias
        if filters is not None:
            if len(filters) < len(strides):
                raise ValueError("length of filters should be no less than the length of strides.")
            self.filters = list(filters[: len(strides)])
        else:  # the nnU-Net rule: 32, 64, ... capped at 320 (3-D) / 512 (2-D)
            self.filters = [min(2 ** (5 + i), 320 if spatial_dims == 3 else 512) for i in range(len(strides))]
        f, ks, st = self.filters, kernel_size, strides
        common = dict(norm_name=norm_name, act_name=act_name, dropout=dropout)
        self.input_block = self.conv_block(spatial_dims, in_channels, f[0], ks[0], st[0], **common)
        self.downsamples = nn.ModuleList(
            [self.conv_block(spatial_dims, cin, cout, k, s, **common) for cin, cout, k, s in zip(f[:-2], f[1:-1], ks[1:-1], st[1:-1])])
        self.bottleneck = self.conv_block(spatial_dims, f[-2], f[-1], ks[-1], st[-1], **common)
        self.upsamples = nn.ModuleList([
            UnetUpBlock(spatial_dims, cin, cout, k, s, upsample_kernel_size=uk, trans_bias=trans_bias, **common)
            for cin, cout, k, s, uk in zip(f[1:][::-1], f[:-1][::-1], ks[1:][::-1], st[1:][::-1], upsample_kernel_size[::-1])])
        self.output_block = UnetOutBlock(spatial_dims, f[0], out_channels, dropout=dropout)
        self.deep_supervision, self.deep_supr_num = deep_supervision, deep_supr_num
        self.heads = [torch.rand(1)] * deep_supr_num
        if deep_supervision:
            self.deep_supervision_heads = nn.ModuleList([UnetOutBlock(spatial_dims, f[i + 1], out_channels, dropout=dropout) for i in range(deep_supr_num)])
            n_up = len(strides) - 1
            if deep_supr_num >= n_up:
                raise ValueError("deep_supr_num should be less than the number of up sample layers.")
            if deep_supr_num < 1:
                raise ValueError("deep_supr_num should be larger than 0.")
        self.apply(self.initialize_weights)
        self._check_kernel_stride()

        downs, ups = [self.input_block] + list(self.downsamples), list(self.upsamples)[::-1]
        heads = list(self.deep_supervision_heads) if deep_supervision else None

        def chain(index: int, downs, ups, heads):
            if len(downs) != len(ups):
                raise ValueError(f"{len(downs)} != {len(ups)}")
            if not downs:
                return self.bottleneck
            head, rest = None, heads
            if heads is not None and index > 0:   # the input block never gets a supervision head
                head, rest = (heads[0], heads[1:]) if heads else (None, [])
            nxt = chain(index + 1, downs[1:], ups[1:], rest)
            if head is not None:
                return DynUNetSkipLayer(index, downs[0], ups[0], nxt, heads=self.heads, super_head=head)
            return DynUNetSkipLayer(index, downs[0], ups[0], nxt)

        self.skip_layers = chain(0, downs, ups, heads)

    def _check_kernel_stride(self) -> None:
        ks, st = self.kernel_size, self.strides
        if len(ks) != len(st) or len(ks) < 3:
            raise ValueError("length of kernel_size and strides should be the same, and no less than 3.")
        for idx, (k, s) in enumerate(zip(ks, st)):
            if not isinstance(k, int) and len(k) != self.spatial_dims:
                raise ValueError(f"length of kernel_size in block {idx} should be the same as spatial_dims.")
            if not isinstance(s, int) and len(s) != self.spatial_dims:
                raise ValueError(f"length of stride in block {idx} should be the same as spatial_dims.")

    @staticmethod
    def initialize_weights(module: nn.Module) -> None:
        if isinstance(module, (nn.Conv3d, nn.Conv2d, nn.ConvTranspose3d, nn.ConvTranspose2d)):
            module.weight = nn.init.kaiming_normal_(module.weight, a=0.01)
            if module.bias is not None:
                module.bias = nn.init.constant_(module.bias, 0)

    def forward(self, x: torch.Tensor) -> torch.Tensor:
        if self.training and self.deep_supervision:
            raise RuntimeError("monai_b200.DynUNet is inference-only: call .eval() (deep-supervision outputs exist in training mode only)")
        return self.output_block(self.skip_layers(x))


DynUnet = Dynunet = DynUNet
