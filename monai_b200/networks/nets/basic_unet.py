"""`BasicUNet` (monai/networks/nets/basic_unet.py:27-282) on the monai_b200 CUDA kernels.

Module tree and state_dict keys follow the reference (`conv_0.conv_0.conv.weight`, `down_1.convs.conv_1.adn.N.weight`,
`upcat_4.upsample.deconv.weight`, `final_conv.weight`, ...).  Only the default `upsample="deconv"` mode is on the
north-star path; the other modes raise NotImplementedError.
"""
from __future__ import annotations

from collections.abc import Sequence

import torch
import torch.nn as nn

from ... import _kernels as K
from ..blocks.convolutions import Convolution, run_conv_module

__all__ = ["BasicUNet", "BasicUnet", "Basicunet", "basicunet", "TwoConv", "Down", "UpCat"]

_CONV = {1: nn.Conv1d, 2: nn.Conv2d, 3: nn.Conv3d}
_CONVT = {1: nn.ConvTranspose1d, 2: nn.ConvTranspose2d, 3: nn.ConvTranspose3d}
_POOL = {1: nn.MaxPool1d, 2: nn.MaxPool2d, 3: nn.MaxPool3d}


class TwoConv(nn.Sequential):
    """two convolutions (basic_unet.py:27-58)."""

    def __init__(self, spatial_dims: int, in_chns: int, out_chns: int, act, norm, bias: bool, dropout=0.0):
        super().__init__()
        self.add_module("conv_0", Convolution(spatial_dims, in_chns, out_chns, act=act, norm=norm, dropout=dropout, bias=bias, padding=1))
        self.add_module("conv_1", Convolution(spatial_dims, out_chns, out_chns, act=act, norm=norm, dropout=dropout, bias=bias, padding=1))


class Down(nn.Sequential):
    """max-pool by 2 then TwoConv (basic_unet.py:61-89)."""

    def __init__(self, spatial_dims: int, in_chns: int, out_chns: int, act, norm, bias: bool, dropout=0.0):
        super().__init__()
        self.add_module("max_pooling", _POOL[spatial_dims](kernel_size=2))
        self.add_module("convs", TwoConv(spatial_dims, in_chns, out_chns, act, norm, bias, dropout))

    def forward(self, x: torch.Tensor) -> torch.Tensor:  # type: ignore[override]
        nd = x.dim() - 2
        x3 = x.reshape(x.shape[0], x.shape[1], *([1] * (3 - nd)), *x.shape[2:])
        if nd != 3:
            raise NotImplementedError("monai_b200 BasicUNet pooling is implemented for 3 spatial dims")
        return self.convs(K.maxpool3d_2(x3))


class _UpSampleDeconv(nn.Module):
    """UpSample(mode="deconv") (monai/networks/blocks/upsample.py:102-116): ConvTranspose(k=2, s=2)."""

    def __init__(self, spatial_dims: int, in_chns: int, out_chns: int):
        super().__init__()
        self.add_module("deconv", _CONVT[spatial_dims](in_chns, out_chns, kernel_size=2, stride=2, bias=True))

    def forward(self, x: torch.Tensor) -> torch.Tensor:
        return run_conv_module(self.deconv, x)


class UpCat(nn.Module):
    """upsample, replicate-pad to the skip's size, concatenate [skip, up], TwoConv (basic_unet.py:92-175)."""

    def __init__(self, spatial_dims, in_chns, cat_chns, out_chns, act, norm, bias, dropout=0.0, upsample="deconv", halves=True, is_pad=True):
        super().__init__()
        if upsample != "deconv":
            raise NotImplementedError("monai_b200 BasicUNet implements upsample='deconv' only")
        up_chns = in_chns // 2 if halves else in_chns
        self.upsample = _UpSampleDeconv(spatial_dims, in_chns, up_chns)
        self.convs = TwoConv(spatial_dims, cat_chns + up_chns, out_chns, act, norm, bias, dropout)
        self.is_pad = is_pad

    def forward(self, x: torch.Tensor, x_e: torch.Tensor | None) -> torch.Tensor:
        x_0 = self.upsample(x)
        if x_e is None:
            return self.convs(x_0)
        n, ce = x_e.shape[:2]
        cu = x_0.shape[1]
        sp = tuple(x_e.shape[2:]) if self.is_pad else tuple(x_0.shape[2:])
        if any(a - b not in (0, 1) for a, b in zip(sp, x_0.shape[2:])):
            raise ValueError(f"skip {tuple(x_e.shape)} and upsampled {tuple(x_0.shape)} tensors differ by more than one voxel")
        cat = torch.empty((n, ce + cu, *sp), device=x.device, dtype=x.dtype)
        K.copy_channels(x_e, cat, 0)
        K.copy_channels(x_0, cat, ce)  # replicate-pads the high side when the skip is one voxel larger
        return self.convs(cat)


class BasicUNet(nn.Module):
    def __init__(
        self,
        spatial_dims: int = 3,
        in_channels: int = 1,
        out_channels: int = 2,
        features: Sequence[int] = (32, 32, 64, 128, 256, 32),
        act=("LeakyReLU", {"negative_slope": 0.1, "inplace": True}),
        norm=("instance", {"affine": True}),
        bias: bool = True,
        dropout=0.0,
        upsample: str = "deconv",
    ):
        super().__init__()
        if spatial_dims != 3:
            raise NotImplementedError("monai_b200 BasicUNet is implemented for spatial_dims=3 (the sliding-window hot path)")
        fea = tuple(features)
        if len(fea) != 6:
            raise ValueError(f"Sequence must have length 6, got {len(fea)}.")
        print(f"BasicUNet features: {fea}.")
        self.conv_0 = TwoConv(spatial_dims, in_channels, fea[0], act, norm, bias, dropout)
        self.down_1 = Down(spatial_dims, fea[0], fea[1], act, norm, bias, dropout)
        self.down_2 = Down(spatial_dims, fea[1], fea[2], act, norm, bias, dropout)
        self.down_3 = Down(spatial_dims, fea[2], fea[3], act, norm, bias, dropout)
        self.down_4 = Down(spatial_dims, fea[3], fea[4], act, norm, bias, dropout)
        self.upcat_4 = UpCat(spatial_dims, fea[4], fea[3], fea[3], act, norm, bias, dropout, upsample)
        self.upcat_3 = UpCat(spatial_dims, fea[3], fea[2], fea[2], act, norm, bias, dropout, upsample)
        self.upcat_2 = UpCat(spatial_dims, fea[2], fea[1], fea[1], act, norm, bias, dropout, upsample)
        self.upcat_1 = UpCat(spatial_dims, fea[1], fea[0], fea[5], act, norm, bias, dropout, upsample, halves=False)
        self.final_conv = _CONV[spatial_dims](fea[5], out_channels, kernel_size=1)

    def forward(self, x: torch.Tensor) -> torch.Tensor:
        if not x.is_cuda:
            raise RuntimeError("monai_b200.BasicUNet runs on CUDA tensors only (there is no CPU fallback)")
        with torch.no_grad():
            x0 = self.conv_0(x.contiguous())
            x1 = self.down_1(x0)
            x2 = self.down_2(x1)
            x3 = self.down_3(x2)
            x4 = self.down_4(x3)
            u4 = self.upcat_4(x4, x3)
            u3 = self.upcat_3(u4, x2)
            u2 = self.upcat_2(u3, x1)
            u1 = self.upcat_1(u2, x0)
            return run_conv_module(self.final_conv, u1)


BasicUnet = Basicunet = basicunet = BasicUNet
