"""UNETR (monai/networks/nets/unetr.py:24-213; Hatamizadeh et al.) behind the reference's constructor, module tree and state_dict
keys: a ViT encoder on 16^3 patches whose hidden states 3 / 6 / 9 / 12 feed a convolutional decoder
(projection up-blocks, transposed convolutions, residual blocks).  Every operation runs on the CUDA kernels of this package in the
fp32-faithful generic forms (b200_conv3d_direct for convolutions AND linear layers, b200_layernorm_cf, b200_mhsa_cf,
b200_instnorm_stats + b200_norm_act); inference only."""
from __future__ import annotations

from typing import Sequence

import torch
import torch.nn as nn

from ..blocks.dynunet_block import UnetOutBlock
from ..blocks.unetr_block import UnetrBasicBlock, UnetrPrUpBlock, UnetrUpBlock
from .vit import ViT

__all__ = ["UNETR"]


class UNETR(nn.Module):
    def __init__(self, in_channels: int, out_channels: int, img_size: Sequence[int] | int, feature_size: int = 16, hidden_size: int = 768,
                 mlp_dim: int = 3072, num_heads: int = 12, proj_type: str = "conv", norm_name="instance", conv_block: bool = True,
                 res_block: bool = True, dropout_rate: float = 0.0, spatial_dims: int = 3, qkv_bias: bool = False, save_attn: bool = False) -> None:
        super().__init__()
        if not (0 <= dropout_rate <= 1):
            raise ValueError("dropout_rate should be between 0 and 1.")
        if hidden_size % num_heads != 0:
            raise ValueError("hidden_size should be divisible by num_heads.")
        self.num_layers = 12
        img_size = (img_size,) * spatial_dims if isinstance(img_size, int) else tuple(img_size)
        self.patch_size = (16,) * spatial_dims
        self.feat_size = tuple(i // p for i, p in zip(img_size, self.patch_size))
        self.hidden_size = hidden_size
        self.classification = False
        self.vit = ViT(in_channels=in_channels, img_size=img_size, patch_size=self.patch_size, hidden_size=hidden_size, mlp_dim=mlp_dim,
                       num_layers=self.num_layers, num_heads=num_heads, proj_type=proj_type, classification=False, dropout_rate=dropout_rate,
                       spatial_dims=spatial_dims, qkv_bias=qkv_bias, save_attn=save_attn)
        sd, f = spatial_dims, feature_size
        self.encoder1 = UnetrBasicBlock(sd, in_channels, f, kernel_size=3, stride=1, norm_name=norm_name, res_block=res_block)
        pr = dict(kernel_size=3, stride=1, upsample_kernel_size=2, norm_name=norm_name, conv_block=conv_block, res_block=res_block)
        self.encoder2 = UnetrPrUpBlock(sd, hidden_size, f * 2, num_layer=2, **pr)
        self.encoder3 = UnetrPrUpBlock(sd, hidden_size, f * 4, num_layer=1, **pr)
        self.encoder4 = UnetrPrUpBlock(sd, hidden_size, f * 8, num_layer=0, **pr)
        up = dict(kernel_size=3, upsample_kernel_size=2, norm_name=norm_name, res_block=res_block)
        self.decoder5 = UnetrUpBlock(sd, hidden_size, f * 8, **up)
        self.decoder4 = UnetrUpBlock(sd, f * 8, f * 4, **up)
        self.decoder3 = UnetrUpBlock(sd, f * 4, f * 2, **up)
        self.decoder2 = UnetrUpBlock(sd, f * 2, f, **up)
        self.out = UnetOutBlock(spatial_dims=sd, in_channels=f, out_channels=out_channels)

    def proj_feat(self, x: torch.Tensor) -> torch.Tensor:
        """tokens -> feature map.  The reference permutes [N, S, hidden] to [N, hidden, *feat_size]; tokens are already channels-first here."""
        return x.reshape(x.shape[0], self.hidden_size, *self.feat_size)

    def forward(self, x_in: torch.Tensor) -> torch.Tensor:
        if tuple(x_in.shape[2:]) != tuple(f * p for f, p in zip(self.feat_size, self.patch_size)):
            raise ValueError(f"UNETR was built for inputs of size {tuple(f * p for f, p in zip(self.feat_size, self.patch_size))}, got {tuple(x_in.shape[2:])}")
        x, hidden = self.vit(x_in)
        enc1 = self.encoder1(x_in)
        enc2 = self.encoder2(self.proj_feat(hidden[3]))
        enc3 = self.encoder3(self.proj_feat(hidden[6]))
        enc4 = self.encoder4(self.proj_feat(hidden[9]))
        dec3 = self.decoder5(self.proj_feat(x), enc4)
        dec2 = self.decoder4(dec3, enc3)
        dec1 = self.decoder3(dec2, enc2)
        return self.out(self.decoder2(dec1, enc1))
