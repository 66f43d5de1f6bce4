"""SegResNet (monai/networks/nets/segresnet.py:30-200; Myronenko 2018, without the VAE branch) behind the reference's constructor,
module tree and state_dict keys, on the CUDA kernels of this package.

convInit, then per level [stride-2 convolution] + pre-activation ResBlocks (GroupNorm by default); the decoder halves the channels
with a 1x1 convolution, upsamples x2 (trilinear or transposed convolution), adds the encoder feature and runs ResBlocks; final
norm - act - 1x1 convolution.  Inference only (dropout is the identity).
"""
from __future__ import annotations

import torch
import torch.nn as nn

from ... import _kernels as K
from ... import _lib as L
from ..blocks.acti_norm import norm_act_from_modules
from ..blocks.segresnet_block import ResBlock, get_conv_layer, get_upsample_layer
from ..layers.factories import get_act_layer, get_norm_layer

__all__ = ["SegResNet"]


class SegResNet(nn.Module):
    def __init__(
        self,
        spatial_dims: int = 3,
        init_filters: int = 8,
        in_channels: int = 1,
        out_channels: int = 2,
        dropout_prob: float | None = None,
        act=("RELU", {"inplace": True}),
        norm=("GROUP", {"num_groups": 8}),
        norm_name: str = "",
        num_groups: int = 8,
        use_conv_final: bool = True,
        blocks_down: tuple = (1, 2, 2, 4),
        blocks_up: tuple = (1, 1, 1),
        upsample_mode: str = "nontrainable",
    ):
        super().__init__()
        if spatial_dims not in (2, 3):
            raise ValueError("`spatial_dims` can only be 2 or 3.")
        self.spatial_dims, self.init_filters, self.in_channels = spatial_dims, init_filters, in_channels
        self.blocks_down, self.blocks_up, self.dropout_prob = blocks_down, blocks_up, dropout_prob
        self.act = act
        self.act_mod = get_act_layer(act)
        if norm_name:
            if norm_name.lower() != "group":
                raise ValueError(f"Deprecating option 'norm_name={norm_name}', please use 'norm' instead.")
            norm = ("group", {"num_groups": num_groups})
        self.norm = norm
        self.upsample_mode = str(getattr(upsample_mode, "value", upsample_mode))
        self.use_conv_final = use_conv_final
        self.convInit = get_conv_layer(spatial_dims, in_channels, init_filters)
        self.down_layers = nn.ModuleList()
        for i, n_blocks in enumerate(blocks_down):
            ch = init_filters * 2**i
            pre = get_conv_layer(spatial_dims, ch // 2, ch, stride=2) if i > 0 else nn.Identity()
            self.down_layers.append(nn.Sequential(pre, *[ResBlock(spatial_dims, ch, norm=norm, act=act) for _ in range(n_blocks)]))
        self.up_layers, self.up_samples = nn.ModuleList(), nn.ModuleList()
        n_up = len(blocks_up)
        for i in range(n_up):
            ch = init_filters * 2 ** (n_up - i)
            self.up_layers.append(nn.Sequential(*[ResBlock(spatial_dims, ch // 2, norm=norm, act=act) for _ in range(blocks_up[i])]))
            self.up_samples.append(nn.Sequential(get_conv_layer(spatial_dims, ch, ch // 2, kernel_size=1),
                                                 get_upsample_layer(spatial_dims, ch // 2, upsample_mode=self.upsample_mode)))
        self.conv_final = nn.Sequential(get_norm_layer(name=norm, spatial_dims=spatial_dims, channels=init_filters), self.act_mod,
                                        get_conv_layer(spatial_dims, init_filters, out_channels, kernel_size=1, bias=True))
        if dropout_prob is not None:
            self.dropout = (nn.Dropout, nn.Dropout2d, nn.Dropout3d)[spatial_dims - 1](dropout_prob)

    def encode(self, x: torch.Tensor):
        if self.training and self.dropout_prob:
            raise RuntimeError("monai_b200.SegResNet is inference-only: call .eval()")
        x = self.convInit(x)
        down_x = []
        for down in self.down_layers:
            for m in down:
                if not isinstance(m, nn.Identity):
                    x = m(x)
            down_x.append(x)
        return x, down_x

    def decode(self, x: torch.Tensor, down_x) -> torch.Tensor:
        for i, (up, upl) in enumerate(zip(self.up_samples, self.up_layers)):
            x = K.norm_act(up[1](up[0](x)), None, res=down_x[i + 1], act=L.ACT_NONE)   # up(x) + skip
            for m in upl:
                x = m(x)
        if self.use_conv_final:
            x = self.conv_final[2](norm_act_from_modules(x, self.conv_final[0], self.conv_final[1]))
        return x

    def forward(self, x: torch.Tensor) -> torch.Tensor:
        x, down_x = self.encode(x)
        down_x.reverse()
        return self.decode(x, down_x)
