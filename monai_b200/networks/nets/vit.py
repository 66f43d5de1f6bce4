"""Vision Transformer encoder (monai/networks/nets/vit.py:24-133) as used by UNETR: patch embedding, `num_layers` TransformerBlocks,
final LayerNorm; returns (tokens, hidden states).  Tokens are channels-first [N, hidden, S] inside this package (the reference's
[N, S, hidden] is the transpose); classification heads are not built (the sliding-window path uses `classification=False`)."""
from __future__ import annotations

from typing import Sequence

import torch
import torch.nn as nn

from ... import _kernels as K
from ..blocks.transformerblock import PatchEmbeddingBlock, TransformerBlock

__all__ = ["ViT"]


class ViT(nn.Module):
    def __init__(self, in_channels: int, img_size: Sequence[int] | int, patch_size: Sequence[int] | int, hidden_size: int = 768, mlp_dim: int = 3072,
                 num_layers: int = 12, num_heads: int = 12, proj_type: str = "conv", pos_embed_type: str = "learnable", classification: bool = False,
                 num_classes: int = 2, dropout_rate: float = 0.0, spatial_dims: int = 3, post_activation="Tanh", qkv_bias: bool = False,
                 save_attn: bool = False) -> None:
        super().__init__()
        if not (0 <= dropout_rate <= 1):
            raise ValueError("dropout_rate should be between 0 and 1.")
        if hidden_size % num_heads != 0:
            raise ValueError("hidden_size should be divisible by num_heads.")
        if classification:
            raise NotImplementedError("monai_b200 ViT is the segmentation encoder (classification=False)")
        self.classification = False
        self.patch_embedding = PatchEmbeddingBlock(in_channels=in_channels, img_size=img_size, patch_size=patch_size, hidden_size=hidden_size,
                                                   num_heads=num_heads, proj_type=proj_type, pos_embed_type=pos_embed_type,
                                                   dropout_rate=dropout_rate, spatial_dims=spatial_dims)
        self.blocks = nn.ModuleList([TransformerBlock(hidden_size, mlp_dim, num_heads, dropout_rate, qkv_bias, save_attn) for _ in range(num_layers)])
        self.norm = nn.LayerNorm(hidden_size)

    def forward(self, x: torch.Tensor):
        x = self.patch_embedding(x)
        hidden_states_out = []
        for blk in self.blocks:
            x = blk(x)
            hidden_states_out.append(x)
        return K.layernorm_cf(x, self.norm.weight, self.norm.bias, self.norm.eps), hidden_states_out
