"""ViT building blocks on CUDA kernels, tokens kept channels-first [N, C, S]:

  * `MLPBlock`        monai/networks/blocks/mlp.py:25-80        linear1 - GELU - linear2
  * `SABlock`         monai/networks/blocks/selfattention.py    combined qkv projection, softmax(q k^T * scale) v, out_proj
  * `TransformerBlock` monai/networks/blocks/transformerblock.py x + attn(norm1(x)); x + mlp(norm2(x))
  * `PatchEmbeddingBlock` monai/networks/blocks/patchembedding.py:28-138 (proj_type "conv") + learnable position embeddings

Same parameter names / shapes as the reference (including the never-used `norm_cross_attn` / `cross_attn` containers every
TransformerBlock registers), so reference checkpoints load key for key.  nn.Linear runs as a 1x1x1 `b200_conv3d_direct` on the
[N, C, 1, 1, S] view (fp32 accumulation), LayerNorm and attention are `b200_layernorm_cf` / `b200_mhsa_cf`, GELU and the residual
adds are passes of `b200_norm_act`.  Inference only: dropout is the identity.
"""
from __future__ import annotations

import math
from typing import Sequence

import torch
import torch.nn as nn

from ... import _kernels as K
from ... import _lib as L

__all__ = ["MLPBlock", "SABlock", "CrossAttentionBlock", "TransformerBlock", "PatchEmbeddingBlock", "linear_cf"]


def linear_cf(x: torch.Tensor, lin: nn.Linear) -> torch.Tensor:
    """nn.Linear over the channel axis of channels-first tokens x[N, Cin, S] -> [N, Cout, S]."""
    N, Cin, S = x.shape
    y = K.conv3d_direct(x.reshape(N, Cin, 1, 1, S), lin.weight.reshape(lin.out_features, lin.in_features, 1, 1, 1), lin.bias)
    return y.reshape(N, lin.out_features, S)


def _add(a: torch.Tensor, b: torch.Tensor) -> torch.Tensor:
    return K.norm_act(a, None, res=b, act=L.ACT_NONE)


class MLPBlock(nn.Module):
    def __init__(self, hidden_size: int, mlp_dim: int, dropout_rate: float = 0.0, act="GELU", dropout_mode: str = "vit") -> None:
        super().__init__()
        if not (0 <= dropout_rate <= 1):
            raise ValueError("dropout_rate should be between 0 and 1.")
        if str(act if isinstance(act, str) else act[0]).upper() != "GELU":
            raise NotImplementedError("monai_b200 MLPBlock implements the GELU activation")
        mlp_dim = mlp_dim or hidden_size
        self.linear1 = nn.Linear(hidden_size, mlp_dim)
        self.linear2 = nn.Linear(mlp_dim, hidden_size)
        self.fn = nn.GELU()
        self.drop1 = nn.Dropout(dropout_rate)
        self.drop2 = nn.Dropout(dropout_rate) if dropout_mode == "vit" else self.drop1

    def forward(self, x: torch.Tensor) -> torch.Tensor:
        return linear_cf(K.norm_act(linear_cf(x, self.linear1), None, act=L.ACT_GELU), self.linear2)


class SABlock(nn.Module):
    def __init__(self, hidden_size: int, num_heads: int, dropout_rate: float = 0.0, qkv_bias: bool = False, save_attn: bool = False) -> None:
        super().__init__()
        if not (0 <= dropout_rate <= 1):
            raise ValueError("dropout_rate should be between 0 and 1.")
        if hidden_size % num_heads != 0:
            raise ValueError("hidden size should be divisible by num_heads.")
        if save_attn:
            raise NotImplementedError("monai_b200 SABlock does not materialise the attention matrix (save_attn)")
        self.num_heads, self.dim_head = num_heads, hidden_size // num_heads
        self.out_proj = nn.Linear(hidden_size, hidden_size)
        self.qkv = nn.Linear(hidden_size, hidden_size * 3, bias=qkv_bias)
        self.to_q = self.to_k = self.to_v = nn.Identity()
        self.drop_output, self.drop_weights = nn.Dropout(dropout_rate), nn.Dropout(dropout_rate)
        self.scale = self.dim_head**-0.5

    def forward(self, x: torch.Tensor) -> torch.Tensor:
        return linear_cf(K.mhsa_cf(linear_cf(x, self.qkv), self.num_heads, self.dim_head, self.scale), self.out_proj)


class CrossAttentionBlock(nn.Module):
    """Parameter container only: the reference's TransformerBlock registers it even when `with_cross_attention` is False."""

    def __init__(self, hidden_size: int, num_heads: int, dropout_rate: float = 0.0, qkv_bias: bool = False) -> None:
        super().__init__()
        self.out_proj = nn.Linear(hidden_size, hidden_size)
        self.to_q = nn.Linear(hidden_size, hidden_size, bias=qkv_bias)
        self.to_k = nn.Linear(hidden_size, hidden_size, bias=qkv_bias)
        self.to_v = nn.Linear(hidden_size, hidden_size, bias=qkv_bias)


class TransformerBlock(nn.Module):
    def __init__(self, hidden_size: int, mlp_dim: int, num_heads: int, dropout_rate: float = 0.0, qkv_bias: bool = False, save_attn: bool = False) -> None:
        super().__init__()
        if not (0 <= dropout_rate <= 1):
            raise ValueError("dropout_rate should be between 0 and 1.")
        if hidden_size % num_heads != 0:
            raise ValueError("hidden_size should be divisible by num_heads.")
        self.mlp = MLPBlock(hidden_size, mlp_dim, dropout_rate)
        self.norm1 = nn.LayerNorm(hidden_size)
        self.attn = SABlock(hidden_size, num_heads, dropout_rate, qkv_bias=qkv_bias, save_attn=save_attn)
        self.norm2 = nn.LayerNorm(hidden_size)
        self.with_cross_attention = False
        self.norm_cross_attn = nn.LayerNorm(hidden_size)
        self.cross_attn = CrossAttentionBlock(hidden_size, num_heads, dropout_rate, qkv_bias)

    def forward(self, x: torch.Tensor) -> torch.Tensor:
        x = _add(self.attn(K.layernorm_cf(x, self.norm1.weight, self.norm1.bias, self.norm1.eps)), x)
        return _add(self.mlp(K.layernorm_cf(x, self.norm2.weight, self.norm2.bias, self.norm2.eps)), x)


class PatchEmbeddingBlock(nn.Module):
    """Strided-convolution patch projection + position embeddings; output channels-first [N, hidden, n_patches]."""

    def __init__(self, in_channels: int, img_size: Sequence[int] | int, patch_size: Sequence[int] | int, hidden_size: int, num_heads: int,
                 proj_type: str = "conv", pos_embed_type: str = "learnable", dropout_rate: float = 0.0, spatial_dims: int = 3) -> None:
        super().__init__()
        if not (0 <= dropout_rate <= 1):
            raise ValueError(f"dropout_rate {dropout_rate} should be between 0 and 1.")
        if hidden_size % num_heads != 0:
            raise ValueError(f"hidden size {hidden_size} should be divisible by num_heads {num_heads}.")
        if proj_type != "conv":
            raise NotImplementedError("monai_b200 PatchEmbeddingBlock implements proj_type='conv'")
        if pos_embed_type not in ("learnable", "none"):
            raise NotImplementedError("monai_b200 PatchEmbeddingBlock implements pos_embed_type 'learnable' and 'none'")
        img = (img_size,) * spatial_dims if isinstance(img_size, int) else tuple(img_size)
        pat = (patch_size,) * spatial_dims if isinstance(patch_size, int) else tuple(patch_size)
        for m, p in zip(img, pat):
            if m < p:
                raise ValueError("patch_size should be smaller than img_size.")
        self.n_patches = int(math.prod(im // p for im, p in zip(img, pat)))
        conv = {1: nn.Conv1d, 2: nn.Conv2d, 3: nn.Conv3d}[spatial_dims]
        self.patch_embeddings = conv(in_channels=in_channels, out_channels=hidden_size, kernel_size=pat, stride=pat)
        self.position_embeddings = nn.Parameter(torch.zeros(1, self.n_patches, hidden_size))
        self.dropout = nn.Dropout(dropout_rate)
        if pos_embed_type == "learnable":
            nn.init.trunc_normal_(self.position_embeddings, mean=0.0, std=0.02, a=-2.0, b=2.0)
        self.apply(self._init_weights)

    @staticmethod
    def _init_weights(m: nn.Module) -> None:
        if isinstance(m, nn.Linear):
            nn.init.trunc_normal_(m.weight, mean=0.0, std=0.02, a=-2.0, b=2.0)
            if m.bias is not None:
                nn.init.constant_(m.bias, 0)
        elif isinstance(m, nn.LayerNorm):
            nn.init.constant_(m.bias, 0)
            nn.init.constant_(m.weight, 1.0)

    def forward(self, x: torch.Tensor) -> torch.Tensor:
        conv = self.patch_embeddings
        nd = x.dim() - 2
        x3 = x.reshape(x.shape[0], x.shape[1], *([1] * (3 - nd)), *x.shape[2:])
        patch = (1,) * (3 - nd) + tuple(int(k) for k in conv.kernel_size)
        # kernel = stride = patch: the convolution is a Linear over the flattened patch (the direct convolution kernel stops at 4096 taps)
        cols = K.patchify(x3, patch)                                                # [N, Cin * prod(patch), S]
        w = conv.weight.reshape(conv.out_channels, -1, 1, 1, 1)
        y = K.conv3d_direct(cols.reshape(cols.shape[0], cols.shape[1], 1, 1, cols.shape[2]), w, conv.bias).reshape(cols.shape[0], conv.out_channels, -1)
        pos = self.position_embeddings.detach().transpose(1, 2).contiguous().to(y.dtype)   # [1, hidden, S]
        return K.norm_act(y, None, res=pos.expand(y.shape[0], -1, -1), act=L.ACT_NONE)
