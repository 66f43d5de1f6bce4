"""ADN: ordered Norm / Dropout / Activation (monai/networks/blocks/acti_norm.py:19-101) on CUDA kernels.

The child modules "N", "D", "A" are torch.nn parameter containers with the reference's names, so state_dict keys
(`...adn.A.weight`, `...adn.N.weight`) are identical.  `forward` never calls them: InstanceNorm statistics come
from `b200_instnorm_stats` and normalise + affine + activation (+ residual) is ONE pass of `b200_norm_act`.
"""
from __future__ import annotations

import torch
import torch.nn as nn

from ... import _kernels as K
from ... import _lib as L
from ..layers.factories import get_act_layer, get_dropout_layer, get_norm_layer

__all__ = ["ADN", "norm_act_from_modules"]


def _act_code(act: nn.Module | None):
    """torch activation container -> (kernel act code, scalar slope, per-channel slope tensor)."""
    if act is None or isinstance(act, nn.Identity):
        return L.ACT_NONE, 0.0, None
    if isinstance(act, nn.PReLU):
        return L.ACT_PRELU, 0.0, act.weight
    if isinstance(act, nn.LeakyReLU):
        return L.ACT_LEAKY, float(act.negative_slope), None
    if isinstance(act, nn.ReLU):
        return L.ACT_RELU, 0.0, None
    if isinstance(act, nn.GELU):
        if getattr(act, "approximate", "none") != "none":
            raise NotImplementedError("only the exact (erf) GELU is implemented")
        return L.ACT_GELU, 0.0, None
    raise NotImplementedError(f"activation {type(act).__name__} has no monai_b200 kernel")


def norm_act_from_modules(
    x: torch.Tensor,
    norm: nn.Module | None,
    act: nn.Module | None,
    res: torch.Tensor | None = None,
    res_norm: nn.Module | None = None,
    out: torch.Tensor | None = None,
) -> torch.Tensor:
    """y = act(norm(x) [+ res_norm(res)]) in one elementwise pass (plus one statistics pass per instance norm)."""
    code, slope, slope_t = _act_code(act)
    stats = gamma = beta = res_stats = None
    eps = 1e-5
    if norm is not None and not isinstance(norm, nn.Identity):
        if isinstance(norm, (nn.InstanceNorm1d, nn.InstanceNorm2d, nn.InstanceNorm3d)):
            if norm.track_running_stats:
                raise NotImplementedError("InstanceNorm with running statistics is not supported")
            stats, eps, gamma, beta = K.instnorm_stats(x), norm.eps, norm.weight, norm.bias
        elif isinstance(norm, nn.GroupNorm):
            # the channels of a group are contiguous in NCDHW: group statistics = "instance" statistics of the [N, G, (C/G)*D, H, W] view;
            # handing every channel its group's sums divided by C/G makes b200_norm_act's mean = sum / S the group mean (same for E[x^2])
            G = norm.num_groups
            N_, C_ = x.shape[:2]
            cg = C_ // G
            if x.dim() < 3 or not x[0].is_contiguous():
                raise NotImplementedError("GroupNorm needs a [N, C, *spatial] input that is contiguous per sample")
            gstats = K.instnorm_stats(x.reshape(N_, G, cg * x.shape[2], *x.shape[3:]))
            stats = (gstats / float(cg)).repeat_interleave(cg, dim=0).contiguous()   # [N*C, 2]: a tiny table
            eps, gamma, beta = norm.eps, norm.weight, norm.bias
        elif isinstance(norm, (nn.BatchNorm1d, nn.BatchNorm2d, nn.BatchNorm3d)):
            if norm.training:
                raise RuntimeError("monai_b200 is inference-only: call .eval() before running BatchNorm layers")
            scale = (norm.weight if norm.affine else 1.0) / torch.sqrt(norm.running_var + norm.eps)
            gamma = scale.float()
            beta = ((norm.bias if norm.affine else 0.0) - norm.running_mean * scale).float()
        else:
            raise NotImplementedError(f"normalisation {type(norm).__name__} has no monai_b200 kernel")
    if res is not None and res_norm is not None and not isinstance(res_norm, nn.Identity):
        if not isinstance(res_norm, (nn.InstanceNorm1d, nn.InstanceNorm2d, nn.InstanceNorm3d)) or res_norm.affine:
            raise NotImplementedError("residual-branch normalisation must be a non-affine InstanceNorm")
        res_stats = K.instnorm_stats(res)
    return K.norm_act(x, stats, eps, gamma, beta, res, res_stats, code, slope, slope_t, out)


class ADN(nn.Sequential):
    def __init__(
        self,
        ordering: str = "NDA",
        in_channels: int | None = None,
        act="RELU",
        norm=None,
        norm_dim: int | None = None,
        dropout=None,
        dropout_dim: int | None = None,
    ) -> None:
        super().__init__()
        ops: dict[str, nn.Module | None] = {"A": None, "D": None, "N": None}
        if norm is not None:
            if norm_dim is None and dropout_dim is None:
                raise ValueError("norm_dim or dropout_dim needs to be specified.")
            ops["N"] = get_norm_layer(name=norm, spatial_dims=norm_dim or dropout_dim, channels=in_channels)
        if act is not None:
            ops["A"] = get_act_layer(act)
        if dropout is not None:
            if norm_dim is None and dropout_dim is None:
                raise ValueError("norm_dim or dropout_dim needs to be specified.")
            ops["D"] = get_dropout_layer(name=dropout, dropout_dim=dropout_dim or norm_dim)
        for item in ordering.upper():
            if item not in ops:
                raise ValueError(f"ordering must be a string of {ops}, got {item} in it.")
            if ops[item] is not None:
                self.add_module(item, ops[item])

    def forward(self, x: torch.Tensor) -> torch.Tensor:  # type: ignore[override]
        seq = list(self.named_children())
        for n, m in seq:
            if n == "D" and m.training and getattr(m, "p", 0.0) > 0:
                raise RuntimeError("monai_b200 is inference-only: dropout layers must be in eval mode")
        seq = [(n, m) for n, m in seq if n != "D"]  # dropout is the identity at inference time
        i = 0
        while i < len(seq):
            n, m = seq[i]
            if n == "N" and i + 1 < len(seq) and seq[i + 1][0] == "A":  # fuse norm + activation
                x = norm_act_from_modules(x, m, seq[i + 1][1])
                i += 2
            elif n == "N":
                x = norm_act_from_modules(x, m, None)
                i += 1
            else:
                x = norm_act_from_modules(x, None, m)
                i += 1
        return x
