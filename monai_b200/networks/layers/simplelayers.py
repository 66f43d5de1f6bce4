"""`SkipConnection` (monai/networks/layers/simplelayers.py:103-135): cat / add / mul of a branch with its input."""
from __future__ import annotations

import torch
import torch.nn as nn

from ... import _kernels as K

__all__ = ["SkipConnection"]


class SkipConnection(nn.Module):
    def __init__(self, submodule: nn.Module, dim: int = 1, mode: str = "cat") -> None:
        super().__init__()
        self.submodule = submodule
        self.dim = dim
        mode = str(getattr(mode, "value", mode)).lower()
        if mode not in ("cat", "add", "mul"):
            raise ValueError(f"Unsupported skip mode {mode!r}; available: cat, add, mul.")
        self.mode = mode

    def forward(self, x: torch.Tensor) -> torch.Tensor:
        y = self.submodule(x)
        if self.mode == "cat":
            if self.dim != 1:
                raise NotImplementedError("monai_b200 SkipConnection concatenates along the channel axis only")
            return K.cat_channels([x, y])
        if self.mode == "add":
            return K.norm_act(x, res=y)
        raise NotImplementedError("SkipConnection(mode='mul') is not on the sliding-window hot path")
