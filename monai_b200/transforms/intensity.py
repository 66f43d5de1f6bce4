"""`GaussianSmooth` / `GaussianSmoothd` (monai/transforms/intensity/array.py:1590-1622, dictionary counterpart) on the
zero-padded separable CUDA filter.  Taps come from `gaussian_1d` (erf-integrated, truncated at 4 sigma, not normalised)."""
from __future__ import annotations

from collections.abc import Hashable, Mapping, Sequence
from typing import Any

import torch

from .. import _kernels as K
from ..data.meta_tensor import rewrap
from ..networks.layers.convutils import gaussian_1d
from .transform import MapTransform, Transform

__all__ = ["GaussianSmooth", "GaussianSmoothd"]


class GaussianSmooth(Transform):
    def __init__(self, sigma: Sequence[float] | float = 1.0, approx: str = "erf") -> None:
        self.sigma = sigma
        self.approx = approx

    def _taps(self, sigma: float) -> torch.Tensor:
        """erf-integrated taps of one axis; a fixed-sigma transform is called once per volume, so the (host) taps are kept"""
        cache = self.__dict__.setdefault("_tap_cache", {})
        key = (sigma, self.approx)
        if key not in cache:
            cache[key] = gaussian_1d(torch.as_tensor(sigma, dtype=torch.float), truncated=4.0, approx=self.approx)
        return cache[key]

    def __call__(self, img: torch.Tensor) -> torch.Tensor:
        if not isinstance(img, torch.Tensor) or not img.is_cuda:
            raise RuntimeError("monai_b200 GaussianSmooth runs on CUDA tensors only (there is no CPU fallback)")
        t = img.as_subclass(torch.Tensor) if type(img) is not torch.Tensor else img
        if t.dtype not in (torch.float16, torch.float32):
            t = t.float()
        nd = t.dim() - 1
        if nd < 1 or nd > 3:
            raise NotImplementedError("GaussianSmooth supports 1-3 spatial dims")
        sig = self.sigma
        if isinstance(sig, Sequence):
            if len(sig) != nd:
                raise ValueError
            sigs = [torch.as_tensor(s, dtype=torch.float) for s in sig]
        else:
            sigs = [torch.as_tensor(sig, dtype=torch.float)] * nd
        taps = [self._taps(float(s)) for s in sigs]
        one = torch.ones(1)
        lift = 3 - nd
        t3 = t.reshape(t.shape[0], *([1] * lift), *t.shape[1:])
        out = K.separable_filter3d(t3.float() if t3.dtype != torch.float32 else t3, [one] * lift + taps)
        out = out.reshape(t.shape).to(t.dtype if t.dtype == torch.float16 else torch.float32)
        return rewrap(out, img)


class GaussianSmoothd(MapTransform):
    def __init__(self, keys, sigma: Sequence[float] | float, approx: str = "erf", allow_missing_keys: bool = False) -> None:
        super().__init__(keys, allow_missing_keys)
        self.converter = GaussianSmooth(sigma, approx=approx)

    def __call__(self, data: Mapping[Hashable, Any]) -> dict:
        d = dict(data)
        for key in self.key_iterator(d):
            d[key] = self.converter(d[key])
        return d
