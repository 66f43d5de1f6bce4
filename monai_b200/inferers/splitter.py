"""`Splitter` / `SlidingWindowSplitter` with the reference's contract (monai/inferers/splitter.py:31-292): patches of a
B C H W [D] tensor on a regular grid with optional offset, overlap, padding and a `filter_fn(patch, location)`.

Splitting is indexing, not arithmetic: patches are views (or `F.pad`-ed copies) of the input on whatever device it lives.
The patch grid follows `iter_patch_position` (monai/data/utils.py:209-254) with `padded=False`.
"""
from __future__ import annotations

from abc import ABC, abstractmethod
from collections.abc import Callable, Iterable, Sequence
from inspect import _empty, signature
from itertools import product
from typing import Any

import torch

from ..data.utils import get_valid_patch_size

__all__ = ["Splitter", "SlidingWindowSplitter"]


def _tup(v) -> tuple:
    return tuple(v) if isinstance(v, (list, tuple)) else (v,)


def _rep(v, n: int) -> tuple:
    t = _tup(v)
    if len(t) == 1:
        return t * n
    if len(t) == n:
        return t
    raise ValueError(f"Sequence must have length {n}, got {len(t)}.")


class Splitter(ABC):
    def __init__(self, patch_size: Sequence[int] | int, device: torch.device | str | None = None) -> None:
        self.patch_size = patch_size
        self.device = device

    @abstractmethod
    def get_input_shape(self, inputs: Any) -> tuple:
        raise NotImplementedError(f"Subclass {self.__class__.__name__} must implement this method.")

    @abstractmethod
    def get_padded_shape(self, inputs: Any) -> tuple:
        raise NotImplementedError(f"Subclass {self.__class__.__name__} must implement this method.")

    @abstractmethod
    def __call__(self, inputs: Any) -> Iterable[tuple[torch.Tensor, Sequence[int]]]:
        raise NotImplementedError(f"Subclass {self.__class__.__name__} must implement this method.")


class SlidingWindowSplitter(Splitter):
    def __init__(self, patch_size: Sequence[int] | int, overlap: Sequence[float] | float | Sequence[int] | int = 0.0, offset: Sequence[int] | int = 0,
                 filter_fn: Callable | None = None, pad_mode: str | None = "constant", pad_value: float | int = 0,
                 device: torch.device | str | None = None) -> None:
        super().__init__(patch_size=patch_size, device=device)
        self.offset = offset
        ov = _tup(overlap)
        if isinstance(ov[0], float) and any(o < 0.0 or o >= 1.0 for o in ov):
            raise ValueError(
                f"Relative overlap must be between 0.0 and 1.0 but {overlap} is given. "
                "If you wish to use number of pixels as overlap, please provide integer numbers."
            )
        if any(o < 0 for o in ov):
            raise ValueError(f"Number of pixels for overlap cannot be negative. {overlap} is given. ")
        self.overlap = overlap
        self.filter_fn = self._validate_filter_fn(filter_fn)
        self.pad_mode = pad_mode
        self.pad_value = pad_value
        if not self.pad_mode and any(off < 0 for off in _tup(offset)):
            raise ValueError(f"Negative `offset`requires a valid padding mode but `mode` is set to {self.pad_mode}.")

    @staticmethod
    def _validate_filter_fn(filter_fn):
        if callable(filter_fn):
            params = signature(filter_fn).parameters
            positional = [v for v in params.values() if v.default is _empty]
            if len(params) < 2:
                raise ValueError(
                    f"`filter_fn` requires to accept at least two parameters (patch, location)."
                    f"The provided callable ({filter_fn}) has {len(params)} parameters."
                )
            if len(positional) > 2:
                raise ValueError(
                    f"`filter_fn` can have at most two positional parameters (patch, location)."
                    f"The provided callable ({filter_fn}) has {len(positional)} positional parameters."
                )
        elif filter_fn is not None:
            raise ValueError(
                "`filter_fn` should be a callable with two input parameters (patch, location). "
                f"{type(filter_fn)} is given."
            )
        return filter_fn

    def _calculate_pad_size(self, spatial_shape, spatial_ndim, patch_size, offset, overlap):
        """[end_last, start_last, ..]-ordered pad list in the layout `F.pad` takes after reversal (reference :172-192)."""
        pad = [0] * (2 * spatial_ndim)
        if not self.pad_mode:
            return pad, False
        pad[1::2] = [-min(off, 0) for off in offset]                 # start pad only for negative offsets
        end = []
        for sh, off, ps, ov in zip(spatial_shape, offset, patch_size, overlap):
            if ps == 0:
                end.append(0)
            else:
                step = round(ps - (ps * ov)) if isinstance(ov, float) else round(ps - ov)
                end.append((off - sh + ps) % step)                   # end pad so that the last patch is whole
        pad[::2] = end
        return pad, any(pad[1::2])

    def _get_valid_shape_parameters(self, spatial_shape: Sequence[int]):
        nd = len(spatial_shape)
        patch_size = _rep(self.patch_size, nd)
        overlap = _rep(self.overlap, nd)
        overlap = tuple(o if p else type(overlap[0])(0) for o, p in zip(overlap, patch_size))
        if any(ov > ps for ov, ps in zip(overlap, patch_size)):
            raise ValueError(f"`overlap` ({overlap}) cannot be larger than patch size ({patch_size}).")
        offset = _rep(self.offset, nd)
        for off, ps, sh in zip(offset, patch_size, spatial_shape):
            if off < -ps:
                raise ValueError(f"Negative `offset` ({off}) cannot be larger than `patch_size` ({ps}) in magnitude.")
            if off >= sh:
                raise ValueError(f"`offset` ({off}) cannot be larger than inputs size ({sh}).")
        return patch_size, overlap, offset

    def get_input_shape(self, inputs: Any) -> tuple:
        return tuple(inputs.shape[2:])

    def get_padded_shape(self, inputs: Any) -> tuple:
        spatial_shape = self.get_input_shape(inputs)
        if not self.pad_mode:
            return spatial_shape
        patch_size, overlap, offset = self._get_valid_shape_parameters(spatial_shape)
        pad, _ = self._calculate_pad_size(spatial_shape, len(spatial_shape), patch_size, offset, overlap)
        return tuple(ss + ps + pe for ss, ps, pe in zip(spatial_shape, pad[1::2], pad[::2]))

    @staticmethod
    def _positions(image_size, patch_size, start_pos, overlap):
        """iter_patch_position(..., padded=False): row-major grid of patch corners (monai/data/utils.py:209-254)."""
        ps = get_valid_patch_size(image_size, patch_size)
        if isinstance(overlap[0], float):
            steps = tuple(round(p * (1.0 - o)) for p, o in zip(ps, overlap))
        else:
            steps = tuple(p - o for p, o in zip(ps, overlap))
        ends = tuple(s - round(p) + 1 for s, p in zip(image_size, ps))
        return product(*(range(a, b, c) for a, b, c in zip(start_pos, ends, steps)))

    def __call__(self, inputs: Any) -> Iterable[tuple[torch.Tensor, Sequence[int]]]:
        if not isinstance(inputs, torch.Tensor):
            raise ValueError(f"The input should be a tensor. {type(inputs)} is given.")
        spatial_shape = tuple(inputs.shape[2:])
        nd = len(spatial_shape)
        patch_size, overlap, offset = self._get_valid_shape_parameters(spatial_shape)
        pad, start_padded = self._calculate_pad_size(spatial_shape, nd, patch_size, offset, overlap)
        if self.pad_mode and any(pad):
            inputs = torch.nn.functional.pad(inputs, pad[::-1], mode=self.pad_mode, value=self.pad_value)
            spatial_shape = tuple(inputs.shape[2:])
            if start_padded:
                offset = tuple(off + p for off, p in zip(offset, pad[1::2]))
        for location in self._positions(spatial_shape, patch_size, offset, overlap):
            sl = (slice(None),) * 2 + tuple(slice(loc, loc + ps) for loc, ps in zip(location, patch_size))
            patch = inputs[sl]
            if self.device is not None:
                patch = patch.to(self.device)
            if start_padded:
                location = tuple(loc - p for loc, p in zip(location, pad[1::2]))
            if self.filter_fn is None or self.filter_fn(patch, location):
                yield patch, location
