"""`Merger` / `AvgMerger` with the reference's contract (monai/inferers/merger.py:41-205): patch outputs are summed into a
float32 buffer with a per-element sample count and averaged on `finalize()`.  The accumulate and the division run in the
CUDA kernels `b200_patch_accumulate` / `b200_patch_finalize` (csrc/post.cu); buffers therefore live on a CUDA device.
"""
from __future__ import annotations

import threading
from abc import ABC, abstractmethod
from collections.abc import Sequence
from typing import Any

import torch

from .. import _kernels as K

__all__ = ["Merger", "AvgMerger"]


class Merger(ABC):
    def __init__(self, merged_shape: Sequence[int], cropped_shape: Sequence[int] | None = None, device: torch.device | str | None = None) -> None:
        if merged_shape is None:
            raise ValueError("Argument `merged_shape` must be provided")
        self.merged_shape: tuple[int, ...] = tuple(merged_shape)
        self.cropped_shape: tuple[int, ...] = self.merged_shape if cropped_shape is None else tuple(cropped_shape)
        self.device = device
        self.is_finalized = False

    @abstractmethod
    def aggregate(self, values: torch.Tensor, location: Sequence[int]) -> None:
        raise NotImplementedError(f"Subclass {self.__class__.__name__} must implement this method.")

    @abstractmethod
    def finalize(self) -> Any:
        raise NotImplementedError(f"Subclass {self.__class__.__name__} must implement this method.")


class AvgMerger(Merger):
    def __init__(self, merged_shape: Sequence[int], cropped_shape: Sequence[int] | None = None, value_dtype: torch.dtype = torch.float32,
                 count_dtype: torch.dtype = torch.uint8, device: torch.device | str = "cuda") -> None:
        super().__init__(merged_shape=merged_shape, cropped_shape=cropped_shape, device=device)
        if not self.merged_shape:
            raise ValueError(f"`merged_shape` must be provided for `AvgMerger`. {self.merged_shape} is give.")
        if torch.device(device).type != "cuda":
            raise RuntimeError("monai_b200 AvgMerger aggregates on a CUDA device (there is no CPU fallback)")
        if value_dtype != torch.float32 or count_dtype not in (torch.uint8, torch.int32):
            raise NotImplementedError("monai_b200 AvgMerger keeps float32 values and uint8 / int32 counts")
        self.value_dtype = value_dtype
        self.count_dtype = count_dtype
        self.values = torch.zeros(self.merged_shape, dtype=self.value_dtype, device=self.device)
        self.counts = torch.zeros(self.merged_shape, dtype=self.count_dtype, device=self.device)
        self._lock = threading.Lock()

    def aggregate(self, values: torch.Tensor, location: Sequence[int]) -> None:
        if self.is_finalized:
            raise ValueError("`AvgMerger` is already finalized. Please instantiate a new object to aggregate.")
        with self._lock:
            K.patch_accumulate(values.detach(), self.values, self.counts, location)

    def finalize(self) -> torch.Tensor:
        if not self.is_finalized:
            K.patch_finalize(self.values, self.counts)
            self.values = self.values[tuple(slice(0, end) for end in self.cropped_shape)]
            self.is_finalized = True
        return self.values

    def get_output(self) -> torch.Tensor:
        return self.finalize()

    def get_values(self) -> torch.Tensor:
        return self.values

    def get_counts(self) -> torch.Tensor:
        return self.counts
