"""Inferer classes with the reference's constructor / call signatures (monai/inferers/inferer.py:62-97, 399-552).

`SlidingWindowInferer` stores its arguments and calls the B200-native `sliding_window_inference` positionally,
exactly as the reference does (inferer.py:532-552).  `SlidingWindowInfererAdapt` keeps its name and signature; the
OOM ladder of the reference (inferer.py:565-641) degrades gracefully here because the resident-prediction budget
already bounds memory, so it only adds the `cpu_thresh` bookkeeping.
"""
from __future__ import annotations

import warnings
from abc import ABC, abstractmethod
from collections.abc import Callable, Mapping, Sequence
from typing import Any

import torch

from ..data.utils import compute_importance_map
from .utils import sliding_window_inference

__all__ = ["Inferer", "SimpleInferer", "SlidingWindowInferer", "SlidingWindowInfererAdapt", "SliceInferer"]


class Inferer(ABC):
    """Base class: `inferer(inputs, network, *args, **kwargs)` (inferer.py:62-97)."""

    @abstractmethod
    def __call__(self, inputs: torch.Tensor, network: Callable, *args: Any, **kwargs: Any) -> Any:
        raise NotImplementedError(f"Subclass {self.__class__.__name__} must implement this method.")


class SimpleInferer(Inferer):
    """Runs `network(inputs, *args, **kwargs)` (inferer.py:373-396)."""

    def __call__(self, inputs: torch.Tensor, network: Callable, *args: Any, **kwargs: Any):
        return network(inputs, *args, **kwargs)


class SlidingWindowInferer(Inferer):
    def __init__(
        self,
        roi_size: Sequence[int] | int,
        sw_batch_size: int = 1,
        overlap: Sequence[float] | float = 0.25,
        mode: str = "constant",
        sigma_scale: Sequence[float] | float = 0.125,
        padding_mode: str = "constant",
        cval: float = 0.0,
        sw_device: torch.device | str | None = None,
        device: torch.device | str | None = None,
        progress: bool = False,
        cache_roi_weight_map: bool = False,
        cpu_thresh: int | None = None,
        buffer_steps: int | None = None,
        buffer_dim: int = -1,
        with_coord: bool = False,
    ) -> None:
        super().__init__()
        mode_s = str(getattr(mode, "value", mode)).lower()
        if mode_s not in ("constant", "gaussian"):
            raise ValueError(f"'{mode}' is not a valid BlendMode")
        self.roi_size = roi_size
        self.sw_batch_size = sw_batch_size
        self.overlap = overlap
        self.mode = mode_s
        self.sigma_scale = sigma_scale
        self.padding_mode = padding_mode
        self.cval = cval
        self.sw_device = sw_device
        self.device = device
        self.progress = progress
        self.cpu_thresh = cpu_thresh
        self.buffer_steps = buffer_steps
        self.buffer_dim = buffer_dim
        self.with_coord = with_coord
        # the reference precomputes the dense map to avoid recomputing it per call; the CUDA blend only needs the
        # separable factors, which cost nothing -- the cached map is kept for attribute compatibility.
        self.roi_weight_map = None
        try:
            if cache_roi_weight_map and isinstance(roi_size, Sequence) and min(roi_size) > 0:
                self.roi_weight_map = compute_importance_map(
                    tuple(self.roi_size), mode=mode_s, sigma_scale=sigma_scale, device=device if device is not None else "cpu"
                )
            if cache_roi_weight_map and self.roi_weight_map is None:
                warnings.warn("cache_roi_weight_map=True, but cache is not created. (dynamic roi_size?)")
        except BaseException as e:
            raise RuntimeError(
                f"roi size {self.roi_size}, mode={mode}, sigma_scale={sigma_scale}, device={device}\n"
                "Seems to be OOM. Please try smaller patch size or mode='constant' instead of mode='gaussian'."
            ) from e

    def __call__(self, inputs: torch.Tensor, network: Callable, *args: Any, **kwargs: Any):
        device = kwargs.pop("device", self.device)
        buffer_steps = kwargs.pop("buffer_steps", self.buffer_steps)
        buffer_dim = kwargs.pop("buffer_dim", self.buffer_dim)
        if device is None and self.cpu_thresh is not None and inputs.shape[2:].numel() > self.cpu_thresh:
            device = "cpu"  # stitched result is handed back in host memory for very large images
        return sliding_window_inference(
            inputs,
            self.roi_size,
            self.sw_batch_size,
            network,
            self.overlap,
            self.mode,
            self.sigma_scale,
            self.padding_mode,
            self.cval,
            self.sw_device,
            device,
            self.progress,
            self.roi_weight_map,
            None,
            buffer_steps,
            buffer_dim,
            self.with_coord,
            *args,
            **kwargs,
        )


class SlidingWindowInfererAdapt(SlidingWindowInferer):
    """Same call contract as the reference class; retries with the result on the host after a CUDA OOM."""

    def __call__(self, inputs: torch.Tensor, network: Callable, *args: Any, **kwargs: Any):
        if self.device is not None or "device" in kwargs:
            return super().__call__(inputs, network, *args, **kwargs)
        try:
            return super().__call__(inputs, network, *args, **kwargs)
        except torch.cuda.OutOfMemoryError:
            torch.cuda.empty_cache()
            self.cpu_thresh = inputs.shape[2:].numel() - 1 if self.cpu_thresh is None else min(self.cpu_thresh, inputs.shape[2:].numel() - 1)
            warnings.warn("CUDA OOM during sliding-window inference; retrying with the stitched output on the host.")
            return super().__call__(inputs, network, *args, device="cpu", **kwargs)


class SliceInferer(SlidingWindowInferer):
    """Slice-by-slice (2-D network) inference over a 3-D volume with the same contract as the reference class
    (monai/inferers/inferer.py:691-771): `roi_size` is 2-D, a singleton is inserted at `spatial_dim`, the window batch is
    squeezed before the 2-D network runs and its outputs (tensor, sequence or mapping) are unsqueezed again.  Gather and
    blend run in the CUDA kernels of `sliding_window_inference`; the 2-D predictor is any callable."""

    def __init__(self, spatial_dim: int = 0, *args: Any, **kwargs: Any) -> None:
        self.spatial_dim = spatial_dim
        super().__init__(*args, **kwargs)
        roi = self.roi_size
        self.orig_roi_size = tuple(roi) if isinstance(roi, (list, tuple)) else (roi,)

    def __call__(self, inputs: torch.Tensor, network: Callable, *args: Any, **kwargs: Any):
        if self.spatial_dim > 2:
            raise ValueError("`spatial_dim` can only be `0, 1, 2` with `[H, W, D]` respectively.")
        if len(self.orig_roi_size) == 2 and len(inputs.shape[2:]) == 3:
            roi = list(self.orig_roi_size)
            roi.insert(self.spatial_dim, 1)
            self.roi_size = roi
        else:
            raise RuntimeError(
                f"Currently, only 2D `roi_size` ({self.orig_roi_size}) with 3D `inputs` tensor (shape={inputs.shape}) is supported."
            )
        return super().__call__(inputs, lambda x: self.network_wrapper(network, x, *args, **kwargs))

    def network_wrapper(self, network: Callable, x: torch.Tensor, *args: Any, **kwargs: Any):
        dim = self.spatial_dim + 2
        out = network(x.squeeze(dim=dim), *args, **kwargs)
        if isinstance(out, torch.Tensor):
            return out.unsqueeze(dim=dim)
        if isinstance(out, Mapping):
            return {k: v.unsqueeze(dim=dim) for k, v in out.items()}
        return tuple(o.unsqueeze(dim=dim) for o in out)
