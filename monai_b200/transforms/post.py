"""`Activations(d)` / `AsDiscrete(d)` (monai/transforms/post/array.py:63-251, dictionary counterparts in post/dictionary.py)
on the channel-wise CUDA kernels of `csrc/post.cu` -- the step that follows the sliding-window inferer in a segmentation
bundle.  Tensors are channel-first without a batch axis, as in the reference.

Same constructor / call arguments and error behaviour; `other` callables run as given (they are user code, not arithmetic of
this package).  `softmax` / `argmax` / one-hot act on dim 0 (the only layout the reference's defaults use); other `dim`
keyword values raise NotImplementedError rather than falling back.
"""
from __future__ import annotations

from collections.abc import Callable, Hashable, Mapping
from typing import Any

import torch

from .. import _kernels as K
from ..data.meta_tensor import rewrap
from .transform import MapTransform, Transform

__all__ = ["Activations", "Activationsd", "AsDiscrete", "AsDiscreted", "Invertd"]


def _cuda_plain(img) -> torch.Tensor:
    if not isinstance(img, torch.Tensor) or not img.is_cuda:
        raise RuntimeError("monai_b200 post transforms run on CUDA tensors only (there is no CPU fallback)")
    t = img.as_subclass(torch.Tensor) if type(img) is not torch.Tensor else img
    return t if t.dtype in (torch.float16, torch.float32) else t.float()


class Activations(Transform):
    def __init__(self, sigmoid: bool = False, softmax: bool = False, other: Callable | None = None, **kwargs) -> None:
        self.sigmoid = sigmoid
        self.softmax = softmax
        self.kwargs = kwargs
        if other is not None and not callable(other):
            raise TypeError(f"other must be None or callable but is {type(other).__name__}.")
        self.other = other

    def __call__(self, img: torch.Tensor, sigmoid: bool | None = None, softmax: bool | None = None, other: Callable | None = None):
        if sigmoid and softmax:
            raise ValueError("Incompatible values: sigmoid=True and softmax=True.")
        if other is not None and not callable(other):
            raise TypeError(f"other must be None or callable but is {type(other).__name__}.")
        if self.kwargs.get("dim", 0) != 0:
            raise NotImplementedError("monai_b200 Activations applies softmax over dim 0 (channel-first) only")
        t = _cuda_plain(img)
        in_dtype = t.dtype
        # the reference converts to float32, applies the activations in that order, and returns the input's dtype
        cur = t
        if sigmoid or self.sigmoid:
            cur = K.channel_post(cur, K.POST_SIGMOID, out_dtype=torch.float32)
        if softmax or self.softmax:
            cur = K.channel_post(cur, K.POST_SOFTMAX, out_dtype=torch.float32)
        act = self.other if other is None else other
        if act is not None:
            cur = act(cur.float())
        out = cur.to(in_dtype) if cur.dtype != in_dtype else cur
        return rewrap(out, img)


class AsDiscrete(Transform):
    def __init__(self, argmax: bool = False, to_onehot: int | None = None, threshold: float | None = None, rounding: str | None = None,
                 **kwargs) -> None:
        self.argmax = argmax
        if isinstance(to_onehot, bool):
            raise ValueError("`to_onehot=True/False` is deprecated, please use `to_onehot=num_classes` instead.")
        self.to_onehot = to_onehot
        self.threshold = threshold
        self.rounding = rounding
        self.kwargs = kwargs

    def __call__(self, img: torch.Tensor, argmax: bool | None = None, to_onehot: int | None = None, threshold: float | None = None,
                 rounding: str | None = None):
        if isinstance(to_onehot, bool):
            raise ValueError("`to_onehot=True/False` is deprecated, please use `to_onehot=num_classes` instead.")
        if self.kwargs.get("dim", 0) != 0 or not self.kwargs.get("keepdim", True):
            raise NotImplementedError("monai_b200 AsDiscrete acts on dim 0 with keepdim=True only")
        out_dtype = self.kwargs.get("dtype", torch.float)
        if out_dtype not in (torch.float32, torch.float16):
            raise NotImplementedError("monai_b200 AsDiscrete returns float32 or float16")
        cur = _cuda_plain(img)
        argmax = self.argmax if argmax is None else argmax
        to_onehot = self.to_onehot if to_onehot is None else to_onehot
        if to_onehot is not None and not isinstance(to_onehot, int):
            raise ValueError(f"the number of classes for One-Hot must be an integer, got {type(to_onehot)}.")
        if argmax and to_onehot is not None:
            cur = K.channel_post(cur, K.POST_ARGMAX, onehot=to_onehot, out_dtype=torch.float32)      # one pass: argmax straight to one-hot
        elif argmax:
            cur = K.channel_post(cur, K.POST_ARGMAX, out_dtype=torch.float32)
        elif to_onehot is not None:
            if cur.shape[0] != 1:
                raise AssertionError("labels should have a channel with length equal to one.")
            cur = K.channel_post(cur, K.POST_ONEHOT, onehot=to_onehot, out_dtype=torch.float32)
        threshold = self.threshold if threshold is None else threshold
        if threshold is not None:
            cur = K.channel_post(cur, K.POST_THRESHOLD, param=float(threshold), out_dtype=torch.float32)
        rounding = self.rounding if rounding is None else rounding
        if rounding is not None:
            if rounding != "torchrounding":
                raise ValueError(f"Unsupported rounding '{rounding}', available options: ['torchrounding'].")
            cur = K.channel_post(cur, K.POST_ROUND, out_dtype=torch.float32)
        out = cur.to(out_dtype) if cur.dtype != out_dtype else cur
        return rewrap(out, img)


class Activationsd(MapTransform):
    def __init__(self, keys, sigmoid: bool = False, softmax: bool = False, other: Callable | None = None, allow_missing_keys: bool = False,
                 **kwargs) -> None:
        super().__init__(keys, allow_missing_keys)
        self.converter = Activations(sigmoid=sigmoid, softmax=softmax, other=other, **kwargs)

    def __call__(self, data: Mapping[Hashable, Any]) -> dict:
        d = dict(data)
        for key in self.key_iterator(d):
            d[key] = self.converter(d[key])
        return d


class AsDiscreted(MapTransform):
    def __init__(self, keys, argmax: bool = False, to_onehot: int | None = None, threshold: float | None = None, rounding: str | None = None,
                 allow_missing_keys: bool = False, **kwargs) -> None:
        super().__init__(keys, allow_missing_keys)
        self.converter = AsDiscrete(argmax=argmax, to_onehot=to_onehot, threshold=threshold, rounding=rounding, **kwargs)

    def __call__(self, data: Mapping[Hashable, Any]) -> dict:
        d = dict(data)
        for key in self.key_iterator(d):
            d[key] = self.converter(d[key])
        return d


class Invertd(MapTransform):
    """Apply the inverse of the pre-processing `transform` to model outputs (monai/transforms/post/dictionary.py, `Invertd`):
    for every key, the prediction takes over the applied operations and the metadata (affine) of `orig_keys`' entry -- the
    pre-processed image the network saw -- and `transform.inverse` maps it back to the original grid (e.g. the inverse of
    Spacingd resamples the logits to the image's native spacing).  `nearest_interp=True` (the reference's default) switches
    the recorded interpolation modes to "nearest" for the inversion; `post_func` runs on the inverted tensor."""

    def __init__(self, keys, transform, orig_keys=None, meta_keys=None, orig_meta_keys=None, meta_key_postfix: str = "meta_dict",
                 nearest_interp: bool | tuple = True, to_tensor: bool | tuple = True, device: Any = None, post_func: Callable | tuple | None = None,
                 allow_missing_keys: bool = False) -> None:
        super().__init__(keys, allow_missing_keys)
        if not hasattr(transform, "inverse"):
            raise ValueError("transform is not invertible, can't invert transform for the data.")
        self.transform = transform
        n = len(self.keys)
        rep = lambda v: tuple(v) if isinstance(v, (list, tuple)) else (v,) * n  # noqa: E731
        self.orig_keys = rep(orig_keys) if orig_keys is not None else self.keys
        self.nearest_interp, self.to_tensor, self.device, self.post_func = rep(nearest_interp), rep(to_tensor), rep(device), rep(post_func)

    @staticmethod
    def _to_nearest(ops: list) -> list:
        """convert_applied_interp_mode(trans_info, mode="nearest", align_corners=None) of monai/transforms/utils.py."""
        import copy

        out = copy.deepcopy(ops)
        for op in out:
            tgt = op.get("extra_info", op)
            if isinstance(tgt, dict) and "mode" in tgt and str(tgt["mode"]).lower() in ("bilinear", "trilinear", "linear", "bicubic", "nearest"):
                tgt["mode"] = "nearest"
                if "align_corners" in tgt:
                    tgt["align_corners"] = "none"
        return out

    def __call__(self, data: Mapping[Hashable, Any]) -> dict:
        import copy

        from ..data.meta_tensor import MetaTensor, is_meta

        d = dict(data)
        for i, key in enumerate(self.key_iterator(d)):
            orig_key = self.orig_keys[i]
            if orig_key not in d or not is_meta(d[orig_key]) or not is_meta(d[key]):
                # the reference skips (with this warning) entries whose transform information is not available: plain-tensor
                # predictions without a trace entry, or a missing pre-processed MetaTensor (post/dictionary.py:677-686)
                import warnings

                warnings.warn(f"transform info of `{orig_key}` is not available or no InvertibleTransform applied.")
                continue
            ops = list(d[orig_key].applied_operations)
            if self.nearest_interp[i]:
                ops = self._to_nearest(ops)
            src = d[key]
            t = src.detach() if isinstance(src, torch.Tensor) else torch.as_tensor(src)
            inputs = MetaTensor(t.as_subclass(torch.Tensor) if type(t) is not torch.Tensor else t)
            inputs.meta = copy.deepcopy(d[orig_key].meta)
            inputs.applied_operations = copy.deepcopy(ops)
            inverted = _inverse_allow_missing(self.transform, {orig_key: inputs})[orig_key]
            if self.device[i] is not None:
                inverted = inverted.to(self.device[i])
            d[key] = self.post_func[i](inverted) if callable(self.post_func[i]) else inverted
        return d


def _inverse_allow_missing(transform, data: dict) -> dict:
    """transform.inverse(data) with `allow_missing_keys` switched on for every dictionary member (allow_missing_keys_mode)."""
    members = list(getattr(transform, "transforms", [transform]))
    saved = [(m, m.allow_missing_keys) for m in members if hasattr(m, "allow_missing_keys")]
    try:
        for m, _ in saved:
            m.allow_missing_keys = True
        return transform.inverse(data)
    finally:
        for m, v in saved:
            m.allow_missing_keys = v
