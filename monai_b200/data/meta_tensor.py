"""Minimal MetaTensor: a torch.Tensor subclass carrying `.affine` (4x4 float64), `.meta` and `.applied_operations`.

The reference's MetaTensor (monai/data/meta_tensor.py:52) is part of the drop-in *boundary*, not of the hot path: when
MONAI itself is importable its MetaTensor can be passed to every monai_b200 transform / inferer (they duck-type on
`.affine`, `.meta`, `copy_meta_from`).  This class exists so the package is usable stand-alone (e.g. on the GPU box).
"""
from __future__ import annotations

import copy
from typing import Any

import numpy as np
import torch

__all__ = ["MetaTensor", "is_meta", "get_affine", "rewrap"]


class MetaTensor(torch.Tensor):
    @staticmethod
    def __new__(cls, x, affine=None, meta: dict | None = None, applied_operations: list | None = None, *args, **kwargs):
        t = torch.as_tensor(x, *args, **kwargs)
        return t.as_subclass(cls)

    def __init__(self, x, affine=None, meta: dict | None = None, applied_operations: list | None = None, *args, **kwargs) -> None:
        self.meta: dict[str, Any] = dict(meta) if meta is not None else dict(getattr(x, "meta", {}) or {})
        self.applied_operations: list = list(applied_operations) if applied_operations is not None else list(getattr(x, "applied_operations", []) or [])
        self.pending_operations: list = list(getattr(x, "pending_operations", []) or [])   # lazy resampling (transforms/lazy.py)
        if affine is not None:
            self.affine = affine
        elif "affine" not in self.meta:
            self.affine = torch.eye(4, dtype=torch.float64)

    @property
    def affine(self) -> torch.Tensor:
        return self.meta.get("affine", torch.eye(4, dtype=torch.float64))

    @affine.setter
    def affine(self, d) -> None:
        self.meta["affine"] = torch.as_tensor(d, dtype=torch.float64, device="cpu")

    def as_tensor(self) -> torch.Tensor:
        return self.as_subclass(torch.Tensor)

    def copy_meta_from(self, other, copy_attr: bool = True):
        self.meta = _copy_tree(other.meta) if copy_attr else dict(other.meta)
        self.applied_operations = _copy_tree(other.applied_operations) if copy_attr else list(getattr(other, "applied_operations", []))
        self.pending_operations = _copy_tree(getattr(other, "pending_operations", [])) if copy_attr else list(getattr(other, "pending_operations", []))
        return self

    # ---- lazy resampling bookkeeping (monai/data/meta_tensor.py:480-507, meta_obj.py push/pop/clear) ----------------------
    def push_pending_operation(self, info: dict) -> None:
        self.pending_operations.append(info)

    def pop_pending_operation(self) -> dict:
        return self.pending_operations.pop()

    def clear_pending_operations(self) -> None:
        self.pending_operations = []

    def push_applied_operation(self, info: dict) -> None:
        self.applied_operations.append(info)

    def pop_applied_operation(self) -> dict:
        return self.applied_operations.pop()

    def peek_pending_shape(self):
        """spatial shape as if all the pending operations were executed"""
        res = self.pending_operations[-1].get("lazy_shape", None) if self.pending_operations else None
        return tuple(int(s) for s in self.shape[1:]) if res is None else tuple(int(s) for s in res)

    def peek_pending_affine(self):
        res = torch.as_tensor(self.affine, dtype=torch.float64)
        r = len(res) - 1
        for p in self.pending_operations:
            m = p.get("lazy_affine")
            if m is None:
                continue
            m = torch.as_tensor(m, dtype=torch.float64)
            full = torch.eye(r + 1, dtype=torch.float64)
            k = min(r, m.shape[0] - 1)
            full[:k, :k], full[:k, -1] = m[:k, :k], m[:k, -1]
            res = res @ full
        return res

    def peek_pending_rank(self) -> int:
        a = self.pending_operations[-1].get("lazy_affine", None) if self.pending_operations else self.affine
        return 1 if a is None else int(max(1, len(a) - 1))

    @classmethod
    def __torch_function__(cls, func, types, args=(), kwargs=None):
        # metadata does not propagate through arbitrary torch ops here: results are plain tensors unless re-wrapped
        ret = super().__torch_function__(func, types, args, kwargs or {})
        return ret

    def __repr__(self, **kw) -> str:  # pragma: no cover
        return f"MetaTensor({self.as_tensor()!r}, affine={self.affine.tolist()})"


_IMMUTABLE = (int, float, str, bool, bytes, type(None), complex)


def _copy_tree(x):
    """Deep copy of the metadata containers (dict / list / tuple of scalars, arrays and small tensors).  Same result as
    copy.deepcopy for these types, without its memo machinery and torch's storage-level __deepcopy__ -- the affine is copied three
    times per transform, and the generic deepcopy was a fifth of the host time of the C4 pipeline."""
    if isinstance(x, _IMMUTABLE):
        return x
    if isinstance(x, torch.Tensor):
        return x.detach().clone() if type(x) is torch.Tensor else copy.deepcopy(x)
    if isinstance(x, dict):
        return {k: _copy_tree(v) for k, v in x.items()} if type(x) is dict else copy.deepcopy(x)
    if isinstance(x, list):
        return [_copy_tree(v) for v in x] if type(x) is list else copy.deepcopy(x)
    if isinstance(x, tuple):
        return tuple(_copy_tree(v) for v in x) if type(x) is tuple else copy.deepcopy(x)
    if type(x) is np.ndarray and x.dtype != object:
        return x.copy()
    return copy.deepcopy(x)   # object arrays, numpy scalars, user objects


def is_meta(x) -> bool:
    return hasattr(x, "meta") and hasattr(x, "copy_meta_from")


def get_affine(x):
    """4x4 (or (r+1)x(r+1)) float64 affine of a MetaTensor-like object, or None for plain tensors."""
    if not is_meta(x):
        return None
    a = x.peek_pending_affine() if hasattr(x, "peek_pending_affine") else x.affine
    return torch.as_tensor(a, dtype=torch.float64).cpu()


def rewrap(out: torch.Tensor, like, affine=None, applied=None):
    """Wrap `out` with the metadata of `like` (same class as `like`), optionally replacing the affine."""
    if not is_meta(like):
        return out
    if type(like) is MetaTensor:
        # same result as MetaTensor(out) + copy_meta_from, without building the default affine that copy_meta_from replaces at once
        w = (out.as_subclass(torch.Tensor) if type(out) is not torch.Tensor else out).as_subclass(MetaTensor)
        w.copy_meta_from(like, copy_attr=True)
    else:
        w = type(like)(out)
        w.copy_meta_from(like, copy_attr=True)
    if affine is not None:
        w.affine = torch.as_tensor(affine, dtype=torch.float64)
    if applied is not None:
        w.applied_operations = list(w.applied_operations) + [applied]
    return w
