"""Transform base classes with the reference's randomisation contract (monai/transforms/transform.py:186-226, 356-369)."""
from __future__ import annotations

from collections.abc import Hashable, Mapping
from typing import Any

import numpy as np

MAX_SEED = np.iinfo(np.uint32).max + 1

__all__ = ["Transform", "Randomizable", "RandomizableTransform", "MapTransform", "Compose", "MAX_SEED"]


class Transform:
    def __call__(self, data: Any):
        raise NotImplementedError(f"Subclass {self.__class__.__name__} must implement this method.")


class Randomizable:
    """Own `np.random.RandomState` stream `R`; `set_random_state(seed)` re-seeds it (seed % 2**32)."""

    R: np.random.RandomState = np.random.RandomState()

    def set_random_state(self, seed: int | None = None, state: np.random.RandomState | None = None):
        if seed is not None:
            _seed = np.int64(id(seed) if not isinstance(seed, (int, np.integer)) else seed)
            self.R = np.random.RandomState(int(_seed % MAX_SEED))
            return self
        if state is not None:
            if not isinstance(state, np.random.RandomState):
                raise TypeError(f"state must be None or a np.random.RandomState but is {type(state).__name__}.")
            self.R = state
            return self
        self.R = np.random.RandomState()
        return self

    def randomize(self, data: Any) -> None:
        raise NotImplementedError(f"Subclass {self.__class__.__name__} must implement this method.")


class RandomizableTransform(Randomizable, Transform):
    def __init__(self, prob: float = 1.0, do_transform: bool = True):
        self._do_transform = do_transform
        self.prob = min(max(prob, 0.0), 1.0)

    def randomize(self, data: Any) -> None:
        self._do_transform = self.R.rand() < self.prob


class MapTransform(Transform):
    def __init__(self, keys, allow_missing_keys: bool = False) -> None:
        self.keys: tuple[Hashable, ...] = tuple(keys) if isinstance(keys, (list, tuple)) else (keys,)
        self.allow_missing_keys = allow_missing_keys
        if not self.keys:
            raise ValueError("keys must be non empty.")

    def key_iterator(self, data: Mapping[Hashable, Any]):
        for key in self.keys:
            if key in data:
                yield key
            elif not self.allow_missing_keys:
                raise KeyError(f"Key `{key}` of transform `{self.__class__.__name__}` was missing in the data and allow_missing_keys==False.")


class Compose(Randomizable, Transform):
    """Sequential composition (monai/transforms/compose.py); `set_random_state` seeds every randomizable member from
    this object's stream, as the reference does.

    `lazy=True` runs the lazy-capable members (those with a `lazy` attribute: Spacing(d), RandAffine(d), SpatialResample)
    without resampling -- they record their matrix -- and executes the composed matrix in ONE resample right before the first
    member that is not lazy-capable and at the end (monai/transforms/compose.py:226-270, lazy/functional.py); `lazy=None`
    honours each member's own flag; `overrides` = {key: {"mode": ..., "padding_mode": ...}} (or a flat dict for tensors).
    `inverse` applies the members' inverses in reverse order."""

    def __init__(self, transforms=None, lazy: bool | None = False, overrides: dict | None = None) -> None:
        self.transforms = tuple(transforms) if isinstance(transforms, (list, tuple)) else ((transforms,) if transforms is not None else ())
        self.lazy, self.overrides = lazy, overrides
        self.set_random_state(seed=int(np.random.randint(MAX_SEED, dtype="uint32")) if False else None)

    def set_random_state(self, seed: int | None = None, state: np.random.RandomState | None = None):
        super().set_random_state(seed=seed, state=state)
        for t in self.transforms:
            if isinstance(t, Randomizable):
                t.set_random_state(seed=int(self.R.randint(MAX_SEED, dtype="uint32")))
        return self

    def __call__(self, data, lazy: bool | None = None):
        from .lazy import apply_pending_transforms

        lazy_ = self.lazy if lazy is None else lazy
        if lazy_ is False:
            for t in self.transforms:
                data = t(data)
            return data
        for t in self.transforms:
            capable = hasattr(t, "lazy")
            run_lazy = capable and (lazy_ is True or bool(getattr(t, "lazy", False)))
            if run_lazy:
                data = t(data, lazy=True)
            else:
                # a member that needs real voxels: execute what is pending on the entries it reads first
                data = apply_pending_transforms(data, getattr(t, "keys", None) if isinstance(t, MapTransform) else None, self.overrides)
                data = t(data, lazy=False) if capable else t(data)
        return apply_pending_transforms(data, None, self.overrides)

    def inverse(self, data):
        for t in reversed(self.transforms):
            if hasattr(t, "inverse"):
                data = t.inverse(data)
        return data
