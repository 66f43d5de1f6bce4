"""Spatial transforms with the reference's operator surface (monai/transforms/spatial/array.py, dictionary.py):
`SpatialResample`, `Spacing`, `Spacingd`, `Resample`, `AffineGrid`, `RandAffineGrid`, `Affine`, `RandAffine`, `RandAffined`.

Every resampling step is ONE launch of `b200_resample_affine`: sampling coordinates are an affine function of the
output voxel index, so the dense coordinate grid the reference materialises (F.affine_grid, create_grid @ affine;
3-4 x voxels x 8 bytes) and its float64 image copy never exist.  Host-side algebra (output shape, affines, RNG draw
order) follows the reference exactly (see transforms/utils.py); metadata (`.affine`, `applied_operations`) is kept.
Outputs are float32 like the reference (spatial/array.py:2116, functional.py:183).
"""
from __future__ import annotations

from collections.abc import Hashable, Mapping, Sequence
from typing import Any

import numpy as np
import torch

from .. import _kernels as K
from .. import _lib as L
from ..data.meta_tensor import get_affine, is_meta, rewrap
from . import utils as U
from .transform import MapTransform, Randomizable, RandomizableTransform, Transform

__all__ = ["SpatialResample", "Spacing", "Spacingd", "Resample", "AffineGrid", "RandAffineGrid", "Affine", "RandAffine", "RandAffined"]

_INTERP = {"bilinear": L.INTERP_LINEAR, "linear": L.INTERP_LINEAR, "trilinear": L.INTERP_LINEAR, "nearest": L.INTERP_NEAREST, 1: L.INTERP_LINEAR, 0: L.INTERP_NEAREST}
_PAD = {"zeros": L.PAD_ZEROS, "border": L.PAD_BORDER, "reflection": L.PAD_REFLECTION, "constant": L.PAD_ZEROS, "nearest": L.PAD_BORDER, "reflect": L.PAD_REFLECTION}


def _mode(v) -> int:
    key = getattr(v, "value", v)
    key = key.lower() if isinstance(key, str) else key
    if key not in _INTERP:
        raise NotImplementedError(f"monai_b200 resampling implements nearest and (bi/tri)linear interpolation, got {v!r}")
    return _INTERP[key]


def _pad(v) -> int:
    key = str(getattr(v, "value", v)).lower()
    if key not in _PAD:
        raise NotImplementedError(f"monai_b200 resampling implements zeros / border / reflection padding, got {v!r}")
    return _PAD[key]


def _fall_back(user, default) -> tuple[int, ...]:
    if user is None:
        return tuple(int(d) for d in default)
    if not isinstance(user, (list, tuple, np.ndarray, torch.Tensor)):
        user = (user,) * len(default)
    return tuple(int(u) if (u is not None and u > 0) else int(d) for u, d in zip(user, default))


def _pop_applied(data, cls_name: str) -> dict:
    """Pop the most recent applied operation of `data` and check that it was recorded by `cls_name` (inverse.py:300-360)."""
    if not is_meta(data) or not getattr(data, "applied_operations", None):
        raise RuntimeError(f"{cls_name}.inverse needs a MetaTensor that went through the forward transform (no applied operations found)")
    op = data.applied_operations[-1]
    if op.get("class") != cls_name:
        raise RuntimeError(f"Error {cls_name} getting the most recently applied invertible transform {op.get('class')}.")
    data.applied_operations.pop()
    return op


def _resample(img: torch.Tensor, mat: np.ndarray, r: int, out_shape: Sequence[int], mode, padding_mode, align_corners: bool) -> torch.Tensor:
    """img [C, *spatial(r)] (CUDA) -> float32 [C, *out_shape] through the 3-D kernel (leading singleton axes for r < 3)."""
    if not img.is_cuda:
        raise RuntimeError("monai_b200 spatial transforms run on CUDA tensors only (there is no CPU fallback)")
    t = img.as_subclass(torch.Tensor) if type(img) is not torch.Tensor else img
    if t.dtype not in (torch.float16, torch.float32):
        t = t.float()
    lift = 3 - r
    extra = tuple(t.shape[1 + r:])  # additional (non-spatial) trailing dims are folded into channels
    if extra:
        t = t.reshape(t.shape[0], *t.shape[1 : 1 + r], -1).movedim(-1, 1).reshape(-1, *t.shape[1 : 1 + r])
    t3 = t.reshape(t.shape[0], *([1] * lift), *t.shape[1:])
    out3 = K.resample_affine(t3, (1,) * lift + tuple(int(s) for s in out_shape), U.lift_to_3d(mat, r).reshape(-1), _mode(mode), _pad(padding_mode), bool(align_corners))
    out = out3.reshape(out3.shape[0], *out_shape)
    if extra:
        out = out.reshape(img.shape[0], -1, *out_shape).movedim(1, -1).reshape(img.shape[0], *out_shape, *extra)
    return out


class SpatialResample(Transform):
    """Resample to a destination affine / shape (spatial/array.py:122-253; functional.py:68-184)."""

    def __init__(self, mode="bilinear", padding_mode="border", align_corners: bool = False, dtype=np.float64, lazy: bool = False):
        self.mode, self.padding_mode, self.align_corners, self.dtype, self.lazy = mode, padding_mode, align_corners, dtype, lazy
        self._trace = True
        self._algebra_cache: dict = {}

    def inverse(self, data):
        """Resample back to the grid the forward call started from (spatial/array.py:238-253)."""
        op = _pop_applied(data, "SpatialResample")
        ex = op["extra_info"]
        inv = SpatialResample(mode=ex["mode"], padding_mode=ex["padding_mode"], align_corners=bool(ex["align_corners"]) if ex["align_corners"] not in (None, "none") else False, dtype=self.dtype)
        inv._trace = False
        return inv(data, dst_affine=ex["src_affine"], spatial_size=op["orig_size"])

    def __call__(self, img, dst_affine=None, spatial_size=None, mode=None, padding_mode=None, align_corners=None, dtype=None, lazy=None):
        align = self.align_corners if align_corners is None else align_corners
        lazy_ = self.lazy if lazy is None else lazy
        src_affine_full = get_affine(img)
        if src_affine_full is None:
            src_affine_full = torch.eye(4, dtype=torch.float64)
        original_shape = tuple(img.peek_pending_shape()) + tuple(img.shape[1 + len(img.peek_pending_shape()):]) if is_meta(img) and getattr(img, "pending_operations", None) else tuple(img.shape[1:])
        rank = min(len(img.shape) - 1, src_affine_full.shape[0] - 1, 3)
        if (not isinstance(spatial_size, int) or spatial_size != -1) and spatial_size is not None:
            rank = min(len(tuple(spatial_size)), 3)
        # The affine algebra below depends on (source affine, destination affine, shapes) only: a dataset resampled to one spacing
        # repeats it volume after volume, and at 1-2 us per numpy call it -- not the kernel -- set the pace of the C4 pipeline.
        src_np = src_affine_full.numpy() if isinstance(src_affine_full, torch.Tensor) else np.asarray(src_affine_full)
        dst_np = None if dst_affine is None else (dst_affine.detach().cpu().numpy() if isinstance(dst_affine, torch.Tensor) else np.asarray(dst_affine))
        size_key = spatial_size if (spatial_size is None or isinstance(spatial_size, int)) else tuple(int(v) for v in spatial_size)
        key = (src_np.shape, src_np.astype(np.float64).tobytes(), None if dst_np is None else (dst_np.shape, dst_np.astype(np.float64).tobytes()),
               original_shape, size_key, rank)
        hit = self._algebra_cache.get(key)
        if hit is None:
            src_affine = U.to_affine_nd(rank, src_affine_full)
            dst = U.to_affine_nd(rank, dst_affine) if dst_affine is not None else src_affine
            in_size = np.asarray(original_shape[:rank])
            if isinstance(spatial_size, int) and spatial_size == -1:
                spatial_size = in_size
            elif spatial_size is None and rank > 1:
                spatial_size, _ = U.compute_shape_offset(in_size, src_affine, dst)
            sp = np.asarray([int(s) if s >= 0 else int(d) for s, d in zip(tuple(spatial_size)[:rank], in_size)])
            try:
                xform = np.eye(rank + 1) if rank < 2 else np.linalg.solve(src_affine, dst)
            except np.linalg.LinAlgError as e:
                raise ValueError(f"src affine is not invertible {src_affine}, {dst}.") from e
            xform = U.to_affine_nd(rank, xform)
            unchanged = bool((np.allclose(src_affine, dst, atol=U.AFFINE_TOL) and np.allclose(sp, in_size)) or (
                np.allclose(xform, np.eye(len(xform)), atol=U.AFFINE_TOL) and np.allclose(sp, in_size)
            ))
            if len(self._algebra_cache) >= 64:
                self._algebra_cache.clear()
            self._algebra_cache[key] = (src_affine, sp, xform, unchanged)
        else:
            src_affine, sp, xform, unchanged = hit
        src_affine, sp, xform = src_affine.copy(), sp.copy(), xform.copy()    # callers keep / edit these (applied_operations, lazy)
        info = {"class": type(self).__name__, "orig_size": original_shape, "extra_info": {"src_affine": src_affine, "align_corners": align,
                "mode": getattr(mode or self.mode, "value", mode or self.mode), "padding_mode": getattr(padding_mode or self.padding_mode, "value", padding_mode or self.padding_mode)}}
        if not self._trace:
            info = None
        if lazy_:   # record the operation, leave the voxels alone (spatial/functional.py:138-150)
            from .lazy import push_pending

            return push_pending(img, info or {"class": type(self).__name__}, xform, sp)
        new_affine = None
        if is_meta(img):
            # track_transform_meta: affine <- affine @ xform (the voxel->world map of the resampled grid)
            new_affine = U.to_affine_nd(len(src_affine_full) - 1, src_affine_full.numpy().copy())
            full = np.eye(len(src_affine_full))
            full[:rank, :rank], full[:rank, -1] = xform[:rank, :rank], xform[:rank, -1]
            new_affine = src_affine_full.numpy() @ full if not unchanged else src_affine_full.numpy()
        if unchanged:
            out = (img.as_subclass(torch.Tensor) if type(img) is not torch.Tensor else img).to(torch.float32)
            return rewrap(out, img, new_affine, info)
        mkey = (key, bool(align))
        mat = self._algebra_cache.get(mkey)
        if mat is None:
            mat = U.sample_matrix_from_xform(xform, original_shape[:rank], sp, align)
            self._algebra_cache[mkey] = mat
        out = _resample(img, mat, rank, tuple(int(s) for s in sp), mode or self.mode, padding_mode or self.padding_mode, align)
        return rewrap(out, img, new_affine, info)


class Spacing(Transform):
    """Resample to the given voxel spacing (spatial/array.py:338-546)."""

    def __init__(self, pixdim, diagonal: bool = False, mode="bilinear", padding_mode="border", align_corners: bool = False, dtype=np.float64,
                 scale_extent: bool = False, recompute_affine: bool = False, min_pixdim=None, max_pixdim=None, lazy: bool = False):
        self.pixdim = np.array(pixdim if isinstance(pixdim, (list, tuple, np.ndarray)) else (pixdim,), dtype=np.float64)
        self.min_pixdim = np.array(min_pixdim if isinstance(min_pixdim, (list, tuple, np.ndarray)) else (min_pixdim,), dtype=np.float64) if min_pixdim is not None else np.array([np.nan])
        self.max_pixdim = np.array(max_pixdim if isinstance(max_pixdim, (list, tuple, np.ndarray)) else (max_pixdim,), dtype=np.float64) if max_pixdim is not None else np.array([np.nan])
        if min_pixdim is None:
            self.min_pixdim = np.full(len(self.pixdim), np.nan)
        if max_pixdim is None:
            self.max_pixdim = np.full(len(self.pixdim), np.nan)
        self.diagonal, self.scale_extent, self.recompute_affine = diagonal, scale_extent, recompute_affine
        for mn, mx in zip(self.min_pixdim, self.max_pixdim):
            if (not np.isnan(mn)) and (not np.isnan(mx)) and ((mx < mn) or (mn < 0)):
                raise ValueError(f"min_pixdim {self.min_pixdim} must be positive, smaller than max {self.max_pixdim}.")
        self.sp_resample = SpatialResample(mode=mode, padding_mode=padding_mode, align_corners=align_corners, dtype=dtype, lazy=lazy)
        self._algebra_cache: dict = {}     # (input affine, shape, scale_extent) -> (affine, new affine, output shape); see SpatialResample

    def inverse(self, data):
        """spatial/array.py:545-546: the inverse of the SpatialResample this transform ran."""
        return self.sp_resample.inverse(data)

    def __call__(self, data_array, mode=None, padding_mode=None, align_corners=None, dtype=None, scale_extent=None, output_spatial_shape=None, lazy=None):
        original_shape = tuple(data_array.peek_pending_shape()) if is_meta(data_array) and getattr(data_array, "pending_operations", None) else tuple(data_array.shape[1:])
        sr = len(original_shape)
        if sr <= 0:
            raise ValueError(f"data_array must have at least one spatial dimension, got {original_shape}.")
        input_affine = get_affine(data_array)
        if input_affine is None:
            import warnings

            warnings.warn("`data_array` is not of type MetaTensor, assuming affine to be identity.")
            input_affine = np.eye(sr + 1, dtype=np.float64)
        scale_extent = self.scale_extent if scale_extent is None else scale_extent
        in_np = input_affine.numpy() if isinstance(input_affine, torch.Tensor) else np.asarray(input_affine)
        ckey = (in_np.shape, in_np.astype(np.float64).tobytes(), original_shape, bool(scale_extent), self.pixdim.tobytes(), self.min_pixdim.tobytes(),
                self.max_pixdim.tobytes(), bool(self.diagonal))
        chit = self._algebra_cache.get(ckey)
        if chit is not None:
            affine_, new_affine, output_shape = chit[0].copy(), chit[1].copy(), list(chit[2])
            return self._finish(data_array, affine_, new_affine, output_shape, original_shape, mode, padding_mode, align_corners, dtype,
                                output_spatial_shape, lazy)
        affine_ = U.to_affine_nd(sr, input_affine)
        out_d = self.pixdim[:sr].copy()
        if out_d.size < sr:
            out_d = np.append(out_d, [out_d[-1]] * (sr - out_d.size))
        orig_d = U.affine_to_spacing(affine_, sr)
        mins = list(self.min_pixdim[:sr]) + [np.nan] * sr
        maxs = list(self.max_pixdim[:sr]) + [np.nan] * sr
        for idx, _d in enumerate(orig_d):
            target = out_d[idx]
            mn = target if np.isnan(mins[idx]) else min(mins[idx], target)
            mx = target if np.isnan(maxs[idx]) else max(maxs[idx], target)
            if mn > mx:
                raise ValueError(f"min_pixdim is larger than max_pixdim at dim {idx}: min {mn} max {mx} out {target}.")
            out_d[idx] = _d if (mn - U.AFFINE_TOL) <= _d <= (mx + U.AFFINE_TOL) else target
        new_affine = U.zoom_affine(affine_, out_d, diagonal=self.diagonal)
        output_shape, offset = U.compute_shape_offset(original_shape, affine_, new_affine, scale_extent)
        new_affine[:sr, -1] = offset[:sr]
        if len(self._algebra_cache) >= 64:
            self._algebra_cache.clear()
        self._algebra_cache[ckey] = (np.array(affine_, copy=True), np.array(new_affine, copy=True), [int(v) for v in output_shape])
        return self._finish(data_array, affine_, new_affine, output_shape, original_shape, mode, padding_mode, align_corners, dtype,
                            output_spatial_shape, lazy)

    def _finish(self, data_array, affine_, new_affine, output_shape, original_shape, mode, padding_mode, align_corners, dtype, output_spatial_shape, lazy):
        actual_shape = list(output_shape) if output_spatial_shape is None else output_spatial_shape
        lazy_ = bool(getattr(self.sp_resample, "lazy", False)) if lazy is None else bool(lazy)
        kw = {"lazy": True} if lazy_ else {}
        out = self.sp_resample(data_array, dst_affine=new_affine, spatial_size=actual_shape, mode=mode, padding_mode=padding_mode,
                               align_corners=align_corners, dtype=dtype, **kw)
        if lazy_:
            return out
        if self.recompute_affine and is_meta(out):
            # scale_affine(original, actual) (monai/transforms/utils.py): centre-preserving rescale of the index grid
            r = len(original_shape)
            scale = np.asarray([o / max(a, 1) for o, a in zip(original_shape, actual_shape)], dtype=np.float64)
            a = np.diag(np.append(scale, 1.0))
            a[:r, -1] = (scale - 1) / 2.0
            out.affine = torch.as_tensor(affine_ @ a)
        return out


class Spacingd(MapTransform):
    """Dictionary version (spatial/dictionary.py:365-531): every key is resampled; with `ensure_same_shape` later keys
    whose input shape equals the first key's reuse the first key's output shape."""

    def __init__(self, keys, pixdim, diagonal: bool = False, mode="bilinear", padding_mode="border", align_corners=False, dtype=np.float64,
                 scale_extent: bool = False, recompute_affine: bool = False, min_pixdim=None, max_pixdim=None, ensure_same_shape: bool = True,
                 allow_missing_keys: bool = False, lazy: bool = False):
        super().__init__(keys, allow_missing_keys)
        self.spacing_transform = Spacing(pixdim, diagonal=diagonal, recompute_affine=recompute_affine, min_pixdim=min_pixdim, max_pixdim=max_pixdim, lazy=lazy)
        n = len(self.keys)
        rep = lambda v: tuple(v) if isinstance(v, (list, tuple)) else (v,) * n  # noqa: E731
        self.mode, self.padding_mode, self.align_corners, self.dtype, self.scale_extent = rep(mode), rep(padding_mode), rep(align_corners), rep(dtype), rep(scale_extent)
        self.ensure_same_shape = ensure_same_shape

    def inverse(self, data: Mapping[Hashable, Any]) -> dict:
        d = dict(data)
        for key in self.key_iterator(d):
            d[key] = self.spacing_transform.inverse(d[key])
        return d

    @property
    def lazy(self) -> bool:
        return self.spacing_transform.sp_resample.lazy

    @lazy.setter
    def lazy(self, v: bool) -> None:
        self.spacing_transform.sp_resample.lazy = bool(v)

    def __call__(self, data: Mapping[Hashable, Any], lazy=None) -> dict:
        d = dict(data)
        _init_shape, _pixdim, should_match = None, None, False
        output_shape_k = None
        lazy_ = self.lazy if lazy is None else lazy
        for i, key in enumerate(self.key_iterator(d)):
            idx = self.keys.index(key)
            if self.ensure_same_shape and is_meta(d[key]):
                # the reference reuses the first key's output shape only for keys with the same input shape AND the same
                # affine-derived spacing (monai/transforms/spatial/dictionary.py:500-512)
                pix_k = U.affine_to_spacing(np.asarray(get_affine(d[key]), dtype=np.float64), len(d[key].shape) - 1)
                shp_k = tuple(d[key].peek_pending_shape()) if hasattr(d[key], "peek_pending_shape") else tuple(d[key].shape[1:])
                if _init_shape is None:
                    _init_shape, _pixdim = shp_k, pix_k
                else:
                    should_match = shp_k == _init_shape and bool(np.allclose(_pixdim, pix_k, atol=1e-3))
            d[key] = self.spacing_transform(
                d[key], mode=self.mode[idx], padding_mode=self.padding_mode[idx], align_corners=self.align_corners[idx], dtype=self.dtype[idx],
                scale_extent=self.scale_extent[idx], output_spatial_shape=output_shape_k if should_match else None, lazy=lazy_,
            )
            if output_shape_k is None:
                output_shape_k = tuple(d[key].peek_pending_shape()) if lazy_ else tuple(d[key].shape[1:])
        return d


class AffineGrid(Transform):
    """Affine matrix builder (spatial/array.py:1662-1783).  The dense grid is only materialised on request
    (`__call__(..., grid=None)` returns (grid, affine) like the reference); the resamplers use the matrix alone."""

    def __init__(self, rotate_params=None, shear_params=None, translate_params=None, scale_params=None, device=None, dtype=np.float32,
                 align_corners: bool = False, affine=None, lazy: bool = False):
        self.rotate_params, self.shear_params, self.translate_params, self.scale_params = rotate_params, shear_params, translate_params, scale_params
        self.device, self.dtype, self.align_corners, self.affine = device, dtype, align_corners, affine

    def matrix(self, spatial_dims: int) -> np.ndarray:
        if self.affine is not None:
            return U.to_affine_nd(spatial_dims, self.affine)
        affine = np.eye(spatial_dims + 1, dtype=np.float32)
        if self.rotate_params:
            affine = affine @ U.create_rotate(spatial_dims, self.rotate_params)
        if self.shear_params:
            affine = affine @ U.create_shear(spatial_dims, self.shear_params)
        if self.translate_params:
            affine = affine @ U.create_translate(spatial_dims, self.translate_params)
        if self.scale_params:
            affine = affine @ U.create_scale(spatial_dims, self.scale_params)
        return U.to_affine_nd(spatial_dims, affine.astype(np.float32))

    def __call__(self, spatial_size=None, grid=None, lazy=None):
        if grid is None and spatial_size is None:
            raise ValueError("Incompatible values: grid=None and spatial_size=None.")
        if grid is None:
            r = len(spatial_size)
            axes = [torch.linspace(-(d - 1.0) / 2.0, (d - 1.0) / 2.0, int(d), dtype=torch.float32) for d in spatial_size]
            coords = torch.meshgrid(*axes, indexing="ij")
            grid = torch.stack([*coords, torch.ones_like(coords[0])])
        r = grid.shape[0] - 1
        affine = torch.as_tensor(self.matrix(r), dtype=grid.dtype, device=grid.device)
        a = affine
        if self.align_corners:
            sc = torch.as_tensor(U.create_scale(r, [max(d, 2) / (max(d, 2) - 1) for d in grid.shape[1:]], dtype=np.float64), dtype=grid.dtype, device=grid.device)
            a = affine @ sc
        out = (a @ grid.reshape(grid.shape[0], -1)).reshape(-1, *grid.shape[1:])
        return out, affine


class RandAffineGrid(Randomizable, Transform):
    """Random affine parameters with the reference's draw order (spatial/array.py:1786-1915): rotate, shear,
    translate, scale (+1.0); each entry `uniform(-f, f)` or `uniform(f[0], f[1])`."""

    def __init__(self, rotate_range=None, shear_range=None, translate_range=None, scale_range=None, device=None, dtype=np.float32, lazy: bool = False):
        _t = lambda v: () if v is None else (tuple(v) if isinstance(v, (list, tuple, np.ndarray)) else (v,))  # noqa: E731
        self.rotate_range, self.shear_range, self.translate_range, self.scale_range = _t(rotate_range), _t(shear_range), _t(translate_range), _t(scale_range)
        self.rotate_params = self.shear_params = self.translate_params = self.scale_params = None
        self.device, self.dtype = device, dtype
        self.affine = torch.eye(4, dtype=torch.float64)

    def _get_rand_param(self, param_range, add_scalar: float = 0.0):
        out = []
        for f in param_range:
            if isinstance(f, (list, tuple, np.ndarray)):
                if len(f) != 2:
                    raise ValueError(f"If giving range as [min,max], should have 2 elements per dim, got {f}.")
                out.append(self.R.uniform(f[0], f[1]) + add_scalar)
            elif f is not None:
                out.append(self.R.uniform(-f, f) + add_scalar)
        return out

    def randomize(self, data=None) -> None:
        self.rotate_params = self._get_rand_param(self.rotate_range)
        self.shear_params = self._get_rand_param(self.shear_range)
        self.translate_params = self._get_rand_param(self.translate_range)
        self.scale_params = self._get_rand_param(self.scale_range, 1.0)

    def matrix(self, spatial_dims: int, randomize: bool = True) -> np.ndarray:
        if randomize:
            self.randomize()
        ag = AffineGrid(self.rotate_params, self.shear_params, self.translate_params, self.scale_params)
        m = ag.matrix(spatial_dims)
        self.affine = torch.as_tensor(m, dtype=torch.float32)
        return m

    def __call__(self, spatial_size=None, grid=None, randomize: bool = True, lazy=None):
        if randomize:
            self.randomize()
        ag = AffineGrid(self.rotate_params, self.shear_params, self.translate_params, self.scale_params)
        g, self.affine = ag(spatial_size, grid)
        return g

    def get_transformation_matrix(self):
        return self.affine


class Resample(Transform):
    """Resample with an explicit grid (spatial/array.py:1962-2117).  Affine grids produced by this package (AffineSpec) are
    resolved to their matrix (no dense grid traffic); a dense grid tensor [3|4, H, W[, D]] (a deformation field, the output of
    the reference's AffineGrid, ...) is sampled by the dense-grid kernel (b200_grid_pull) with the torch path's conventions:
    centred voxel coordinates scaled by 2 / max(2, size) when `norm_coords`, values in [-1, 1] otherwise, grid_sample's
    un-normalisation for `align_corners`, zeros / border / reflection padding, bilinear or nearest interpolation."""

    def __init__(self, mode="bilinear", padding_mode="border", norm_coords: bool = True, device=None, align_corners: bool = False, dtype=np.float64):
        self.mode, self.padding_mode, self.norm_coords, self.device, self.align_corners, self.dtype = mode, padding_mode, norm_coords, device, align_corners, dtype

    def _dense(self, img, grid, mode, padding_mode, align: bool):
        if not img.is_cuda:
            raise RuntimeError("monai_b200 spatial transforms run on CUDA tensors only (there is no CPU fallback)")
        t = img.as_subclass(torch.Tensor) if type(img) is not torch.Tensor else img
        if t.dtype not in (torch.float16, torch.float32):
            t = t.float()
        sr = min(t.dim() - 1, 3)
        g = torch.as_tensor(grid)
        g = (g.as_subclass(torch.Tensor) if type(g) is not torch.Tensor else g).to(t.device)
        if g.shape[0] < sr or tuple(g.shape[1:]) == () or g.dim() != sr + 1:
            raise ValueError(f"grid must be [{sr} or {sr + 1}, *spatial], got {tuple(g.shape)}")
        if g.dtype not in (torch.float32, torch.float64):
            g = g.to(torch.float64 if self.dtype in (np.float64, torch.float64, None) else torch.float32)
        in_sp = tuple(int(s) for s in t.shape[1 : 1 + sr])
        # voxel coordinate of axis i = a_i * grid[i] + b_i: (x 2/max(2,size)) then grid_sample's un-normalisation
        scale, shift = [], []
        for size in in_sp:
            k = 2.0 / max(2, size) if self.norm_coords else 1.0
            if align:
                scale.append(k * (size - 1) / 2.0); shift.append((size - 1) / 2.0)
            else:
                scale.append(k * size / 2.0); shift.append((size - 1) / 2.0)
        pad_code = _pad(padding_mode)
        bound = {L.PAD_ZEROS: 7, L.PAD_BORDER: 0, L.PAD_REFLECTION: 1 if align else 2}[pad_code]   # reflection: about centres (dct1) / edges (dct2)
        order = 1 if _mode(mode) == L.INTERP_LINEAR else 0
        lift = 3 - sr
        out_sp = tuple(int(s) for s in g.shape[1:])
        t3 = t.reshape(1, t.shape[0], *([1] * lift), *in_sp)
        g3 = g[:sr].reshape(sr, *([1] * lift), *out_sp)
        if lift:
            g3 = torch.cat([torch.zeros((lift, *g3.shape[1:]), dtype=g3.dtype, device=g3.device), g3], 0)
        out = K.grid_pull(t3, g3.contiguous(), [0] * lift + [bound] * sr, [0] * lift + [order] * sr, extrapolate=True, channel_last=False,
                          scale=[1.0] * lift + scale, shift=[0.0] * lift + shift, half_even=True, out_dtype=torch.float32)
        return rewrap(out.reshape(t.shape[0], *out_sp), img)

    def __call__(self, img, grid=None, mode=None, padding_mode=None, dtype=None, align_corners=None):
        if grid is None:
            return img
        align = self.align_corners if align_corners is None else align_corners
        if not isinstance(grid, AffineSpec):
            return self._dense(img, grid, self.mode if mode is None else mode, self.padding_mode if padding_mode is None else padding_mode, bool(align))
        if not self.norm_coords:
            raise NotImplementedError("norm_coords=False with an affine grid specification is not implemented (pass the dense grid)")
        r = len(grid.spatial_size)
        mat = U.sample_matrix_from_centered_affine(grid.affine if not (grid.grid_align_corners) else grid.affine, tuple(img.shape[1 : 1 + r]), grid.spatial_size, align)
        out = _resample(img, mat, r, grid.spatial_size, self.mode if mode is None else mode, self.padding_mode if padding_mode is None else padding_mode, align)
        return rewrap(out, img)


class AffineSpec:
    """An affine sampling grid in closed form: `affine` acts on the centred index grid of `spatial_size`."""

    def __init__(self, affine: np.ndarray, spatial_size: Sequence[int], grid_align_corners: bool = False):
        self.affine, self.spatial_size, self.grid_align_corners = np.asarray(affine, dtype=np.float64), tuple(int(s) for s in spatial_size), grid_align_corners


def _update_affine(img, affine_centered: np.ndarray, src_shape, dst_shape):
    """metadata update of affine_func (spatial/functional.py:548-613): new = old @ (T_src_center @ A @ T_dst_center^-1)."""
    if not is_meta(img):
        return None
    r = len(dst_shape)
    old = get_affine(img).numpy()
    t_src, t_dst = np.eye(r + 1), np.eye(r + 1)
    t_src[:r, -1] = (np.asarray(src_shape[:r], dtype=np.float64) - 1) / 2.0
    t_dst[:r, -1] = -(np.asarray(dst_shape, dtype=np.float64) - 1) / 2.0
    m = t_src @ U.to_affine_nd(r, affine_centered) @ t_dst
    full = np.eye(len(old))
    full[:r, :r], full[:r, -1] = m[:r, :r], m[:r, -1]
    return old @ full


class Affine(Transform):
    """Deterministic affine transform (spatial/array.py:2120-2316)."""

    def __init__(self, rotate_params=None, shear_params=None, translate_params=None, scale_params=None, affine=None, spatial_size=None, mode="bilinear",
                 padding_mode="reflection", normalized: bool = False, device=None, dtype=np.float32, align_corners: bool = False, image_only: bool = False, lazy: bool = False):
        if normalized:
            raise NotImplementedError("Affine(normalized=True) is not implemented")
        self.affine_grid = AffineGrid(rotate_params, shear_params, translate_params, scale_params, affine=affine, dtype=dtype, align_corners=align_corners)
        self.image_only, self.spatial_size, self.mode, self.padding_mode, self.align_corners = image_only, spatial_size, mode, padding_mode, align_corners

    def __call__(self, img, spatial_size=None, mode=None, padding_mode=None, lazy=None):
        ori = tuple(img.shape[1:])
        r = min(len(ori), 3)
        sp = _fall_back(self.spatial_size if spatial_size is None else spatial_size, ori[:r])
        a = self.affine_grid.matrix(r)
        a_eff = a
        if self.align_corners:
            n = np.asarray(sp, dtype=np.float64)
            a_eff = a @ np.diag(np.append(np.maximum(n, 2) / (np.maximum(n, 2) - 1), 1.0))
        n_s = np.asarray(ori[:r], dtype=np.float64)
        norm = np.diag(np.append(2.0 / np.maximum(2.0, n_s), 1.0))
        center = np.eye(r + 1)
        center[:-1, -1] = -(np.asarray(sp, dtype=np.float64) - 1) / 2.0
        mat = U._unnormalize(ori[:r], self.align_corners) @ norm @ a_eff @ center
        out = _resample(img, mat, r, sp, self.mode if mode is None else mode, self.padding_mode if padding_mode is None else padding_mode, self.align_corners)
        out = rewrap(out, img, _update_affine(img, a, ori, sp), {"class": "Affine", "affine": a})
        return out if self.image_only else (out, torch.as_tensor(a))


class RandAffine(RandomizableTransform):
    """Random affine transform (spatial/array.py:2317-2577)."""

    def __init__(self, prob: float = 0.1, rotate_range=None, shear_range=None, translate_range=None, scale_range=None, spatial_size=None, mode="bilinear",
                 padding_mode="reflection", cache_grid: bool = False, device=None, lazy: bool = False):
        RandomizableTransform.__init__(self, prob)
        self.rand_affine_grid = RandAffineGrid(rotate_range, shear_range, translate_range, scale_range, device=device)
        self.resampler = Resample(device=device)
        self.spatial_size, self.mode, self.padding_mode, self.cache_grid, self.lazy = spatial_size, mode, padding_mode, cache_grid, lazy

    def inverse(self, data):
        """spatial/array.py:2545-2577: resample with the inverse of the recorded matrix back to the original size."""
        op = _pop_applied(data, "RandAffine")
        if not op.get("do_resampling", True):
            return data
        orig_size = tuple(int(v) for v in op["orig_size"])
        inv = np.linalg.inv(np.asarray(op["affine"], dtype=np.float64))
        r = len(orig_size)
        out = self.resampler(data, grid=AffineSpec(inv, orig_size), mode=op["mode"], padding_mode=op["padding_mode"])
        return rewrap(out.as_subclass(torch.Tensor) if type(out) is not torch.Tensor else out, data, _update_affine(data, inv, tuple(data.shape[1 : 1 + r]), orig_size))

    def set_random_state(self, seed=None, state=None):
        self.rand_affine_grid.set_random_state(seed, state)
        super().set_random_state(seed, state)
        return self

    def randomize(self, data=None) -> None:
        super().randomize(None)
        if not self._do_transform:
            return None
        self.rand_affine_grid.randomize()

    def get_identity_grid(self, spatial_size, lazy: bool = False) -> AffineSpec:
        return AffineSpec(np.eye(len(spatial_size) + 1), spatial_size)

    def __call__(self, img, spatial_size=None, mode=None, padding_mode=None, randomize: bool = True, grid=None, lazy=None):
        if randomize:
            self.randomize()
        lazy_ = self.lazy if lazy is None else lazy
        ori = tuple(img.peek_pending_shape()) if is_meta(img) and getattr(img, "pending_operations", None) else tuple(img.shape[1:])
        r = min(len(ori), 3)
        sp = _fall_back(self.spatial_size if spatial_size is None else spatial_size, ori[:r])
        do_resampling = self._do_transform or (sp != tuple(ori[:r]))
        _mode_ = self.mode if mode is None else mode
        _pad_ = self.padding_mode if padding_mode is None else padding_mode
        if grid is None or not isinstance(grid, AffineSpec):
            grid = self.get_identity_grid(sp)
            if self._do_transform:
                grid = AffineSpec(self.rand_affine_grid.matrix(r, randomize=randomize), sp)
        affine = self.rand_affine_grid.get_transformation_matrix()
        a = np.asarray(grid.affine)
        info = {"class": "RandAffine", "affine": a, "rand_affine_matrix": affine, "orig_size": ori[:r], "mode": getattr(_mode_, "value", _mode_),
                "padding_mode": getattr(_pad_, "value", _pad_), "do_resampling": bool(do_resampling)}
        if lazy_:   # affine_func(lazy=True): pending matrix = centre shifts around the centred affine (Affine.compute_w_affine)
            from .lazy import push_pending

            t_src, t_dst = np.eye(r + 1), np.eye(r + 1)
            t_src[:r, -1] = (np.asarray(ori[:r], dtype=np.float64) - 1) / 2.0
            t_dst[:r, -1] = -(np.asarray(sp, dtype=np.float64) - 1) / 2.0
            return push_pending(img, info, t_src @ U.to_affine_nd(r, a) @ t_dst, sp)
        if not do_resampling:
            t = img.as_subclass(torch.Tensor) if type(img) is not torch.Tensor else img
            return rewrap(t.to(torch.float32), img, None, info)
        out = self.resampler(img, grid=grid, mode=_mode_, padding_mode=_pad_)
        return rewrap(out.as_subclass(torch.Tensor) if type(out) is not torch.Tensor else out, img, _update_affine(img, a, ori, sp), info)


class RandAffined(RandomizableTransform, MapTransform):
    """Dictionary version (spatial/dictionary.py:1001-1175): one set of random parameters shared by all keys."""

    def __init__(self, keys, spatial_size=None, prob: float = 0.1, rotate_range=None, shear_range=None, translate_range=None, scale_range=None,
                 mode="bilinear", padding_mode="reflection", cache_grid: bool = False, device=None, allow_missing_keys: bool = False, lazy: bool = False):
        MapTransform.__init__(self, keys, allow_missing_keys)
        RandomizableTransform.__init__(self, prob)
        self.rand_affine = RandAffine(prob=1.0, rotate_range=rotate_range, shear_range=shear_range, translate_range=translate_range, scale_range=scale_range,
                                      spatial_size=spatial_size, cache_grid=cache_grid, device=device)
        n = len(self.keys)
        rep = lambda v: tuple(v) if isinstance(v, (list, tuple)) else (v,) * n  # noqa: E731
        self.mode, self.padding_mode = rep(mode), rep(padding_mode)

    def set_random_state(self, seed=None, state=None):
        self.rand_affine.set_random_state(seed, state)
        super().set_random_state(seed, state)
        return self

    def inverse(self, data: Mapping[Hashable, Any]) -> dict:
        d = dict(data)
        for key in self.key_iterator(d):
            d[key] = self.rand_affine.inverse(d[key])
        return d

    @property
    def lazy(self) -> bool:
        return self.rand_affine.lazy

    @lazy.setter
    def lazy(self, v: bool) -> None:
        self.rand_affine.lazy = bool(v)

    def __call__(self, data: Mapping[Hashable, Any], lazy=None) -> dict:
        d = dict(data)
        keys = [k for k in self.key_iterator(d)]
        if not keys:
            return d
        lazy_ = self.lazy if lazy is None else lazy
        self.randomize(None)
        self.rand_affine.randomize()  # all the keys share the same random affine factor
        item = d[keys[0]]
        ori = tuple(item.peek_pending_shape()) if is_meta(item) and getattr(item, "pending_operations", None) else tuple(item.shape[1:])
        r = min(len(ori), 3)
        sp = _fall_back(self.rand_affine.spatial_size, ori[:r])
        do_resampling = self._do_transform or (sp != tuple(ori[:r]))
        grid = None
        if do_resampling:
            grid = self.rand_affine.get_identity_grid(sp)
            if self._do_transform:
                grid = AffineSpec(self.rand_affine.rand_affine_grid.matrix(r, randomize=True), sp)
        for key in keys:
            idx = self.keys.index(key)
            if do_resampling:
                # the reference passes randomize=True here too (dictionary.py:1156): the per-key redraw only consumes
                # RNG state, the shared grid is what is applied
                d[key] = self.rand_affine(d[key], None, self.mode[idx], self.padding_mode[idx], True, grid, lazy=lazy_)
            elif not lazy_:
                t = d[key].as_subclass(torch.Tensor) if type(d[key]) is not torch.Tensor else d[key]
                d[key] = rewrap(t.to(torch.float32), d[key])
        return d
