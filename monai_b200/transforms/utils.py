"""Host-side affine algebra for the spatial transforms (float64 numpy unless noted) -- SURVEY.md §8 rows a16-a18.

Restates, with identical arithmetic and rounding:
  affine_to_spacing / zoom_affine / compute_shape_offset / to_affine_nd  (monai/data/utils.py:737-760, 823-872, 875-935, 938-984)
  create_rotate / create_shear / create_scale / create_translate          (monai/transforms/utils.py:859-1075)
and composes, for the CUDA resampler, the single 3x4 matrix that maps an OUTPUT voxel index to a SOURCE voxel
coordinate -- the composition the reference reaches through normalize_transform / to_norm_affine
(monai/networks/utils.py:243-326), F.affine_grid and F.grid_sample's un-normalisation.
"""
from __future__ import annotations

from typing import Sequence

import numpy as np
import torch

AFFINE_TOL = 1e-3

__all__ = [
    "AFFINE_TOL", "to_affine_nd", "affine_to_spacing", "zoom_affine", "compute_shape_offset", "create_rotate", "create_shear",
    "create_scale", "create_translate", "sample_matrix_from_xform", "sample_matrix_from_centered_affine", "lift_to_3d",
]


def _np(a) -> np.ndarray:
    if isinstance(a, torch.Tensor):
        a = a.detach().cpu().numpy()
    return np.asarray(a, dtype=np.float64)


def to_affine_nd(r, affine) -> np.ndarray:
    """(r+1)x(r+1) affine whose top-left kxk block and last column come from `affine` (k = min(r, len(affine)-1))."""
    a = _np(affine).copy()
    if a.ndim != 2:
        raise ValueError(f"affine must have 2 dimensions, got {a.ndim}.")
    new = np.array(r, dtype=np.float64, copy=True)
    if new.ndim == 0:
        sr = int(new)
        if sr < 0:
            raise ValueError(f"r must be positive, got {sr}.")
        new = np.eye(sr + 1, dtype=np.float64)
    d = max(min(len(new) - 1, len(a) - 1), 1)
    new[:d, :d] = a[:d, :d]
    if d > 1:
        new[:d, -1] = a[:d, -1]
    return new


def affine_to_spacing(affine, r: int = 3, suppress_zeros: bool = True) -> np.ndarray:
    a = _np(affine)
    if a.ndim != 2 or a.shape[0] != a.shape[1]:
        raise ValueError(f"affine must be a square matrix, got {a.shape}.")
    sub = a[:r, :r]
    spacing = np.sqrt(np.sum(sub * sub, axis=0))
    if suppress_zeros:
        spacing[spacing == 0] = 1.0
    return spacing


def zoom_affine(affine, scale, diagonal: bool = True) -> np.ndarray:
    a = np.array(_np(affine), dtype=float, copy=True)
    if len(a) != len(a[0]):
        raise ValueError(f"affine must be n x n, got {len(a)} x {len(a[0])}.")
    scale_np = np.array(scale, dtype=float, copy=True)
    d = len(a) - 1
    norm = affine_to_spacing(a, r=d)
    if len(scale_np) < d:
        scale_np = np.append(scale_np, norm[len(scale_np):])
    scale_np = scale_np[:d]
    scale_np = np.asarray([s if s > 0 else n for s, n in zip(scale_np, norm)], dtype=float)  # fall_back_tuple
    scale_np[scale_np == 0] = 1.0
    if diagonal:
        return np.diag(np.append(scale_np, [1.0]))
    rzs = a[:-1, :-1]
    zs = np.linalg.cholesky(rzs.T @ rzs).T
    rotation = rzs @ np.linalg.inv(zs)
    s = np.sign(np.diag(zs)) * np.abs(scale_np)
    new_affine = np.eye(len(a))
    new_affine[:-1, :-1] = rotation @ np.diag(s)
    return new_affine


def compute_shape_offset(spatial_shape, in_affine, out_affine, scale_extent: bool = False):
    shape = np.array(spatial_shape, copy=True, dtype=float)
    sr = len(shape)
    in_a, out_a = to_affine_nd(sr, in_affine), to_affine_nd(sr, out_affine)
    in_coords = [(-0.5, dim - 0.5) if scale_extent else (0.0, dim - 1.0) for dim in shape]
    corners = np.asarray(np.meshgrid(*in_coords, indexing="ij")).reshape((len(shape), -1))
    corners = np.concatenate((corners, np.ones_like(corners[:1])))
    try:
        corners_out = np.linalg.solve(out_a, in_a) @ corners
    except np.linalg.LinAlgError as e:
        raise ValueError(f"Affine {out_a} is not invertible") from e
    corners = in_a @ corners
    all_dist = corners_out[:-1].copy()
    corners_out = corners_out[:-1] / corners_out[-1]
    out_shape = np.round(np.ptp(corners_out, axis=1)) if scale_extent else np.round(np.ptp(corners_out, axis=1) + 1.0)
    offset = None
    for i in range(corners.shape[1]):
        min_corner = np.min(all_dist - all_dist[:, i : i + 1], 1)
        if np.allclose(min_corner, 0.0, rtol=AFFINE_TOL):
            offset = corners[:-1, i]
            break
    if offset is None:
        offset = in_a[:-1, :-1] @ (shape / 2.0) + in_a[:-1, -1] - out_a[:-1, :-1] @ (out_shape / 2.0)
    if scale_extent:
        in_offset = np.append(0.5 * (shape / out_shape - 1.0), 1.0)
        offset = np.abs((in_a @ in_offset / in_offset[-1])[:-1]) * np.sign(offset)
    return out_shape.astype(int, copy=False), offset


# The reference builds these with the torch backend in float32 (AffineGrid, spatial/array.py:1758-1770); float32 is kept.
def _f32(v) -> np.float32:
    return np.float32(v)


def create_rotate(spatial_dims: int, radians) -> np.ndarray:
    radians = tuple(radians) if isinstance(radians, (list, tuple, np.ndarray)) else (radians,)
    sin = lambda t: torch.sin(torch.as_tensor(t, dtype=torch.float32)).item()  # noqa: E731
    cos = lambda t: torch.cos(torch.as_tensor(t, dtype=torch.float32)).item()  # noqa: E731
    if spatial_dims == 2:
        if len(radians) < 1:
            raise ValueError("radians must be non empty.")
        out = np.eye(3, dtype=np.float32)
        s, c = sin(radians[0]), cos(radians[0])
        out[0, 0], out[0, 1], out[1, 0], out[1, 1] = c, -s, s, c
        return out
    if spatial_dims != 3:
        raise ValueError(f"Unsupported spatial_dims: {spatial_dims}, available options are [2, 3].")
    if len(radians) < 1:
        raise ValueError("radians must be non empty.")
    s, c = sin(radians[0]), cos(radians[0])
    affine = np.eye(4, dtype=np.float32)
    affine[1, 1], affine[1, 2], affine[2, 1], affine[2, 2] = c, -s, s, c
    if len(radians) >= 2:
        s, c = sin(radians[1]), cos(radians[1])
        m = np.eye(4, dtype=np.float32)
        m[0, 0], m[0, 2], m[2, 0], m[2, 2] = c, s, -s, c
        affine = affine @ m
    if len(radians) >= 3:
        s, c = sin(radians[2]), cos(radians[2])
        m = np.eye(4, dtype=np.float32)
        m[0, 0], m[0, 1], m[1, 0], m[1, 1] = c, -s, s, c
        affine = affine @ m
    return affine


def create_shear(spatial_dims: int, coefs) -> np.ndarray:
    coefs = tuple(coefs) if isinstance(coefs, (list, tuple, np.ndarray)) else (coefs,)
    if spatial_dims == 2:
        coefs = (coefs + (0.0,) * 2)[:2]
        out = np.eye(3, dtype=np.float32)
        out[0, 1], out[1, 0] = coefs[0], coefs[1]
        return out
    if spatial_dims == 3:
        coefs = (coefs + (0.0,) * 6)[:6]
        out = np.eye(4, dtype=np.float32)
        out[0, 1], out[0, 2], out[1, 0], out[1, 2], out[2, 0], out[2, 1] = coefs
        return out
    raise NotImplementedError("Currently only spatial_dims in [2, 3] are supported.")


def create_scale(spatial_dims: int, scaling_factor, dtype=np.float32) -> np.ndarray:
    f = tuple(scaling_factor) if isinstance(scaling_factor, (list, tuple, np.ndarray)) else (scaling_factor,)
    f = (f + (1.0,) * spatial_dims)[:spatial_dims]
    return np.diag(np.asarray(f + (1.0,), dtype=dtype))


def create_translate(spatial_dims: int, shift, dtype=np.float32) -> np.ndarray:
    shift = tuple(shift) if isinstance(shift, (list, tuple, np.ndarray)) else (shift,)
    out = np.eye(int(spatial_dims) + 1, dtype=dtype)
    for i, a in enumerate(shift[:spatial_dims]):
        out[i, spatial_dims] = a
    return out


# --------------------------------------------------------------------------------- output index -> source coordinate
def _norm_false(shape) -> np.ndarray:
    """normalize_transform(shape, align_corners=False, zero_centered=False): [-0.5, d-0.5] -> [-1, 1]."""
    shape = np.asarray(shape, dtype=np.float64)
    norm = shape.copy()
    norm[norm <= 0.0] = 2.0
    m = np.diag(np.append(2.0 / norm, 1.0))
    m[:-1, -1] = 1.0 / shape - 1.0
    return m


def _base_grid(shape, align_corners: bool) -> np.ndarray:
    """index j -> F.affine_grid base coordinate: (2j+1)/n - 1, or 2j/(n-1) - 1 when align_corners."""
    n = np.asarray(shape, dtype=np.float64)
    if align_corners:
        den = np.where(n > 1, n - 1, 1.0)
        m = np.diag(np.append(np.where(n > 1, 2.0 / den, 0.0), 1.0))
        m[:-1, -1] = np.where(n > 1, -1.0, 0.0)
    else:
        m = np.diag(np.append(2.0 / n, 1.0))
        m[:-1, -1] = 1.0 / n - 1.0
    return m


def _unnormalize(shape, align_corners: bool) -> np.ndarray:
    """grid_sample's un-normalisation: ((g+1)*S-1)/2, or (g+1)/2*(S-1) when align_corners."""
    s = np.asarray(shape, dtype=np.float64)
    if align_corners:
        m = np.diag(np.append((s - 1) / 2.0, 1.0))
        m[:-1, -1] = (s - 1) / 2.0
    else:
        m = np.diag(np.append(s / 2.0, 1.0))
        m[:-1, -1] = (s - 1) / 2.0
    return m


def sample_matrix_from_xform(xform, src_shape, dst_shape, align_corners: bool) -> np.ndarray:
    """spatial_resample's torch path (spatial/functional.py:174-179): AffineTransform(normalized=False) always
    normalises with align_corners=False, then affine_grid / grid_sample run with the caller's align_corners."""
    x = _np(xform)
    theta = _norm_false(src_shape) @ x @ np.linalg.inv(_norm_false(dst_shape))
    return _unnormalize(src_shape, align_corners) @ theta @ _base_grid(dst_shape, align_corners)


def sample_matrix_from_centered_affine(affine, src_shape, dst_shape, align_corners: bool) -> np.ndarray:
    """Affine / RandAffine path (spatial/array.py:1758-1783, 2102-2115): grid = A @ centred index grid
    (create_grid: linspace(-(d-1)/2, (d-1)/2)), with AffineGrid's align_corners pre-scale, then x 2/max(2,dim) and
    grid_sample's un-normalisation."""
    a = _np(affine)
    n = np.asarray(dst_shape, dtype=np.float64)
    s = np.asarray(src_shape, dtype=np.float64)
    r = len(dst_shape)
    center = np.eye(r + 1)
    center[:-1, -1] = -(n - 1) / 2.0
    if align_corners:
        a = a @ np.diag(np.append(np.maximum(n, 2) / (np.maximum(n, 2) - 1), 1.0))
    norm = np.diag(np.append(2.0 / np.maximum(2.0, s), 1.0))
    return _unnormalize(src_shape, align_corners) @ norm @ a @ center


def lift_to_3d(m: np.ndarray, r: int) -> np.ndarray:
    """(r+1)x(r+1) index->coordinate matrix -> 3x4 row-major for the 3-D kernel (leading singleton axes)."""
    out = np.zeros((3, 4), dtype=np.float64)
    lift = 3 - r
    for i in range(lift):
        out[i, i] = 1.0
    out[lift:, lift:3] = m[:r, :r]
    out[lift:, 3] = m[:r, r]
    return out


def _normalize_matrix(shape, align_corners: bool, zero_centered: bool) -> np.ndarray:
    """normalize_transform (monai/networks/utils.py:243-286): voxel index -> [-1, 1] for the four (align_corners, zero_centered) cases."""
    shape = np.asarray(shape, dtype=np.float64)
    norm = shape.copy()
    r = len(shape)
    m = np.eye(r + 1)
    with np.errstate(divide="ignore"):   # a unit-size axis with zero_centered gives 2 / 0 = inf, as the reference's torch code does
        if align_corners:
            norm[norm <= 1.0] = 2.0
            m[np.arange(r), np.arange(r)] = 2.0 / (norm if zero_centered else norm - 1.0)
            if not zero_centered:
                m[:-1, -1] = -1.0
        else:
            norm[norm <= 0.0] = 2.0
            m[np.arange(r), np.arange(r)] = 2.0 / (norm - 1.0 if zero_centered else norm)
            if not zero_centered:
                m[:-1, -1] = 1.0 / shape - 1.0
    return m


def affine_transform_matrix(theta, src_shape, dst_shape, normalized: bool, reverse_indexing: bool, align_corners: bool,
                            zero_centered: bool = False) -> np.ndarray:
    """The output-index -> source-index matrix of AffineTransform.forward (monai/networks/layers/spatial_transforms.py:556-592) for one
    homogeneous (r+1)x(r+1) `theta`, in the image's own (i, j, k) axis order:
      * normalized=False: theta is taken to normalised coordinates by to_norm_affine(align_corners=False, zero_centered)
        (networks/utils.py:289-326: norm(src) @ theta @ inv(norm(dst)));
      * reverse_indexing=True flips rows / columns so that affine_grid's (x, y, z) = (k, j, i) convention sees an (i, j, k) theta --
        in (i, j, k) order that is the identity; with reverse_indexing=False theta is GIVEN in (x, y, z) order and is flipped here;
      * affine_grid's base coordinates and grid_sample's un-normalisation with the layer's align_corners close the chain."""
    t = np.asarray(_np(theta), dtype=np.float64)
    r = len(src_shape)
    if t.shape != (r + 1, r + 1):
        raise ValueError(f"theta must be {(r + 1, r + 1)} for {r} spatial dims, got {t.shape}")
    if not normalized:
        t = _normalize_matrix(src_shape, False, zero_centered) @ t @ np.linalg.inv(_normalize_matrix(dst_shape, False, zero_centered))
    if not reverse_indexing:
        rev = list(range(r - 1, -1, -1)) + [r]
        t = t[rev][:, rev]
    return _unnormalize(src_shape, align_corners) @ t @ _base_grid(dst_shape, align_corners)
