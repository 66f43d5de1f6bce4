"""Host-side coordinate helpers of the resampling path with the reference's names and call contracts
(monai/networks/utils.py:243-326): `normalize_transform` and `to_norm_affine`.  Small float64 matrices computed on the host; the CUDA
resampler never sees them separately -- monai_b200.transforms.utils folds them into one output-index -> source-index matrix."""
from __future__ import annotations

from typing import Sequence

import numpy as np
import torch

from ..transforms.utils import _normalize_matrix

__all__ = ["normalize_transform", "to_norm_affine"]


def normalize_transform(shape, device=None, dtype=None, align_corners: bool = False, zero_centered: bool = False) -> torch.Tensor:
    """The 1 x (d+1) x (d+1) affine that takes voxel indices of an image of spatial `shape` to [-1, 1]
    (monai/networks/utils.py:243-286).  Source ranges: align_corners=False / zero_centered=False: [-0.5, d-0.5]; True / False: [0, d-1];
    False / True: [-(d-1)/2, (d-1)/2]; True / True: [-d/2, d/2]."""
    shape = [float(s) for s in (shape.tolist() if isinstance(shape, (torch.Tensor, np.ndarray)) else shape)]
    m = _normalize_matrix(shape, bool(align_corners), bool(zero_centered))
    return torch.as_tensor(m, dtype=torch.float64)[None].to(device=device, dtype=dtype if dtype is not None else torch.float64)


def to_norm_affine(affine: torch.Tensor, src_size: Sequence[int], dst_size: Sequence[int], align_corners: bool = False,
                   zero_centered: bool = False) -> torch.Tensor:
    """`affine` (N x d x d, voxel coordinates) expressed for normalised coordinates:
    normalize_transform(src) @ affine @ inv(normalize_transform(dst)) (monai/networks/utils.py:289-326; same exceptions)."""
    if not isinstance(affine, torch.Tensor):
        raise TypeError(f"affine must be a torch.Tensor but is {type(affine).__name__}.")
    if affine.ndimension() != 3 or affine.shape[1] != affine.shape[2]:
        raise ValueError(f"affine must be Nxdxd, got {tuple(affine.shape)}.")
    sr = affine.shape[1] - 1
    if sr != len(src_size) or sr != len(dst_size):
        raise ValueError(f"affine suggests {sr}D, got src={len(src_size)}D, dst={len(dst_size)}D.")
    src_xform = normalize_transform(src_size, affine.device, affine.dtype, align_corners, zero_centered)
    dst_inv = np.linalg.inv(_normalize_matrix([float(s) for s in dst_size], bool(align_corners), bool(zero_centered)))
    return src_xform @ affine @ torch.as_tensor(dst_inv, dtype=affine.dtype, device=affine.device)
