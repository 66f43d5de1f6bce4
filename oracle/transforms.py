"""torch-CPU restatement of the reference's spatial-transform path -- TEST INFRASTRUCTURE (see oracle/__init__.py).

The oracle follows the reference's *literal* algorithm (dense coordinate grid + F.grid_sample in float64, separable
F.conv3d for the Gaussian), which is deliberately a different formulation from the product's closed-form 3x4 matrix:
  spatial_resample  -- monai/transforms/spatial/functional.py:68-184 + networks/layers/spatial_transforms.py:502-592
  spacing           -- monai/transforms/spatial/array.py:437-543 (+ data/utils.py affine helpers)
  rand_affine       -- monai/transforms/spatial/array.py:2476-2548, 1858-1915, 1758-1783, 2061-2116; transforms/utils.py:758-831
  gaussian_smooth   -- monai/transforms/intensity/array.py:1610-1622; networks/layers/simplelayers.py:170-249, 589-595
"""
from __future__ import annotations

import numpy as np
import torch
import torch.nn.functional as F

AFFINE_TOL = 1e-3


def to_affine_nd(r, affine):
    a = np.asarray(affine, dtype=np.float64).copy()
    new = np.eye(int(r) + 1) if np.ndim(r) == 0 else np.array(r, dtype=np.float64, copy=True)
    d = max(min(len(new) - 1, len(a) - 1), 1)
    new[:d, :d] = a[:d, :d]
    if d > 1:
        new[:d, -1] = a[:d, -1]
    return new


def affine_to_spacing(affine, r=3):
    a = np.asarray(affine, dtype=np.float64)[:r, :r]
    s = np.sqrt(np.sum(a * a, axis=0))
    s[s == 0] = 1.0
    return s


def zoom_affine(affine, scale, diagonal=True):
    a = np.array(affine, dtype=float, copy=True)
    d = len(a) - 1
    norm = affine_to_spacing(a, d)
    sc = np.array(scale, dtype=float, copy=True)
    if len(sc) < d:
        sc = np.append(sc, norm[len(sc):])
    sc = sc[:d]
    sc = np.asarray([s if s > 0 else n for s, n in zip(sc, norm)])
    sc[sc == 0] = 1.0
    if diagonal:
        return np.diag(np.append(sc, [1.0]))
    rzs = a[:-1, :-1]
    zs = np.linalg.cholesky(rzs.T @ rzs).T
    rot = rzs @ np.linalg.inv(zs)
    out = np.eye(len(a))
    out[:-1, :-1] = rot @ np.diag(np.sign(np.diag(zs)) * np.abs(sc))
    return out


def compute_shape_offset(spatial_shape, in_affine, out_affine, scale_extent=False):
    shape = np.array(spatial_shape, copy=True, dtype=float)
    sr = len(shape)
    ia, oa = to_affine_nd(sr, in_affine), to_affine_nd(sr, out_affine)
    in_coords = [(-0.5, d - 0.5) if scale_extent else (0.0, d - 1.0) for d in shape]
    corners = np.asarray(np.meshgrid(*in_coords, indexing="ij")).reshape((sr, -1))
    corners = np.concatenate((corners, np.ones_like(corners[:1])))
    corners_out = np.linalg.solve(oa, ia) @ corners
    corners = ia @ corners
    all_dist = corners_out[:-1].copy()
    corners_out = corners_out[:-1] / corners_out[-1]
    out_shape = np.round(np.ptp(corners_out, axis=1)) if scale_extent else np.round(np.ptp(corners_out, axis=1) + 1.0)
    offset = None
    for i in range(corners.shape[1]):
        if np.allclose(np.min(all_dist - all_dist[:, i : i + 1], 1), 0.0, rtol=AFFINE_TOL):
            offset = corners[:-1, i]
            break
    if offset is None:
        offset = ia[:-1, :-1] @ (shape / 2.0) + ia[:-1, -1] - oa[:-1, :-1] @ (out_shape / 2.0)
    if scale_extent:
        in_offset = np.append(0.5 * (shape / out_shape - 1.0), 1.0)
        offset = np.abs((ia @ in_offset / in_offset[-1])[:-1]) * np.sign(offset)
    return out_shape.astype(int), offset


def _normalize_transform(shape, align_corners=False):
    shape = torch.as_tensor(shape, dtype=torch.float64)
    norm = shape.clone()
    if align_corners:
        norm[norm <= 1.0] = 2.0
        norm = 2.0 / (norm - 1.0)
        m = torch.diag(torch.cat((norm, torch.ones(1, dtype=torch.float64))))
        m[:-1, -1] = -1.0
    else:
        norm[norm <= 0.0] = 2.0
        norm = 2.0 / norm
        m = torch.diag(torch.cat((norm, torch.ones(1, dtype=torch.float64))))
        m[:-1, -1] = 1.0 / shape - 1.0
    return m


def spatial_resample(img: torch.Tensor, src_affine, dst_affine, spatial_size=None, mode="bilinear", padding_mode="border", align_corners=False):
    """img [C, *spatial] -> (float32 result, xform)."""
    r = min(img.dim() - 1, 3)
    src = to_affine_nd(r, src_affine)
    dst = to_affine_nd(r, dst_affine)
    in_size = np.asarray(img.shape[1 : 1 + r])
    if spatial_size is None:
        spatial_size, _ = compute_shape_offset(in_size, src, dst)
    spatial_size = [int(s) for s in spatial_size]
    xform = np.linalg.solve(src, dst)
    if (np.allclose(src, dst, atol=AFFINE_TOL) or np.allclose(xform, np.eye(r + 1), atol=AFFINE_TOL)) and np.allclose(spatial_size, in_size):
        return img.float(), xform
    theta = torch.as_tensor(xform, dtype=torch.float64)[None]
    src_x = _normalize_transform(in_size, False)
    dst_x = _normalize_transform(spatial_size, False)
    theta = src_x @ theta @ torch.linalg.inv(dst_x)          # to_norm_affine(align_corners=False)
    rev = list(range(r - 1, -1, -1))
    theta2 = theta.clone()
    theta2[:, :r] = theta[:, rev]
    theta3 = theta2.clone()
    theta3[:, :, :r] = theta2[:, :, rev]                     # reverse_indexing=True
    x = img[None].double()
    grid = F.affine_grid(theta3[:, :r], [1, x.shape[1], *spatial_size], align_corners=align_corners)
    out = F.grid_sample(x, grid, mode=mode, padding_mode=padding_mode, align_corners=align_corners)[0]
    return out.float(), xform


def spacing(img: torch.Tensor, affine, pixdim, diagonal=False, mode="bilinear", padding_mode="border", align_corners=False):
    sr = img.dim() - 1
    a = to_affine_nd(sr, affine)
    out_d = np.asarray(pixdim, dtype=np.float64)[:sr].copy()
    if out_d.size < sr:
        out_d = np.append(out_d, [out_d[-1]] * (sr - out_d.size))
    new_affine = zoom_affine(a, out_d, diagonal=diagonal)
    out_shape, offset = compute_shape_offset(img.shape[1:], a, new_affine, False)
    new_affine[:sr, -1] = offset[:sr]
    out, _ = spatial_resample(img, a, new_affine, out_shape, mode, padding_mode, align_corners)
    return out, new_affine


def _create_rotate3(radians):
    t = lambda v: torch.as_tensor(v, dtype=torch.float32)  # noqa: E731
    aff = torch.eye(4)
    s, c = torch.sin(t(radians[0])), torch.cos(t(radians[0]))
    aff[1, 1], aff[1, 2], aff[2, 1], aff[2, 2] = c, -s, s, c
    if len(radians) >= 2:
        s, c = torch.sin(t(radians[1])), torch.cos(t(radians[1]))
        m = torch.eye(4)
        m[0, 0], m[0, 2], m[2, 0], m[2, 2] = c, s, -s, c
        aff = aff @ m
    if len(radians) >= 3:
        s, c = torch.sin(t(radians[2])), torch.cos(t(radians[2]))
        m = torch.eye(4)
        m[0, 0], m[0, 1], m[1, 0], m[1, 1] = c, -s, s, c
        aff = aff @ m
    return aff


def rand_affine_params(seed, rotate_range=(), shear_range=(), translate_range=(), scale_range=(), dict_version=True):
    """RNG draw order of RandAffined (dictionary.py:1135-1151): grid parameters are drawn by rand_affine.randomize()
    and AGAIN by rand_affine_grid(...) (randomize=True); the second draw is the one applied."""
    R = np.random.RandomState(seed)

    def draw(rng, add=0.0):
        out = []
        for f in rng:
            if isinstance(f, (list, tuple)):
                out.append(R.uniform(f[0], f[1]) + add)
            elif f is not None:
                out.append(R.uniform(-f, f) + add)
        return out

    for _ in range(2):
        rot, shear, trans, scale = draw(rotate_range), draw(shear_range), draw(translate_range), draw(scale_range, 1.0)
    return rot, shear, trans, scale


def rand_affine(img: torch.Tensor, seed, rotate_range=(), shear_range=(), translate_range=(), scale_range=(), spatial_size=None,
                mode="bilinear", padding_mode="reflection"):
    """RandAffined(prob=1) on one key: identity grid (create_grid, float32) -> affine @ grid -> Resample (float64).
    3-D images [C, D, H, W] (config C4) and 2-D images [C, H, W] (the 2-D goldens of the reference's unit tests):
    create_rotate / _create_shear / create_translate / create_scale of monai/transforms/utils.py:859-1075 for both ranks."""
    rot, shear, trans, scale = rand_affine_params(seed, rotate_range, shear_range, translate_range, scale_range)
    nd = img.dim() - 1
    if nd not in (2, 3):
        raise NotImplementedError("oracle rand_affine restates the 2-D and 3-D paths")
    if spatial_size is None or isinstance(spatial_size, int):   # None / -1: the image's own size (fall_back_tuple)
        sp = list(img.shape[1:])
    else:
        sp = [int(s) if int(s) > 0 else int(d) for s, d in zip(spatial_size, img.shape[1:])]
    axes = [torch.linspace(-(d - 1.0) / 2.0, (d - 1.0) / 2.0, int(d), dtype=torch.float32) for d in sp]
    coords = torch.meshgrid(*axes, indexing="ij")
    grid = torch.stack([*coords, torch.ones_like(coords[0])])
    n = nd + 1
    affine = torch.eye(n)
    t32 = lambda v: torch.as_tensor(v, dtype=torch.float32)  # noqa: E731
    if rot:
        if nd == 3:
            affine = affine @ _create_rotate3(rot)
        else:
            s, c = torch.sin(t32(rot[0])), torch.cos(t32(rot[0]))
            m = torch.eye(3)
            m[0, 0], m[0, 1], m[1, 0], m[1, 1] = c, -s, s, c
            affine = affine @ m
    if shear:
        m = torch.eye(n)
        if nd == 3:
            c = (list(shear) + [0.0] * 6)[:6]
            m[0, 1], m[0, 2], m[1, 0], m[1, 2], m[2, 0], m[2, 1] = [t32(v) for v in c]
        else:
            c = (list(shear) + [0.0] * 2)[:2]
            m[0, 1], m[1, 0] = t32(c[0]), t32(c[1])
        affine = affine @ m
    if trans:
        m = torch.eye(n)
        for i, a in enumerate(trans[:nd]):
            m[i, nd] = a
        affine = affine @ m
    if scale:
        f = (list(scale) + [1.0] * nd)[:nd]
        affine = affine @ torch.diag(torch.as_tensor(f + [1.0], dtype=torch.float32))
    grid = (affine @ grid.reshape(n, -1)).reshape(n, *sp)
    x = img[None].double()
    g = grid[list(range(nd - 1, -1, -1))].movedim(0, -1)[None].double().clone()   # xyz order for grid_sample
    for i, dim in enumerate(x.shape[:1:-1]):
        g[0, ..., i] *= 2.0 / max(2, dim)
    out = F.grid_sample(x, g, mode=mode, padding_mode=padding_mode, align_corners=False)[0]
    return out.float(), affine


def gaussian_1d_erf(sigma: float, truncated: float = 4.0):
    s = torch.as_tensor(sigma, dtype=torch.float)
    tail = int(max(float(s) * truncated, 0.5) + 0.5)
    x = torch.arange(-tail, tail + 1, dtype=torch.float)
    t = 0.70710678 / torch.abs(s)
    return (0.5 * ((t * (x + 0.5)).erf() - (t * (x - 0.5)).erf())).clamp(min=0)


def gaussian_smooth(img: torch.Tensor, sigma):
    nd = img.dim() - 1
    sig = list(sigma) if isinstance(sigma, (list, tuple)) else [sigma] * nd
    x = img[None].float()
    c = x.shape[1]
    conv = [F.conv1d, F.conv2d, F.conv3d][nd - 1]
    for d in range(nd):  # first spatial axis first (the recursion of _separable_filtering_conv applies d=0 innermost)
        k = gaussian_1d_erf(sig[d])
        shape = [1] * (nd + 2)
        shape[d + 2] = -1
        w = k.reshape(shape).repeat([c, 1] + [1] * nd)
        pad = [0] * nd
        pad[d] = (k.numel() - 1) // 2
        x = conv(x, w, padding=pad, groups=c)
    return x[0]


def activations(img: torch.Tensor, sigmoid: bool = False, softmax: bool = False) -> torch.Tensor:
    """Activations.__call__ (monai/transforms/post/array.py:91-128): float32, sigmoid then softmax over dim 0."""
    if sigmoid and softmax:
        raise ValueError("Incompatible values: sigmoid=True and softmax=True.")
    t = img.float()
    if sigmoid:
        t = torch.sigmoid(t)
    if softmax:
        t = torch.softmax(t, dim=0)
    return t


def as_discrete(img: torch.Tensor, argmax: bool = False, to_onehot: int | None = None, threshold: float | None = None,
                rounding: str | None = None) -> torch.Tensor:
    """AsDiscrete.__call__ (monai/transforms/post/array.py:190-251): argmax(dim 0, keepdim), one-hot (`one_hot`,
    monai/networks/utils.py:170-230: scatter of ones along dim 0), `>= threshold`, torch.round; float32 result."""
    t = img
    if argmax:
        t = torch.argmax(t, dim=0, keepdim=True)
    if to_onehot is not None:
        if t.shape[0] != 1:
            raise AssertionError("labels should have a channel with length equal to one.")
        oh = torch.zeros((to_onehot, *t.shape[1:]), dtype=torch.float)
        t = oh.scatter_(0, t.long(), 1.0)
    if threshold is not None:
        t = t >= threshold
    if rounding is not None:
        t = torch.round(t)
    return t.float()


def affine_transform(src: torch.Tensor, theta: torch.Tensor, spatial_size=None, normalized: bool = False, mode: str = "bilinear",
                     padding_mode: str = "zeros", align_corners: bool = True, reverse_indexing: bool = True, zero_centered: bool = False):
    """AffineTransform.forward (monai/networks/layers/spatial_transforms.py:502-592) restated: theta padded to homogeneous form,
    to_norm_affine (networks/utils.py:289-326) when it is not normalised, the (i, j, k) -> (x, y, z) flip, F.affine_grid + F.grid_sample."""
    if theta.dim() == 2:
        theta = theta[None]
    theta = theta.clone()
    sr = src.dim() - 2
    if tuple(theta.shape[1:]) in ((2, 3), (3, 4)):
        pad = torch.zeros((theta.shape[0], 1, sr + 1), dtype=theta.dtype)
        pad[:, 0, -1] = 1
        theta = torch.cat([theta, pad], dim=1)
    src_size = tuple(src.shape)
    dst_size = src_size if spatial_size is None else src_size[:2] + tuple(spatial_size)
    if not normalized:
        def norm(shape):   # normalize_transform(align_corners=False, zero_centered)
            s = torch.tensor(shape, dtype=torch.float64)
            n = s.clone()
            n[n <= 0.0] = 2.0
            m = torch.diag(torch.cat((2.0 / (n - 1.0 if zero_centered else n), torch.ones(1, dtype=torch.float64))))
            if not zero_centered:
                m[:-1, -1] = 1.0 / s - 1.0
            return m.to(theta.dtype)[None]
        theta = norm(src_size[2:]) @ theta @ torch.linalg.inv(norm(dst_size[2:]))
    if reverse_indexing:
        rev = list(range(sr - 1, -1, -1))
        theta[:, :sr] = theta[:, rev]
        theta[:, :, :sr] = theta[:, :, rev]
    if theta.shape[0] == 1 and src_size[0] > 1:
        theta = theta.repeat(src_size[0], 1, 1)
    grid = F.affine_grid(theta[:, :sr], list(dst_size), align_corners=align_corners)
    return F.grid_sample(src.contiguous(), grid, mode=mode, padding_mode=padding_mode, align_corners=align_corners)
