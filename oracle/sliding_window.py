"""numpy restatement of monai/inferers/utils.py::sliding_window_inference (lines 42-321) and its helpers.

TEST INFRASTRUCTURE (see oracle/__init__.py).  N-D, fp32 (or the input dtype), single-output predictors plus
tuple / dict outputs at other resolutions -- everything the reference's tests/inferers/test_sliding_window_inference.py pins.
"""
from __future__ import annotations

import itertools
import math
from typing import Callable, Sequence

import numpy as np


def fall_back_tuple(user, default):
    """monai/utils/misc.py fall_back_tuple: non-positive / None -> default."""
    if not isinstance(user, (list, tuple, np.ndarray)):
        user = (user,) * len(default)
    if len(user) != len(default):
        raise ValueError(f"Sequence must have length {len(default)}, got {len(user)}.")
    return tuple(int(u) if (u is not None and u > 0) else int(d) for u, d in zip(user, default))


def get_valid_patch_size(image_size, patch_size):
    """monai/data/utils.py:343-354."""
    ndim = len(image_size)
    if not isinstance(patch_size, (list, tuple, np.ndarray)):
        patch_size = (patch_size,)
    patch = (tuple(patch_size) + (0,) * ndim)[:ndim]
    return tuple(min(ms, ps or ms) for ms, ps in zip(image_size, patch))


def get_scan_interval(image_size, roi_size, overlap):
    """monai/inferers/utils.py:363-384."""
    out = []
    for i, r, o in zip(image_size, roi_size, overlap):
        if r == i:
            out.append(int(r))
        else:
            interval = int(r * (1 - o))
            out.append(interval if interval > 0 else 1)
    return tuple(out)


def dense_patch_starts(image_size, patch_size, scan_interval):
    """monai/data/utils.py:166-206 (start positions; windows are their product in "ij" order)."""
    patch_size = get_valid_patch_size(image_size, patch_size)
    starts = []
    for i in range(len(image_size)):
        if scan_interval[i] == 0:
            n = 1
        else:
            num = int(math.ceil(float(image_size[i]) / scan_interval[i]))
            n = 1
            for d in range(num):
                if d * scan_interval[i] + patch_size[i] >= image_size[i]:
                    n = d + 1
                    break
        ax = []
        for idx in range(n):
            s = idx * scan_interval[i]
            s -= max(s + patch_size[i] - image_size[i], 0)
            ax.append(s)
        starts.append(ax)
    return starts


def dense_patch_slices(image_size, patch_size, scan_interval):
    patch_size = get_valid_patch_size(image_size, patch_size)
    starts = dense_patch_starts(image_size, patch_size, scan_interval)
    return [tuple(slice(s, s + p) for s, p in zip(st, patch_size)) for st in itertools.product(*starts)]


def compute_importance_map(patch_size, mode="constant", sigma_scale=0.125):
    """monai/data/utils.py:1084-1134, float32 arithmetic in the same order."""
    import torch  # tiny vectors only: torch's fp32 exp is what the reference evaluates, so the map is bit-identical

    if mode == "constant":
        m = np.ones(patch_size, dtype=np.float32)
    elif mode == "gaussian":
        if not isinstance(sigma_scale, (list, tuple)):
            sigma_scale = (sigma_scale,) * len(patch_size)
        m = None
        for i, (p, s) in enumerate(zip(patch_size, sigma_scale)):
            sigma = p * s
            x = torch.arange(start=-(p - 1) / 2.0, end=(p - 1) / 2.0 + 1, dtype=torch.float)
            g = torch.exp(x**2 / (-2 * sigma**2)).numpy()
            m = g if m is None else (m[..., None] * g[(None,) * i]).astype(np.float32)
    else:
        raise ValueError(f"Unsupported mode: {mode}")
    lo = max(float(m.min()), 1e-3)
    return np.maximum(m, np.float32(lo)).astype(np.float32)


def _nearest_exact_resize(w: np.ndarray, out_shape) -> np.ndarray:
    """F.interpolate(mode="nearest-exact"): src index = floor((dst + 0.5) * in/out)."""
    idx = [np.minimum(np.floor((np.arange(o) + 0.5) * (i / o)).astype(np.int64), i - 1) for i, o in zip(w.shape, out_shape)]
    return w[np.ix_(*idx)]


def sliding_window_inference(
    inputs: np.ndarray,
    roi_size,
    sw_batch_size: int,
    predictor: Callable,
    overlap=0.25,
    mode="constant",
    sigma_scale=0.125,
    padding_mode="constant",
    cval=0.0,
    roi_weight_map=None,
):
    """inputs [B,C,*spatial] numpy; predictor maps [n,C,*roi] -> array | tuple | dict of arrays."""
    nsd = inputs.ndim - 2
    if not isinstance(overlap, (list, tuple)):
        overlap = (overlap,) * nsd
    for o in overlap:
        if o < 0 or o >= 1:
            raise ValueError(f"overlap must be >= 0 and < 1, got {overlap}.")
    dtype = inputs.dtype
    batch_size = inputs.shape[0]
    image_size_ = list(inputs.shape[2:])
    roi_size = fall_back_tuple(roi_size, image_size_)
    image_size = tuple(max(i, r) for i, r in zip(image_size_, roi_size))
    pad_width = [(0, 0), (0, 0)]
    pad_size = []
    for k in range(inputs.ndim - 1, 1, -1):  # utils.py:165-168 (last axis first)
        diff = max(roi_size[k - 2] - inputs.shape[k], 0)
        half = diff // 2
        pad_size.extend([half, diff - half])
    for ax in range(nsd):
        j = nsd - 1 - ax
        pad_width.append((pad_size[2 * j], pad_size[2 * j + 1]))
    if any(pad_size):
        np_mode = {"constant": "constant", "reflect": "reflect", "replicate": "edge", "circular": "wrap"}[padding_mode]
        kw = {"constant_values": cval} if np_mode == "constant" else {}
        inputs = np.pad(inputs, pad_width, mode=np_mode, **kw)
    scan_interval = get_scan_interval(image_size, roi_size, overlap)
    slices = dense_patch_slices(image_size, roi_size, scan_interval)
    num_win = len(slices)
    total = num_win * batch_size
    valid = get_valid_patch_size(image_size, roi_size)
    if valid == tuple(roi_size) and roi_weight_map is not None:
        imp = np.asarray(roi_weight_map)
    else:
        imp = compute_importance_map(valid, mode, sigma_scale)
    imp = imp.astype(dtype)

    outs, counts, keys = [], [], None
    for g in range(0, total, sw_batch_size):
        rng = range(g, min(g + sw_batch_size, total))
        win = np.concatenate([inputs[(slice(i // num_win, i // num_win + 1), slice(None)) + slices[i % num_win]] for i in rng])
        seg = predictor(win)
        if isinstance(seg, dict):
            keys = sorted(seg.keys())
            seg = tuple(seg[k] for k in keys)
        elif not isinstance(seg, (tuple, list)):
            seg = (seg,)
        for ss, sg in enumerate(seg):
            sg = np.asarray(sg).astype(dtype)
            seg_shape = sg.shape[2:]
            z = None
            w = imp
            if tuple(seg_shape) != tuple(roi_size):
                z = [o / float(i) for o, i in zip(seg_shape, roi_size)]
                w = _nearest_exact_resize(imp, seg_shape)
            if len(outs) <= ss:
                oshape = [batch_size, sg.shape[1]] + ([int(i * zz) for i, zz in zip(image_size, z)] if z else list(image_size))
                outs.append(np.zeros(oshape, dtype=dtype))
                cm = np.zeros([1, 1] + oshape[2:], dtype=dtype)
                for s in slices:
                    if z is not None:
                        s = tuple(slice(int(si.start * zz), int(si.stop * zz)) for si, zz in zip(s, z))
                    cm[(slice(None), slice(None)) + s] += w
                counts.append(cm)
            sg = sg * w[None, None]
            for i, p in zip(rng, sg):
                s = slices[i % num_win]
                if z is not None:
                    s = tuple(slice(int(si.start * zz), int(si.stop * zz)) for si, zz in zip(s, z))
                outs[ss][(i // num_win, slice(None)) + s] += p
    for ss in range(len(outs)):
        outs[ss] = outs[ss] / counts[ss]
    if any(pad_size):
        for ss, o in enumerate(outs):
            zoom = [a / b for a, b in zip(o.shape[2:], roi_size)]  # utils.py:304
            sl = []
            for sp in range(nsd):
                si = nsd - sp - 1
                sl.insert(0, slice(int(round(pad_size[sp * 2] * zoom[si])), int(round((pad_size[sp * 2] + image_size_[si]) * zoom[si]))))
            outs[ss] = o[(slice(None), slice(None)) + tuple(sl)]
    if keys is not None:
        return dict(zip(keys, outs))
    return outs[0] if len(outs) == 1 else tuple(outs)
