"""Depth-sharded sliding-window inference over the GPUs of one node (BASELINE.json configs[4]).

The reference has no single-volume sharding; its result on one device is the oracle.  Here the
window set is partitioned by *start layer* along the first spatial axis ("depth"): rank k runs the windows of a
contiguous group of layers and accumulates fp32 numerators for its slab only.  The only coupling between ranks is
the overlap-add at slab boundaries, so exactly one exchange step follows: every rank sends each peer the part of its
slab that the peer *owns* (NCCL P2P over NVLink; gloo in the CPU tests), adds what it receives, normalises its owned
rows with the analytic count map (no count-map exchange) and the slabs are all-gathered.

`make_shard_plan` / `exchange_partials` are device-agnostic and are unit-tested with gloo on CPU (world_size 2, 3).
"""
from __future__ import annotations

import math

from collections.abc import Callable, Sequence
from dataclasses import dataclass
from typing import Any

import numpy as np
import torch
import torch.distributed as dist

__all__ = ["ShardPlan", "make_shard_plan", "exchange_partials", "allgather_owned", "ShardedSlidingWindowInferer"]


@dataclass
class ShardPlan:
    world: int
    win_range: list[tuple[int, int]]   # per rank: [first, last) flat window index within one batch item (depth-major order)
    slab: list[tuple[int, int]]        # per rank: depth rows covered by its windows [lo, hi)
    owned: list[tuple[int, int]]       # per rank: depth rows it finalises [lo, hi); a partition of [0, D)


def make_shard_plan(starts_d: Sequence[int], roi_d: int, D: int, world: int, per_layer: int = 1) -> ShardPlan:
    """Balanced partition of the depth-major window list: rank k runs flat windows [k*n/world, (k+1)*n/world) where
    n = len(starts_d) * per_layer (per_layer = windows per depth start-layer).  A layer may be split between two
    ranks, so the load imbalance is at most one window (8 GPUs on 2100 windows -> 262/263).
    Ownership cuts are monotone depth positions near the window-range boundaries; correctness does not depend on
    where they sit because `exchange_partials` routes every slab/owner intersection."""
    nl = len(starts_d)
    if world < 1:
        raise ValueError("world must be >= 1")
    n = nl * per_layer
    bounds = [(k * n) // world for k in range(world + 1)]
    win_range = [(bounds[k], bounds[k + 1]) for k in range(world)]
    slab = []
    for a, b in win_range:
        slab.append((int(starts_d[a // per_layer]), int(starts_d[(b - 1) // per_layer]) + roi_d) if b > a else (0, 0))
    cuts = [0]
    for k in range(1, world):
        b = bounds[k]
        if b <= 0:
            cuts.append(0)
        elif b >= n:
            cuts.append(D)
        else:
            layer, frac = b // per_layer, (b % per_layer) / per_layer
            nxt = starts_d[layer + 1] if layer + 1 < nl else starts_d[layer] + roi_d
            pos = starts_d[layer] + frac * (nxt - starts_d[layer]) + (roi_d // 2 if frac == 0 else roi_d / 2)
            cuts.append(int(min(max(round(pos), cuts[-1]), D)))
    cuts.append(D)
    owned = [(cuts[k], cuts[k + 1]) for k in range(world)]
    return ShardPlan(world, win_range, slab, owned)


def _intersect(a, b):
    lo, hi = max(a[0], b[0]), min(a[1], b[1])
    return (lo, hi) if hi > lo else None


def _planes(t: torch.Tensor, lo: int, hi: int):
    """The (batch, channel) pieces of t[:, :, lo:hi]: each is a contiguous block of rows, so it can be sent / received in place."""
    return [t[b, c, lo:hi] for b in range(t.shape[0]) for c in range(t.shape[1])]


def _add_inplace(dst: torch.Tensor, src: torch.Tensor) -> None:
    if dst.is_cuda:
        from .. import _kernels as K

        K.add_f32(dst, src)
    else:   # gloo unit tests of the host logic run on CPU tensors
        dst += src


def exchange_partials(acc: torch.Tensor, plan: ShardPlan, rank: int, group=None) -> None:
    """acc [B, C, D, H, W] fp32 holds this rank's partial numerators on its slab rows (zeros elsewhere).  After the
    call the rows this rank OWNS hold the complete sums.  Sends: my slab ∩ peer's owned rows; receives the converse.
    Every (batch, channel) plane of a row range is contiguous, so the sends go straight out of the accumulator (no staging
    copy); ONE grouped P2P call carries all transfers; received planes are added by a CUDA kernel."""
    world = plan.world
    if world == 1:
        return
    ops, recv_bufs = [], []
    for peer in range(world):
        if peer == rank:
            continue
        s = _intersect(plan.slab[rank], plan.owned[peer])
        if s is not None:
            for piece in _planes(acc, s[0], s[1]):
                ops.append(dist.P2POp(dist.isend, piece, peer, group=group))
        r = _intersect(plan.slab[peer], plan.owned[rank])
        if r is not None:
            for piece in _planes(acc, r[0], r[1]):
                buf = torch.empty_like(piece)
                recv_bufs.append((piece, buf))
                ops.append(dist.P2POp(dist.irecv, buf, peer, group=group))
    if ops:
        for req in dist.batch_isend_irecv(ops):
            req.wait()
    for piece, buf in recv_bufs:
        _add_inplace(piece, buf)


def allgather_owned(out: torch.Tensor, plan: ShardPlan, rank: int, group=None) -> None:
    """Every rank holds valid result rows plan.owned[rank] of out [B, C, D, H, W]; after the call every rank holds all rows.
    One grouped P2P call (an all-gather with uneven, in-place pieces): each rank sends the contiguous (batch, channel) planes
    of its owned rows to every peer and receives the peers' planes directly into `out`."""
    world = plan.world
    if world == 1:
        return
    ops = []
    lo, hi = plan.owned[rank]
    mine = _planes(out, lo, hi) if hi > lo else []
    for peer in range(world):
        if peer == rank:
            continue
        for piece in mine:
            ops.append(dist.P2POp(dist.isend, piece, peer, group=group))
        plo, phi = plan.owned[peer]
        if phi > plo:
            for piece in _planes(out, plo, phi):
                ops.append(dist.P2POp(dist.irecv, piece, peer, group=group))
    if ops:
        for req in dist.batch_isend_irecv(ops):
            req.wait()


class ShardedSlidingWindowInferer:
    """`SlidingWindowInferer` whose windows are partitioned over the ranks of the default process group.

    Every rank passes the SAME full volume (or a tensor whose slab rows are valid) and receives the SAME full result.
    Constructor arguments match `SlidingWindowInferer` (roi_size, sw_batch_size, overlap, mode, sigma_scale).
    """

    def __init__(self, roi_size, sw_batch_size: int = 1, overlap=0.25, mode="constant", sigma_scale=0.125, gather: bool = True):
        self.roi_size, self.sw_batch_size, self.overlap, self.mode, self.sigma_scale, self.gather = roi_size, sw_batch_size, overlap, mode, sigma_scale, gather

    def plan(self, spatial_shape: Sequence[int], world: int | None = None) -> ShardPlan:
        """The shard plan `__call__` uses for a volume of this spatial shape: `plan.slab[rank]` are the depth rows of the
        INPUT a rank reads (callers may upload only those), `plan.owned[rank]` the rows of the result it finalises (with
        `gather=False` only those rows of the returned tensor are valid)."""
        from ..data.utils import dense_patch_starts
        from ..inferers.utils import _ensure_tuple_rep, _fall_back_tuple, _get_scan_interval

        if world is None:
            world = dist.get_world_size() if dist.is_initialized() else 1
        D, H, W = (int(s) for s in spatial_shape)
        roi = _fall_back_tuple(self.roi_size, (D, H, W))
        interval = _get_scan_interval((D, H, W), roi, 3, _ensure_tuple_rep(self.overlap, 3))
        starts = dense_patch_starts((D, H, W), roi, interval)
        return make_shard_plan(starts[0], roi[0], D, world, per_layer=len(starts[1]) * len(starts[2]))

    def __call__(self, inputs: torch.Tensor, network: Callable[..., torch.Tensor], *args: Any, **kwargs: Any) -> torch.Tensor:
        from .. import _kernels as K
        from ..data.utils import dense_patch_starts, importance_factors
        from ..inferers.utils import _RESIDENT_BYTES, _ensure_tuple_rep, _fall_back_tuple, _get_scan_interval

        if inputs.dim() != 5 or not inputs.is_cuda:
            raise RuntimeError("ShardedSlidingWindowInferer takes CUDA tensors [B, C, D, H, W]")
        rank = dist.get_rank() if dist.is_initialized() else 0
        world = dist.get_world_size() if dist.is_initialized() else 1
        B, _, D, H, W = inputs.shape
        roi = _fall_back_tuple(self.roi_size, (D, H, W))
        if any(r > s for r, s in zip(roi, (D, H, W))):
            raise NotImplementedError("sharded inference expects the volume to be at least one roi in every axis")
        overlap = _ensure_tuple_rep(self.overlap, 3)
        interval = _get_scan_interval((D, H, W), roi, 3, overlap)
        starts = dense_patch_starts((D, H, W), roi, interval)
        nh, nw = len(starts[1]), len(starts[2])
        plan = make_shard_plan(starts[0], roi[0], D, world, per_layer=nh * nw)
        wa, wb = plan.win_range[rank]
        num_win = len(starts[0]) * nh * nw
        if num_win < world:   # every rank computes the same numbers: all of them raise, none enters a collective
            raise ValueError(f"sharded inference over {world} ranks needs at least {world} windows, the volume has {num_win}")
        dev = inputs.device
        mode_s = str(getattr(self.mode, "value", self.mode)).lower()
        factors, clamp = importance_factors(roi, mode_s, self.sigma_scale)
        factors = [f.to(dev) for f in factors]
        starts_t = [torch.tensor(s, dtype=torch.int32, device=dev) for s in starts]
        starts_t[2]._align = math.gcd(8, *[int(v) for v in starts[2]])
        starts_t[2]._max_cover = max(max(sum(1 for s0 in ax if s0 <= v < s0 + r) for v in ax) for ax, r in zip(starts, roi))
        x = inputs.detach()
        x = x.as_subclass(torch.Tensor) if type(x) is not torch.Tensor else x
        acc = None
        out_c = None
        store = None      # this rank's predictions stay resident (fp16/fp32 as produced) and are blended in one pass
        per_layer = nh * nw
        for b in range(B):
            ids = list(range(b * num_win + wa, b * num_win + wb))
            tab = torch.tensor([(b, starts[0][(i % num_win) // per_layer], starts[1][((i % num_win) // nw) % nh], starts[2][(i % num_win) % nw]) for i in ids],
                               dtype=torch.int32, device=dev).reshape(-1, 4)
            held, held_first = 0, ids[0] if ids else 0

            def flush():
                # numerators of the held windows are added to the fp32 accumulators over the depth rows they touch
                # (row range from the host-side start table: no device synchronisation on the way)
                lo_l, hi_l = (held_first - b * num_win) // per_layer, (held_first + held - 1 - b * num_win) // per_layer
                K.sw_blend(1, store[:held], held_first, held_first + held, (B, out_c, D, H, W), roi, starts_t, factors, clamp, None, acc,
                           box=(int(starts[0][lo_l]), int(starts[0][hi_l]) + roi[0], 0, 0))

            for g in range(0, len(ids), self.sw_batch_size):
                win = K.sw_gather(x, tab[g : g + self.sw_batch_size], roi, w_align=math.gcd(16, *[int(v) for v in starts[2]]))
                seg = network(win, *args, **kwargs)
                if not isinstance(seg, torch.Tensor) or tuple(seg.shape[2:]) != tuple(roi):
                    raise NotImplementedError("sharded inference supports single-tensor predictors at the input resolution")
                seg = seg.detach()
                if acc is None:
                    out_c = seg.shape[1]
                    acc = torch.zeros((B, out_c, D, H, W), device=dev, dtype=torch.float32)
                    per_win = out_c * int(np.prod(roi)) * seg.element_size()
                    cap = max(self.sw_batch_size, min(len(ids), _RESIDENT_BYTES // max(1, per_win)))
                    store = torch.empty((cap, out_c, *roi), device=dev, dtype=seg.dtype)
                n = seg.shape[0]
                if held + n > store.shape[0]:
                    flush()
                    held_first, held = ids[g], 0
                store[held : held + n].copy_(seg)
                held += n
            if held:
                flush()
        if acc is None:  # unreachable: the window count was validated against the world size before any collective
            raise RuntimeError("rank without windows")
        exchange_partials(acc, plan, rank)
        o_lo, o_hi = plan.owned[rank]
        out = torch.empty((B, out_c, D, H, W), device=dev, dtype=inputs.dtype if inputs.dtype in (torch.float16, torch.float32) else torch.float32)
        K.sw_blend(2, None, 0, B * num_win, (B, out_c, D, H, W), roi, starts_t, factors, clamp, None, out, acc=acc, box=(o_lo, o_hi, 0, 0))
        if world > 1 and self.gather:
            allgather_owned(out, plan, rank)
        return out
