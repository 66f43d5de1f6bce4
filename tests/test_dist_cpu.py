"""World-size-2/3 gloo tests (CPU) of the depth-sharding host logic: partition, ownership cuts and the exchange step.

Per-rank partial numerators are produced by a numpy restatement of the weighted scatter-add restricted to the rank's
window layers; after `exchange_partials` over gloo + the analytic count, the stitched volume must equal the
single-process oracle (`oracle.sliding_window`)."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from monai_b200.parallel import exchange_partials, make_shard_plan
from monai_b200.parallel.sharded import allgather_owned
from oracle import sliding_window as osw


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _pred(x):
    return np.concatenate([x * 2.0 + 1.0, np.tanh(x)], axis=1).astype(np.float32)


def _worker(rank, world, port, shape, roi, overlap, ret):
    os.environ["MASTER_ADDR"], os.environ["MASTER_PORT"] = "127.0.0.1", str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        x = np.random.default_rng(0).standard_normal((1, 1, *shape)).astype(np.float32)
        interval = osw.get_scan_interval(shape, roi, (overlap,) * 3)
        starts = osw.dense_patch_starts(shape, roi, interval)
        nh, nw = len(starts[1]), len(starts[2])
        plan = make_shard_plan(starts[0], roi[0], shape[0], world, per_layer=nh * nw)
        imp = osw.compute_importance_map(roi, "gaussian", 0.125)
        wa, wb = plan.win_range[rank]
        acc = np.zeros((1, 2, *shape), dtype=np.float32)
        for i in range(wa, wb):
            sd, sh, sw = starts[0][i // (nh * nw)], starts[1][(i // nw) % nh], starts[2][i % nw]
            sl = (slice(None), slice(None), slice(sd, sd + roi[0]), slice(sh, sh + roi[1]), slice(sw, sw + roi[2]))
            acc[sl] += _pred(x[sl]) * imp
        # rows outside the slab must be untouched
        lo, hi = plan.slab[rank]
        assert not acc[:, :, :lo].any() and not acc[:, :, hi:].any()
        t = torch.from_numpy(acc)
        exchange_partials(t, plan, rank)
        cnt = np.zeros((1, 1, *shape), dtype=np.float32)
        for sd in starts[0]:
            for sh in starts[1]:
                for sw in starts[2]:
                    cnt[:, :, sd : sd + roi[0], sh : sh + roi[1], sw : sw + roi[2]] += imp
        o_lo, o_hi = plan.owned[rank]
        mine = t.numpy()[:, :, o_lo:o_hi] / cnt[:, :, o_lo:o_hi]
        # the owned rows are then shared with every rank in ONE grouped exchange (uneven, in-place all-gather)
        res = torch.full((1, 2, *shape), float("nan"))
        res[:, :, o_lo:o_hi] = torch.from_numpy(mine)
        allgather_owned(res, plan, rank)
        ret[rank] = (o_lo, o_hi, mine, res.numpy())
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("world,shape,roi,overlap", [(2, (40, 20, 18), (16, 16, 16), 0.5), (3, (50, 16, 16), (16, 16, 16), 0.75), (2, (33, 17, 16), (16, 16, 16), 0.25)])
def test_sharded_exchange_matches_single_process_oracle(world, shape, roi, overlap):
    mgr = mp.Manager()
    ret = mgr.dict()
    mp.spawn(_worker, args=(world, _free_port(), shape, roi, overlap, ret), nprocs=world, join=True)
    x = np.random.default_rng(0).standard_normal((1, 1, *shape)).astype(np.float32)
    want = osw.sliding_window_inference(x, roi, 4, _pred, overlap, "gaussian")
    got = np.zeros_like(want)
    covered = np.zeros(shape[0], dtype=int)
    for r in range(world):
        lo, hi, part, gathered = ret[r]
        got[:, :, lo:hi] = part
        covered[lo:hi] += 1
        np.testing.assert_allclose(gathered, want, rtol=1e-5, atol=1e-6, err_msg=f"all-gathered result on rank {r}")
    assert (covered == 1).all()  # owned rows partition the depth axis
    np.testing.assert_allclose(got, want, rtol=1e-5, atol=1e-6)


def test_shard_plan_balance_and_coverage():
    # config C5: 1024 depth rows, roi 96, interval 48 -> 21 layers x 100 windows; 8 ranks -> 262/263 windows each
    starts = list(range(0, 913, 48)) + [928]
    plan = make_shard_plan(starts, 96, 1024, 8, per_layer=100)
    sizes = [b - a for a, b in plan.win_range]
    assert sum(sizes) == 2100 and max(sizes) - min(sizes) <= 1
    assert plan.owned[0][0] == 0 and plan.owned[-1][1] == 1024
    for r in range(7):
        assert plan.owned[r][1] == plan.owned[r + 1][0] and plan.owned[r][1] >= plan.owned[r][0]
    one = make_shard_plan(starts, 96, 1024, 1, per_layer=100)
    assert one.owned == [(0, 1024)] and one.win_range == [(0, 2100)]
    # more ranks than windows: empty ranks own nothing and the plan stays consistent
    tiny = make_shard_plan([0], 16, 16, 3, per_layer=2)
    assert sum(b - a for a, b in tiny.win_range) == 2


def test_sharded_inferer_plan_helper_matches_make_shard_plan():
    """`plan()` tells a caller which input rows a rank reads and which result rows it owns (used by bench.py's e2e arm)."""
    from monai_b200.data.utils import dense_patch_starts
    from monai_b200.parallel import ShardedSlidingWindowInferer, make_shard_plan

    inf = ShardedSlidingWindowInferer((96, 96, 96), 8, 0.5, "gaussian")
    for world in (1, 2, 4, 8):
        p = inf.plan((512, 512, 512), world)
        starts = dense_patch_starts((512, 512, 512), (96, 96, 96), (48, 48, 48))
        q = make_shard_plan(starts[0], 96, 512, world, per_layer=len(starts[1]) * len(starts[2]))
        assert p.win_range == q.win_range and p.slab == q.slab and p.owned == q.owned
        assert p.owned[0][0] == 0 and p.owned[-1][1] == 512 and all(a[1] == b[0] for a, b in zip(p.owned, p.owned[1:]))
        assert sum(b - a for a, b in p.win_range) == 1000
        for (s0, s1), (o0, o1) in zip(p.slab, p.owned):
            assert s0 <= o0 < o1 <= s1 or o1 <= o0   # a rank's owned rows lie inside the rows its windows cover
