"""2-GPU NCCL test of the depth-sharded sliding-window inferer (skipped when fewer than 2 GPUs are visible)."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

pytestmark = pytest.mark.gpu


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _pred(x):
    return torch.cat([x * 2.0 + 1.0, torch.tanh(x)], dim=1)


def _worker(rank, world, port, ret):
    os.environ["MASTER_ADDR"], os.environ["MASTER_PORT"] = "127.0.0.1", str(port)
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", rank))
    try:
        from monai_b200.inferers import SlidingWindowInferer
        from monai_b200.parallel import ShardedSlidingWindowInferer

        x = torch.randn(1, 1, 80, 40, 48, generator=torch.Generator().manual_seed(0)).cuda()
        single = SlidingWindowInferer((32, 32, 32), 4, 0.5, "gaussian")(x, _pred)
        sharded = ShardedSlidingWindowInferer((32, 32, 32), 4, 0.5, "gaussian")(x, _pred)
        ret[rank] = float((single - sharded).abs().max())
    finally:
        dist.destroy_process_group()


def test_sharded_inferer_matches_single_gpu():
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    mgr = mp.Manager()
    ret = mgr.dict()
    mp.spawn(_worker, args=(2, _free_port(), ret), nprocs=2, join=True)
    assert len(ret) == 2 and max(ret.values()) < 1e-5, dict(ret)
