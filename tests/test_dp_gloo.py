"""World-size-2 gloo test of the data-parallel gradient reducer (host logic, CPU): bucket
averaging with a per-rank un-scale factor, skipping parameters that received no gradient
(SURVEY.md fact 7), and rank-0 broadcast of the initial parameters."""
import os
import sys

import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _worker(rank, world, port, ret):
    sys.path.insert(0, os.path.join(ROOT, "myrtle-vision_b200"))
    from myrtle_vision.utils.parallel import GradReducer
    dist.init_process_group("gloo", init_method="tcp://127.0.0.1:%d" % port, world_size=world,
                            rank=rank)
    try:
        red = GradReducer()
        assert red.world == world
        # a flat bucket whose per-rank scale differs (each rank picks its own power of two)
        scale = 2.0 ** (rank + 3)
        true = torch.arange(10, dtype=torch.float32) * (rank + 1)
        flat = true * scale
        red.reduce_slice(flat, 1.0 / scale)
        red.wait()
        want = sum(torch.arange(10, dtype=torch.float32) * (r + 1) for r in range(world)) / world
        assert torch.allclose(flat, want)
        # parameters: one never gets a gradient and must be skipped without deadlock
        a = torch.nn.Parameter(torch.ones(3)); b = torch.nn.Parameter(torch.ones(2, 2))
        unused = torch.nn.Parameter(torch.ones(5))
        a.grad = torch.full((3,), float(rank)); b.grad = torch.full((2, 2), float(10 * rank))
        n = red.reduce_params([a, unused, b])
        assert n == 2 and unused.grad is None
        assert torch.allclose(a.grad, torch.full((3,), 0.5))
        assert torch.allclose(b.grad, torch.full((2, 2), 5.0))
        ret[rank] = True
    finally:
        dist.destroy_process_group()


def test_reducer_world_size_2():
    import random
    port = 29500 + random.randint(0, 2000)
    mgr = mp.Manager()
    ret = mgr.dict()
    mp.spawn(_worker, args=(2, port, ret), nprocs=2, join=True)
    assert ret.get(0) and ret.get(1)
