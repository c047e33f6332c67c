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


def _det_worker(rank, world, port, ret):
    """Detection criterion under data parallelism: box losses are normalised by the all-reduced mean number of
    boxes (reference models/detector.py:134-138), identically on the list path and the padded path."""
    sys.path.insert(0, os.path.join(ROOT, "myrtle-vision_b200"))
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    from myrtle_vision.models.detector import SetCriterion
    from myrtle_vision.models.matcher import HungarianMatcher, pad_targets
    from test_detection_criterion import scipy_match_padded
    dist.init_process_group("gloo", init_method="tcp://127.0.0.1:%d" % port, world_size=world, rank=rank)
    try:
        g = torch.Generator().manual_seed(100 + rank)
        n_boxes = [3, 1] if rank == 0 else [6, 2]               # 4 boxes on rank 0, 8 on rank 1: mean 6
        out = {"pred_logits": torch.randn(2, 12, 5, generator=g),
               "pred_boxes": torch.rand(2, 12, 4, generator=g) * 0.4 + 0.2}
        targets = [{"labels": torch.randint(0, 4, (n,), generator=g),
                    "boxes": torch.rand(n, 4, generator=g) * 0.4 + 0.2} for n in n_boxes]
        matcher = HungarianMatcher(1, 5, 2)
        matcher.match_padded = scipy_match_padded(matcher)      # CPU stand-in for the device assignment
        crit = SetCriterion(4, matcher, {"loss_ce": 1, "loss_bbox": 5, "loss_giou": 2}, 0.1,
                            ["labels", "boxes", "cardinality"])
        listed = crit(out, targets)
        padded = crit(out, pad_targets(targets, capacity=8))
        for k in listed:
            assert abs(float(listed[k]) - float(padded[k])) < 1e-5, k
        # same matching, normalised by this rank's own count: the ratio is local boxes / mean boxes
        local = float(sum(n_boxes))
        idx = matcher(out, targets)
        src = torch.cat([out["pred_boxes"][b, i] for b, (i, _) in enumerate(idx)])
        tgt = torch.cat([t["boxes"][j] for t, (_, j) in zip(targets, idx)])
        l1_local = float((src - tgt).abs().sum() / local)
        assert abs(float(listed["loss_bbox"]) - l1_local * local / 6.0) < 1e-5
        ret[rank] = True
    finally:
        dist.destroy_process_group()


def test_detection_criterion_world_size_2():
    import random
    port = 31500 + random.randint(0, 2000)
    mgr = mp.Manager()
    ret = mgr.dict()
    mp.spawn(_det_worker, args=(2, port, ret), nprocs=2, join=True)
    assert ret.get(0) and ret.get(1)
