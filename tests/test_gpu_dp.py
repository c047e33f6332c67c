"""Data-parallel parity on real GPUs (needs >= 2 devices, NCCL): gradients of a 2-rank step equal
the 1-GPU gradients on the concatenated batch; the step runs twice, which
is exactly where stock DDP fails on the reference's unused det tokens."""
import os
import sys

import pytest
import torch

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _worker(rank, world, port, ret):
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "myrtle-vision_b200"))
    import torch.distributed as dist
    import torch.nn.functional as F
    from myrtle_vision.models.vit import ViT
    from myrtle_vision.utils.parallel import DataParallel
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", init_method="tcp://127.0.0.1:%d" % port, world_size=world, rank=rank)
    try:
        dev = torch.device("cuda", rank)
        torch.manual_seed(11)
        kw = dict(decoder="classification", image_size=96, patch_size=16, num_classes=7, dim=128, depth=2,
                  heads=2, mlp_dim=256, q_format="FP16_32")
        model = ViT(**kw).to(dev).train()
        g = torch.Generator().manual_seed(5)
        img = torch.randn(8, 3, 96, 96, generator=g).clamp(-1, 1)
        y = torch.randint(0, 7, (8,), generator=g)
        # single-GPU reference on the full batch (mean loss over 8)
        F.cross_entropy(model(img.to(dev)), y.to(dev)).backward()
        ref = {n: p.grad.clone() for n, p in model.named_parameters() if p.grad is not None}
        model.zero_grad()
        dp = DataParallel(model)
        shard = slice(rank * 4, rank * 4 + 4)
        for it in range(2):                                   # second iteration must also work
            model.zero_grad()
            F.cross_entropy(dp(img[shard].to(dev)), y[shard].to(dev)).backward()
        torch.cuda.synchronize()
        worst = 0.0
        for n, p in model.named_parameters():
            if n in ref:
                worst = max(worst, ((p.grad - ref[n]).norm() / (ref[n].norm() + 1e-20)).item())
            else:
                assert p.grad is None, n
        ret[rank] = worst
        # overflow flag under data parallelism: a saturation that only rank 1 sees reaches every rank through the slot
        # in front of the last gradient bucket (no extra collective), so all replicas skip the same step
        eng = model.engine()
        assert float(eng.found_inf) == 0.0
        if rank == 1:
            eng.scaler_state[1] = 2.0 ** 40               # rank 1 only: its fp16 gradient operands saturate
        model.zero_grad()
        F.cross_entropy(dp(img[shard].to(dev)), y[shard].to(dev)).backward()
        torch.cuda.synchronize()
        ret["found_inf_%d" % rank] = float(eng.found_inf)
    finally:
        dist.destroy_process_group()


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs 2 GPUs")
def test_two_rank_gradients_match_single_gpu():
    import random
    import torch.multiprocessing as mp
    port = 29500 + random.randint(0, 2000)
    ret = mp.Manager().dict()
    mp.spawn(_worker, args=(2, port, ret), nprocs=2, join=True)
    # fp16 tensor-core operands with per-rank power-of-two scales: 5e-3 relative L2
    assert ret[0] < 5e-3 and ret[1] < 5e-3, dict(ret)
    assert ret["found_inf_0"] == 1.0 and ret["found_inf_1"] == 1.0, dict(ret)
