"""Device-side detection matching: mv_linear_sum_assignment against SciPy (the
reference's matcher, models/matcher.py:83-86), the CUDA HungarianMatcher / padded SetCriterion against
the fixtures made by the unmodified reference, and the detection step captured as ONE CUDA graph."""
import numpy as np
import pytest
import torch
from scipy.optimize import linear_sum_assignment

from test_detection_criterion import WEIGHTS, load_case

pytestmark = pytest.mark.gpu


def scipy_match(cost, sizes):
    B, Q, T = cost.shape
    want = np.full((B, T), -1, dtype=np.int32)
    for b in range(B):
        ri, ci = linear_sum_assignment(cost[b, :, :sizes[b]].astype(np.float64))
        want[b, ci] = ri
    return want


@pytest.mark.parametrize("B,Q,T", [(8, 100, 24), (3, 100, 128), (64, 100, 16), (2, 1, 1), (5, 37, 40), (2, 1000, 1024)])
def test_assignment_kernel_equals_scipy(B, Q, T):
    import mv_native
    rng = np.random.default_rng(B * 7 + T)
    cost = (rng.standard_normal((B, Q, T)) * 2).astype(np.float32)
    sizes = rng.integers(0, T + 1, size=B).astype(np.int32)
    sizes[0], sizes[-1] = T, 0 if B > 1 else T
    flag = torch.zeros(1, dtype=torch.int32, device="cuda")
    match = mv_native.linear_sum_assignment(torch.from_numpy(cost).cuda(), torch.from_numpy(sizes).cuda(), flag=flag)
    assert int(flag) == 0
    assert np.array_equal(match.cpu().numpy(), scipy_match(cost, sizes))


def test_assignment_kernel_flags_infeasible_blocks_and_empty_batches():
    import mv_native
    cost = torch.full((2, 4, 8), float("inf"), device="cuda")
    flag = torch.zeros(1, dtype=torch.int32, device="cuda")
    mv_native.linear_sum_assignment(cost, torch.tensor([3, 0], dtype=torch.int32, device="cuda"), flag=flag)
    assert int(flag) == 1
    out = mv_native.linear_sum_assignment(torch.zeros(0, 4, 8, device="cuda"), torch.zeros(0, dtype=torch.int32, device="cuda"))
    assert out.shape == (0, 8)
    with pytest.raises(mv_native.MvError):
        mv_native.linear_sum_assignment(torch.zeros(1, 4, 8), torch.zeros(1, dtype=torch.int32))


@pytest.mark.parametrize("seed", [1, 2, 3])
def test_cuda_matcher_and_padded_criterion_equal_the_reference(seed):
    from myrtle_vision.models.detector import SetCriterion
    from myrtle_vision.models.matcher import HungarianMatcher, pad_targets
    z, p, logits, boxes, targets = load_case(seed)
    dev = torch.device("cuda")
    logits, boxes = logits.to(dev), boxes.to(dev)
    targets = [{k: v.to(dev) for k, v in t.items()} for t in targets]
    matcher = HungarianMatcher(*z[p + "costs"].tolist())   # seeds 1, 2: the defaults, as detection/train.py:199
    indices = matcher({"pred_logits": logits, "pred_boxes": boxes}, targets)     # the reference's result format
    assert torch.cat([i for i, _ in indices]).tolist() == z[p + "match_src"].tolist()
    assert torch.cat([j for _, j in indices]).tolist() == z[p + "match_tgt"].tolist()
    crit = SetCriterion(20, matcher, WEIGHTS, 0.1, ["labels", "boxes", "cardinality"]).to(dev)
    for tgt in (targets, pad_targets(targets, capacity=32 if seed != 2 else None)):
        lg, bx = logits.clone().requires_grad_(True), boxes.clone().requires_grad_(True)
        losses = crit({"pred_logits": lg, "pred_boxes": bx}, tgt)
        for k in losses:
            assert abs(float(losses[k]) - float(z[p + k])) <= 1e-5 * max(1.0, abs(float(z[p + k]))), k
        sum(losses[k] * WEIGHTS[k] for k in WEIGHTS).backward()
        np.testing.assert_allclose(lg.grad.cpu().numpy(), z[p + "grad_logits"], rtol=1e-4, atol=1e-7)
        np.testing.assert_allclose(bx.grad.cpu().numpy(), z[p + "grad_boxes"], rtol=1e-4, atol=1e-7)


def test_detection_step_is_one_graph_and_equals_the_eager_step():
    """fwd + SetCriterion (device matching) + bwd captured once; replays on new batches give the eager
    step's loss and gradients."""
    from myrtle_vision.datasets.synthetic import SyntheticVision, detection_collate
    from myrtle_vision.models.matcher import pad_targets
    from myrtle_vision.models.vit import ViT
    from myrtle_vision.utils.graph import GraphedTrainStep
    from myrtle_vision.utils.trainer import build_criterion, to_device
    dev = torch.device("cuda")
    torch.manual_seed(0)
    model = ViT(decoder="detection", image_size=176, patch_size=16, num_classes=3, dim=128, depth=2, heads=2,
                mlp_dim=256, q_format="FP16_32").to(dev).train()
    ds = SyntheticVision("detection", 8, 176, 3, seed=3)
    cfg = {"loss_ce": 1.0, "loss_bbox": 5.0, "loss_giou": 2.0, "eos_coef": 0.1}
    criterion = build_criterion("detection", cfg, 3, dev)
    batches = []
    for part in (range(0, 4), range(4, 8)):
        img, ts = detection_collate([ds[i] for i in part])
        batches.append((img.to(dev), to_device(ts, dev), to_device(pad_targets(ts, capacity=32), dev)))
    step = GraphedTrainStep(model, criterion, batches[0][0], batches[0][2])
    for img, ts, padded in batches[::-1]:
        loss_g = float(step(img, padded))
        grads_g = {n: p.grad.clone() for n, p in model.named_parameters() if p.grad is not None}
        model.zero_grad(set_to_none=True)
        loss_e = criterion(model(img), ts)            # list targets: the reference-format path
        loss_e.backward()
        assert abs(loss_g - float(loss_e)) <= 1e-5 * abs(float(loss_e))
        for n, p in model.named_parameters():
            if p.grad is not None:
                # fp16 gradient operands: last-bit differences of the criterion's fp32 sums flip a few roundings
                assert (grads_g[n] - p.grad).norm() <= 2e-3 * p.grad.norm() + 1e-7, n
