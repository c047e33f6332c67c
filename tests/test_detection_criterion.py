"""Detection loss (SetCriterion + HungarianMatcher, SURVEY.md §8a last row) against fixtures made by
the unmodified reference criterion (oracle/make_golden_det.py -> tests/golden/det_criterion.npz)."""
import os

import numpy as np
import pytest
import torch

from oracle.golden_cases import GOLD

WEIGHTS = {"loss_ce": 1, "loss_bbox": 5, "loss_giou": 2}


def load_case(seed):
    z = np.load(os.path.join(GOLD, "det_criterion.npz"))
    p = "s%d_" % seed
    n = z[p + "n_tgt"].tolist()
    labels = torch.from_numpy(z[p + "tgt_labels"]).split(n)
    boxes = torch.from_numpy(z[p + "tgt_boxes"]).split(n)
    targets = [{"labels": l, "boxes": b} for l, b in zip(labels, boxes)]
    return z, p, torch.from_numpy(z[p + "logits"]), torch.from_numpy(z[p + "boxes"]), targets


@pytest.mark.parametrize("seed", [1, 2, 3])
def test_matcher_and_losses_equal_the_reference(seed):
    from myrtle_vision.models.detector import SetCriterion
    from myrtle_vision.models.matcher import HungarianMatcher
    z, p, logits, boxes, targets = load_case(seed)
    matcher = HungarianMatcher(*z[p + "costs"].tolist())   # seeds 1, 2: the defaults, as detection/train.py:199
    indices = matcher({"pred_logits": logits, "pred_boxes": boxes}, targets)
    assert [len(i) for i, _ in indices] == [min(100, len(t["labels"])) for t in targets]
    assert torch.cat([i for i, _ in indices]).numpy().tolist() == z[p + "match_src"].tolist()
    assert torch.cat([j for _, j in indices]).numpy().tolist() == z[p + "match_tgt"].tolist()

    logits.requires_grad_(True)
    boxes.requires_grad_(True)
    crit = SetCriterion(20, matcher, WEIGHTS, 0.1, ["labels", "boxes", "cardinality"])
    losses = crit({"pred_logits": logits, "pred_boxes": boxes}, targets)
    assert set(losses) == {"loss_ce", "class_error", "loss_bbox", "loss_giou", "cardinality_error"}
    for k in losses:
        assert abs(float(losses[k]) - float(z[p + k])) <= 1e-5 * max(1.0, abs(float(z[p + k]))), k
    total = sum(losses[k] * WEIGHTS[k] for k in WEIGHTS)
    total.backward()
    assert abs(float(total) - float(z[p + "total"])) < 1e-5 * abs(float(z[p + "total"]))
    np.testing.assert_allclose(logits.grad.numpy(), z[p + "grad_logits"], rtol=1e-4, atol=1e-7)
    np.testing.assert_allclose(boxes.grad.numpy(), z[p + "grad_boxes"], rtol=1e-4, atol=1e-7)


def test_criterion_edge_cases_and_postprocess():
    from myrtle_vision.models.detector import PostProcess, SetCriterion
    from myrtle_vision.models.matcher import HungarianMatcher
    with pytest.raises(AssertionError):
        HungarianMatcher(0, 0, 0)
    g = torch.Generator().manual_seed(0)
    out = {"pred_logits": torch.randn(2, 7, 4, generator=g), "pred_boxes": torch.rand(2, 7, 4, generator=g) * 0.5 + 0.2}
    empty = [{"labels": torch.zeros(0, dtype=torch.int64), "boxes": torch.zeros(0, 4)} for _ in range(2)]
    crit = SetCriterion(3, HungarianMatcher(), WEIGHTS, 0.1, ["labels", "boxes", "cardinality"])
    losses = crit(out, empty)                      # no objects anywhere: num_boxes clamps to 1
    assert float(losses["loss_bbox"]) == 0.0 and float(losses["loss_giou"]) == 0.0
    assert torch.isfinite(losses["loss_ce"])
    with pytest.raises(AssertionError):
        crit.get_loss("masks", out, empty, [], 1)
    res = PostProcess()(out, torch.tensor([[100, 200], [50, 60]]))
    assert len(res) == 2 and res[0]["boxes"].shape == (7, 4)
    cx, w = out["pred_boxes"][0, 0, 0], out["pred_boxes"][0, 0, 2]
    assert abs(float(res[0]["boxes"][0, 0]) - float((cx - w / 2) * 200)) < 1e-4
    assert int(res[0]["labels"].max()) <= 2       # the no-object class is never a label


def scipy_match_padded(matcher):
    """Stand-in for the device assignment (csrc/assign.cu) so that the padded criterion's masked
    reductions can be checked on the CPU; the kernel itself is held to SciPy in test_assign_host.py
    (same source, host build) and tests/test_gpu_detection.py."""
    from scipy.optimize import linear_sum_assignment

    def match_padded(outputs, padded):
        cost = matcher._cost(outputs["pred_logits"].detach(), outputs["pred_boxes"].detach(),
                             padded["labels"], padded["boxes"])
        match = torch.full(padded["labels"].shape, -1, dtype=torch.int32)
        for b, n in enumerate(padded["sizes"].tolist()):
            ri, ci = linear_sum_assignment(cost[b, :, :n].numpy())
            match[b, torch.as_tensor(ci, dtype=torch.int64)] = torch.as_tensor(ri, dtype=torch.int32)
        return match
    return match_padded


@pytest.mark.parametrize("seed,capacity", [(1, None), (2, 64), (3, 32)])
def test_padded_criterion_equals_the_reference(seed, capacity, monkeypatch):
    from myrtle_vision.models.detector import SetCriterion
    from myrtle_vision.models.matcher import HungarianMatcher, pad_targets
    z, p, logits, boxes, targets = load_case(seed)
    matcher = HungarianMatcher(*z[p + "costs"].tolist())   # seeds 1, 2: the defaults, as detection/train.py:199
    monkeypatch.setattr(matcher, "match_padded", scipy_match_padded(matcher))
    padded = pad_targets(targets, capacity=capacity)
    assert padded["labels"].shape[1] % 8 == 0 and padded["sizes"].tolist() == [len(t["labels"]) for t in targets]
    logits.requires_grad_(True)
    boxes.requires_grad_(True)
    crit = SetCriterion(20, matcher, WEIGHTS, 0.1, ["labels", "boxes", "cardinality"])
    losses = crit({"pred_logits": logits, "pred_boxes": boxes}, padded)
    assert set(losses) == {"loss_ce", "class_error", "loss_bbox", "loss_giou", "cardinality_error"}
    for k in losses:
        assert abs(float(losses[k]) - float(z[p + k])) <= 1e-5 * max(1.0, abs(float(z[p + k]))), k
    total = sum(losses[k] * WEIGHTS[k] for k in WEIGHTS)
    total.backward()
    np.testing.assert_allclose(logits.grad.numpy(), z[p + "grad_logits"], rtol=1e-4, atol=1e-7)
    np.testing.assert_allclose(boxes.grad.numpy(), z[p + "grad_boxes"], rtol=1e-4, atol=1e-7)


def test_padded_criterion_without_objects(monkeypatch):
    from myrtle_vision.models.detector import SetCriterion
    from myrtle_vision.models.matcher import HungarianMatcher, pad_targets
    g = torch.Generator().manual_seed(0)
    out = {"pred_logits": torch.randn(2, 7, 4, generator=g), "pred_boxes": torch.rand(2, 7, 4, generator=g) * 0.5 + 0.2}
    empty = [{"labels": torch.zeros(0, dtype=torch.int64), "boxes": torch.zeros(0, 4)} for _ in range(2)]
    matcher = HungarianMatcher()
    monkeypatch.setattr(matcher, "match_padded", scipy_match_padded(matcher))
    crit = SetCriterion(3, matcher, WEIGHTS, 0.1, ["labels", "boxes", "cardinality"])
    want = crit(out, empty)
    got = crit(out, pad_targets(empty))
    for k in want:
        assert abs(float(got[k]) - float(want[k])) < 1e-6, k
    with pytest.raises(AssertionError):
        pad_targets([{"labels": torch.zeros(9, dtype=torch.int64), "boxes": torch.rand(9, 4)}], capacity=8)
