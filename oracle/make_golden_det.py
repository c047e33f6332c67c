"""Generate tests/golden/det_criterion.npz by running the UNMODIFIED reference criterion here.

    python oracle/make_golden_det.py

Imports SetCriterion / HungarianMatcher from /root/reference/src/myrtle_vision/models/
{detector,matcher}.py (torchvision + SciPy, both in this container), feeds them seeded synthetic
DIOR-shaped predictions/targets (100 queries, 20 classes, 1..20 boxes per
image, cxcywh in (0.1,0.9)x(0.05,0.3)) and stores inputs, the matching and every loss term with
its gradient w.r.t. the predictions.  Test infrastructure only; the GPU box never reads
/root/reference.
"""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, "/root/reference/src")

from myrtle_vision.models.detector import SetCriterion  # noqa: E402  (the reference)
from myrtle_vision.models.matcher import HungarianMatcher  # noqa: E402

WEIGHTS = {"loss_ce": 1, "loss_bbox": 5, "loss_giou": 2}     # detection/train_configs/yolos_tiny.json:25-30


def make_case(seed, B=4, Q=100, C=20):
    g = torch.Generator().manual_seed(seed)
    logits = torch.randn(B, Q, C + 1, generator=g)
    boxes = torch.rand(B, Q, 4, generator=g) * 0.8 + 0.1
    boxes[..., 2:] = boxes[..., 2:] * 0.3
    targets = []
    for b in range(B):
        k = int(torch.randint(1, 21, (1,), generator=g))
        if seed % 2 == 1 and b == 1:
            k = 0                                           # an image without objects
        cxcy = torch.rand(k, 2, generator=g) * 0.8 + 0.1
        wh = torch.rand(k, 2, generator=g) * 0.25 + 0.05
        targets.append({"labels": torch.randint(0, C, (k,), generator=g),
                        "boxes": torch.cat([cxcy, wh], dim=1)})
    return logits, boxes, targets


def main():
    out = {}
    # seeds 1, 2: HungarianMatcher() with its default costs 1 / 1 / 1, exactly as detection/train.py:199 builds it
    # (the weight_dict only weighs the loss terms); seed 3: explicit non-default costs
    for seed, costs in ((1, None), (2, None), (3, (1, 5, 2))):
        logits, boxes, targets = make_case(seed)
        logits.requires_grad_(True)
        boxes.requires_grad_(True)
        matcher = HungarianMatcher() if costs is None else HungarianMatcher(*costs)
        crit = SetCriterion(20, matcher, WEIGHTS, 0.1, ["labels", "boxes", "cardinality"])
        losses = crit({"pred_logits": logits, "pred_boxes": boxes}, targets)
        total = sum(losses[k] * WEIGHTS[k] for k in WEIGHTS)
        total.backward()
        indices = crit.matcher({"pred_logits": logits.detach(), "pred_boxes": boxes.detach()}, targets)
        p = "s%d_" % seed
        out[p + "costs"] = np.array([matcher.cost_class, matcher.cost_bbox, matcher.cost_giou], dtype=np.float64)
        out[p + "logits"] = logits.detach().numpy()
        out[p + "boxes"] = boxes.detach().numpy()
        out[p + "n_tgt"] = np.array([len(t["labels"]) for t in targets])
        out[p + "tgt_labels"] = torch.cat([t["labels"] for t in targets]).numpy()
        out[p + "tgt_boxes"] = torch.cat([t["boxes"] for t in targets]).numpy()
        out[p + "match_src"] = torch.cat([i for i, _ in indices]).numpy()
        out[p + "match_tgt"] = torch.cat([j for _, j in indices]).numpy()
        for k, v in losses.items():
            out[p + k] = np.array(float(v))
        out[p + "total"] = np.array(float(total))
        out[p + "grad_logits"] = logits.grad.numpy()
        out[p + "grad_boxes"] = boxes.grad.numpy()
    path = os.path.join(ROOT, "tests", "golden", "det_criterion.npz")
    np.savez_compressed(path, **out)
    print("wrote", path, {k: v.shape for k, v in out.items() if v.ndim == 0 or k.endswith("n_tgt")})


if __name__ == "__main__":
    main()
