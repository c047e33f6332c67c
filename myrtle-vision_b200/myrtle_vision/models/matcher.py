"""Bipartite matching between predicted and ground-truth boxes (config 5 loss, host side).

Drop-in for src/myrtle_vision/models/matcher.py:13-87 of the reference (`HungarianMatcher`, same
constructor, same `forward(outputs, targets)` result: one `(prediction_idx, target_idx)` int64
pair per image, CPU tensors).  The matching cost is the reference's
`cost_bbox * L1 + cost_class * (-prob[target class]) + cost_giou * (-GIoU)`, but it is built as
ONE padded `[B, Q, Tmax]` block on the device (every image only against its own targets) and
leaves the GPU in a single copy; the reference builds the `[B*Q, sum(T)]` cross product of
every prediction with every image's targets and throws the off-diagonal blocks away.
For CPU tensors the assignment is SciPy's `linear_sum_assignment`, as in the reference (:83-86).
For CUDA tensors it is `mv_linear_sum_assignment` (csrc/assign.cu, one warp per image): `match_padded`
returns the matching as a device tensor without any host synchronisation, so the detection train step
(SetCriterion's padded path) can be captured in one CUDA graph; `forward` converts the same result into
the reference's list of CPU index pairs.
"""
import torch
from scipy.optimize import linear_sum_assignment
from torch import nn


def cxcywh_to_xyxy(b):
    cx, cy, w, h = b.unbind(-1)
    return torch.stack((cx - 0.5 * w, cy - 0.5 * h, cx + 0.5 * w, cy + 0.5 * h), dim=-1)


def generalized_iou(a, b):
    """GIoU of xyxy boxes, broadcasting over leading dimensions (a[..., 4] vs b[..., 4])."""
    area_a = (a[..., 2] - a[..., 0]) * (a[..., 3] - a[..., 1])
    area_b = (b[..., 2] - b[..., 0]) * (b[..., 3] - b[..., 1])
    iw = (torch.minimum(a[..., 2], b[..., 2]) - torch.maximum(a[..., 0], b[..., 0])).clamp(min=0)
    ih = (torch.minimum(a[..., 3], b[..., 3]) - torch.maximum(a[..., 1], b[..., 1])).clamp(min=0)
    inter = iw * ih
    union = area_a + area_b - inter
    iou = inter / union
    cw = (torch.maximum(a[..., 2], b[..., 2]) - torch.minimum(a[..., 0], b[..., 0])).clamp(min=0)
    ch = (torch.maximum(a[..., 3], b[..., 3]) - torch.minimum(a[..., 1], b[..., 1])).clamp(min=0)
    hull = cw * ch
    return iou - (hull - union) / hull


def pad_targets(targets, capacity=None, multiple=8):
    """list of {'labels': [n], 'boxes': [n, 4]} -> {'labels': int64 [B, T], 'boxes': fp32 [B, T, 4],
    'sizes': int32 [B]} on the targets' device, T = max(n) rounded up to `multiple` (or `capacity`).
    Padding boxes are the valid box (0.5, 0.5, 1, 1); sizes come from tensor shapes (no device sync)."""
    sizes = [int(t["boxes"].shape[0]) for t in targets]
    tmax = max(sizes) if sizes else 0
    T = max(multiple, -(-tmax // multiple) * multiple)
    if capacity is not None:
        assert capacity >= tmax, f"{tmax} targets in one image exceed the capacity {capacity}"
        T = capacity
    dev = targets[0]["boxes"].device if targets else torch.device("cpu")
    labels = torch.zeros(len(targets), T, dtype=torch.int64, device=dev)
    boxes = torch.empty(len(targets), T, 4, dtype=torch.float32, device=dev)
    boxes[..., :2] = 0.5
    boxes[..., 2:] = 1.0
    for b, t in enumerate(targets):
        if sizes[b]:
            labels[b, :sizes[b]] = t["labels"]
            boxes[b, :sizes[b]] = t["boxes"]
    return {"labels": labels, "boxes": boxes, "sizes": torch.tensor(sizes, dtype=torch.int32).to(dev)}


class HungarianMatcher(nn.Module):
    def __init__(self, cost_class: float = 1, cost_bbox: float = 1, cost_giou: float = 1):
        super().__init__()
        self.cost_class = cost_class
        self.cost_bbox = cost_bbox
        self.cost_giou = cost_giou
        assert cost_class != 0 or cost_bbox != 0 or cost_giou != 0, "all costs cant be 0"

    @torch.no_grad()
    def cost_blocks(self, outputs, targets):
        """[B, Q, Tmax] cost of matching prediction q of image b to its own target t (padding
        columns hold copies of valid-shaped dummy boxes and are cut off by the caller)."""
        logits, boxes = outputs["pred_logits"], outputs["pred_boxes"]
        B, Q = logits.shape[:2]
        sizes = [int(t["boxes"].shape[0]) for t in targets]
        tmax = max(sizes) if sizes else 0
        if tmax == 0:
            return torch.zeros(B, Q, 0), sizes
        dev = logits.device
        ids = torch.zeros(B, tmax, dtype=torch.int64, device=dev)
        tb = torch.empty(B, tmax, 4, dtype=boxes.dtype, device=dev)
        tb[..., :2] = 0.5
        tb[..., 2:] = 1.0
        for b, t in enumerate(targets):
            if sizes[b]:
                ids[b, :sizes[b]] = t["labels"].to(dev)
                tb[b, :sizes[b]] = t["boxes"].to(dev)
        return self._cost(logits, boxes, ids, tb).cpu(), sizes

    def _cost(self, logits, boxes, ids, tb):
        B, Q = logits.shape[:2]
        prob = logits.softmax(-1)                                         # [B, Q, C+1]
        c_class = -prob.gather(2, ids[:, None, :].expand(B, Q, ids.shape[1]))
        c_bbox = (boxes[:, :, None, :] - tb[:, None, :, :]).abs().sum(-1)
        c_giou = -generalized_iou(cxcywh_to_xyxy(boxes)[:, :, None, :],
                                  cxcywh_to_xyxy(tb)[:, None, :, :])
        return self.cost_bbox * c_bbox + self.cost_class * c_class + self.cost_giou * c_giou

    @torch.no_grad()
    def match_padded(self, outputs, padded):
        """padded: pad_targets(...) on the outputs' CUDA device -> int32 [B, T], the prediction matched to
        every target (-1 in padding columns).  Stream-ordered, no host synchronisation."""
        import mv_native
        logits, boxes = outputs["pred_logits"].float(), outputs["pred_boxes"].float()
        cost = self._cost(logits, boxes, padded["labels"], padded["boxes"])
        return mv_native.linear_sum_assignment(cost, padded["sizes"])

    @torch.no_grad()
    def forward(self, outputs, targets):
        if outputs["pred_logits"].is_cuda:
            dev = outputs["pred_logits"].device
            padded = pad_targets([{k: t[k].to(dev) for k in ("labels", "boxes")} for t in targets])
            match = self.match_padded(outputs, padded).cpu()
            result = []
            for b, t in enumerate(targets):
                m = match[b, :int(t["boxes"].shape[0])].to(torch.int64)
                j = torch.nonzero(m >= 0).flatten()
                order = torch.argsort(m[j])                # SciPy returns pairs sorted by prediction
                result.append((m[j][order], j[order]))
            return result
        C, sizes = self.cost_blocks(outputs, targets)
        result = []
        for b, n in enumerate(sizes):
            i, j = linear_sum_assignment(C[b, :, :n].numpy())
            result.append((torch.as_tensor(i, dtype=torch.int64), torch.as_tensor(j, dtype=torch.int64)))
        return result
