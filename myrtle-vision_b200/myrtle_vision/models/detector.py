"""Detection loss and post-processing (config 5; reference src/myrtle_vision/models/detector.py).

`SetCriterion(num_classes, matcher, weight_dict, eos_coef, losses)` keeps the reference's
constructor, `forward(outputs, targets) -> dict` and loss names (`loss_ce`, `class_error`,
`cardinality_error`, `loss_bbox`, `loss_giou`; reference :42-96, :121-145).  Differences are
internal: matched pairs are gathered once and shared by all losses, and the GIoU of matched pairs
is computed pairwise instead of as the diagonal of an all-pairs matrix (:89-94).
`num_boxes` is averaged over ranks exactly as the reference does (:134-138).

Padded path: when `targets` is the dict made by `matcher.pad_targets` (fixed
[B, T] capacity, on the CUDA device), matching runs on the device (`HungarianMatcher.match_padded`)
and every loss is a masked reduction over the [B, T] pairs, so `forward` contains no host
synchronisation and no data-dependent shape: the whole detection step is capturable in one CUDA
graph.  The values equal the list path's (tests/test_gpu_detection.py).
"""
import torch
import torch.nn.functional as F
from torch import nn

from myrtle_vision.models.matcher import cxcywh_to_xyxy, generalized_iou
from myrtle_vision.utils.utils import accuracy, get_world_size, is_dist_avail_and_initialized


class SetCriterion(nn.Module):
    def __init__(self, num_classes, matcher, weight_dict, eos_coef, losses):
        super().__init__()
        self.num_classes = num_classes
        self.matcher = matcher
        self.weight_dict = weight_dict
        self.eos_coef = eos_coef
        self.losses = losses
        empty_weight = torch.ones(self.num_classes + 1)
        empty_weight[-1] = self.eos_coef
        self.register_buffer("empty_weight", empty_weight)

    # ---- index helpers (same results as the reference's _get_src/_get_tgt_permutation_idx)
    def _get_src_permutation_idx(self, indices):
        batch_idx = torch.cat([torch.full_like(src, i) for i, (src, _) in enumerate(indices)])
        src_idx = torch.cat([src for (src, _) in indices])
        return batch_idx, src_idx

    def _get_tgt_permutation_idx(self, indices):
        batch_idx = torch.cat([torch.full_like(tgt, i) for i, (_, tgt) in enumerate(indices)])
        tgt_idx = torch.cat([tgt for (_, tgt) in indices])
        return batch_idx, tgt_idx

    def loss_labels(self, outputs, targets, indices, num_boxes, log=True):
        assert "pred_logits" in outputs
        logits = outputs["pred_logits"]
        idx = self._get_src_permutation_idx(indices)
        matched = torch.cat([t["labels"][j] for t, (_, j) in zip(targets, indices)]).to(logits.device)
        classes = torch.full(logits.shape[:2], self.num_classes, dtype=torch.int64, device=logits.device)
        classes[idx] = matched
        losses = {"loss_ce": F.cross_entropy(logits.transpose(1, 2), classes, self.empty_weight)}
        if log:
            losses["class_error"] = 100 - accuracy(logits[idx], matched)[0]
        return losses

    @torch.no_grad()
    def loss_cardinality(self, outputs, targets, indices, num_boxes):
        logits = outputs["pred_logits"]
        lengths = torch.as_tensor([len(t["labels"]) for t in targets], device=logits.device)
        predicted = (logits.argmax(-1) != logits.shape[-1] - 1).sum(1)
        return {"cardinality_error": F.l1_loss(predicted.float(), lengths.float())}

    def loss_boxes(self, outputs, targets, indices, num_boxes):
        assert "pred_boxes" in outputs
        idx = self._get_src_permutation_idx(indices)
        src = outputs["pred_boxes"][idx]
        tgt = torch.cat([t["boxes"][j] for t, (_, j) in zip(targets, indices)], dim=0).to(src.device)
        giou = generalized_iou(cxcywh_to_xyxy(src), cxcywh_to_xyxy(tgt))
        return {"loss_bbox": (src - tgt).abs().sum() / num_boxes,
                "loss_giou": (1 - giou).sum() / num_boxes}

    def get_loss(self, loss, outputs, targets, indices, num_boxes, **kwargs):
        loss_map = {"labels": self.loss_labels, "cardinality": self.loss_cardinality,
                    "boxes": self.loss_boxes}
        assert loss in loss_map, f"do you really want to compute {loss} loss?"
        return loss_map[loss](outputs, targets, indices, num_boxes, **kwargs)

    def forward_padded(self, outputs, padded):
        logits, boxes = outputs["pred_logits"], outputs["pred_boxes"]
        B, Q, C1 = logits.shape
        labels, tb, sizes = padded["labels"], padded["boxes"], padded["sizes"]
        T = labels.shape[1]
        match = self.matcher.match_padded(outputs, padded)                          # int32 [B, T]
        m = (torch.arange(T, device=logits.device)[None, :] < sizes[:, None]) & (match >= 0)
        q = match.clamp(min=0).to(torch.int64)
        num_boxes = sizes.sum().to(torch.float32)
        if is_dist_avail_and_initialized():
            torch.distributed.all_reduce(num_boxes)
        num_boxes = torch.clamp(num_boxes / get_world_size(), min=1)
        losses = {}
        if "labels" in self.losses:
            classes = torch.full((B, Q + 1), self.num_classes, dtype=torch.int64, device=logits.device)
            classes.scatter_(1, torch.where(m, q, Q), torch.where(m, labels, self.num_classes))
            losses["loss_ce"] = F.cross_entropy(logits.transpose(1, 2), classes[:, :Q], self.empty_weight)
            with torch.no_grad():
                top = logits.gather(1, q[..., None].expand(B, T, C1)).argmax(-1)
                pairs = m.sum()
                hit = ((top == labels) & m).sum().float() * 100.0 / pairs.clamp(min=1)
                losses["class_error"] = 100 - hit
        if "cardinality" in self.losses:
            with torch.no_grad():
                predicted = (logits.argmax(-1) != C1 - 1).sum(1)
                losses["cardinality_error"] = F.l1_loss(predicted.float(), sizes.float())
        if "boxes" in self.losses:
            src = boxes.gather(1, q[..., None].expand(B, T, 4))
            zero = torch.zeros((), dtype=src.dtype, device=src.device)
            giou = generalized_iou(cxcywh_to_xyxy(src), cxcywh_to_xyxy(tb))
            losses["loss_bbox"] = torch.where(m, (src - tb).abs().sum(-1), zero).sum() / num_boxes
            losses["loss_giou"] = torch.where(m, 1 - giou, zero).sum() / num_boxes
        unknown = [k for k in self.losses if k not in ("labels", "cardinality", "boxes")]
        assert not unknown, f"do you really want to compute {unknown[0]} loss?"
        return losses

    def forward(self, outputs, targets):
        outputs = {k: v for k, v in outputs.items() if k != "aux_outputs"}
        if isinstance(targets, dict):
            return self.forward_padded(outputs, targets)
        indices = self.matcher(outputs, targets)
        device = next(iter(outputs.values())).device
        num_boxes = torch.as_tensor([sum(len(t["labels"]) for t in targets)], dtype=torch.float,
                                    device=device)
        if is_dist_avail_and_initialized():
            torch.distributed.all_reduce(num_boxes)
        num_boxes = torch.clamp(num_boxes / get_world_size(), min=1).item()
        losses = {}
        for loss in self.losses:
            losses.update(self.get_loss(loss, outputs, targets, indices, num_boxes))
        return losses


class PostProcess(nn.Module):
    """Model output -> per-image {'scores','labels','boxes'} in absolute xyxy pixels (reference :148-176)."""

    @torch.no_grad()
    def forward(self, outputs, target_sizes):
        logits, boxes = outputs["pred_logits"], outputs["pred_boxes"]
        assert len(logits) == len(target_sizes)
        assert target_sizes.shape[1] == 2
        scores, labels = F.softmax(logits, -1)[..., :-1].max(-1)
        h, w = target_sizes.unbind(1)
        scale = torch.stack([w, h, w, h], dim=1)[:, None, :]
        xyxy = cxcywh_to_xyxy(boxes) * scale
        return [{"scores": s, "labels": l, "boxes": b} for s, l, b in zip(scores, labels, xyxy)]
