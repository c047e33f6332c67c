"""Fused task losses of the B200 path.

`upsampled_cross_entropy(patch_logits, labels)` equals
`F.cross_entropy(nn.Upsample(size=labels.shape[-2:], mode="bilinear")(logits_bchw), labels)` — the
reference's segmentation head + criterion (models/vit.py:355-371, segmentation/train.py:188, 261) —
without materialising the full-resolution logits: one kernel produces the loss and its gradient
w.r.t. the patch logits.  `patch_logits` is the decoder Linear's output `[B, gh*gw, C]`, which
`ViT.forward` returns instead of the upsampled map while `vit.decoder.fused_loss` is set and the
model is in training mode.
"""
import torch

import mv_native as mv


class _UpsampledCrossEntropy(torch.autograd.Function):
    @staticmethod
    def forward(ctx, patch_logits, labels, ignore_index):
        acc, dy = mv.upsample_ce(patch_logits.detach().float(), labels, ignore_index)
        count = acc[1].clamp_min(1.0)
        ctx.save_for_backward(dy, count)
        return acc[0] / count

    @staticmethod
    def backward(ctx, grad):
        dy, count = ctx.saved_tensors
        return dy * (grad / count), None, None


def upsampled_cross_entropy(patch_logits, labels, ignore_index=-100):
    return _UpsampledCrossEntropy.apply(patch_logits, labels, ignore_index)
