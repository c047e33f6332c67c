"""Functional CPU restatement of the myrtle-vision quantised ViT forward.

TEST INFRASTRUCTURE ONLY (the model-level oracle and the CPU baseline "port").
Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl
reference legs may import it.  It is pinned against the reference itself:
oracle/make_golden.py imports the unmodified /root/reference/src/myrtle_vision
(through oracle/shim/qtorch) and tests/test_oracle_model.py compares this file
with the fixtures that script wrote under tests/golden/.  The fake-quant
arithmetic underneath (oracle/quant_oracle.c) is PARITY UNPINNED — see its
header.

What is restated (reference file:line, relative to /root/reference):
  * ViT.forward                        src/myrtle_vision/models/vit.py:267-320
  * Attention.forward                  src/myrtle_vision/models/vit.py:84-99
  * FeedForward / PreNorm / Residual   src/myrtle_vision/models/vit.py:17-56
  * the three decoders                 src/myrtle_vision/models/vit.py:325-396
  * quantiser placement per q_format   src/myrtle_vision/utils/quantize.py:253-327
    (verified site by site in SURVEY.md Appendix A)
  * straight-through backward          src/myrtle_vision/utils/quantize.py:77-89

Parameters are a flat dict keyed by the reference's FP32 state_dict names
(`canonical_key` maps the per-q_format names onto them).
All tensor math is stock PyTorch CPU fp32 — for GEMM / LayerNorm / softmax /
GELU / interpolate those ops *are* the reference's arithmetic.
"""
import re

import numpy as np
import torch
import torch.nn.functional as F

from . import quant_oracle as _qo

HALF = (5, 10)   # NumberFormat.HalfPrecisionFloat  -> FloatingPoint(exp=5, man=10)
TF32 = (8, 10)   # NumberFormat.TensorFloat32       -> FloatingPoint(exp=8, man=10)

# q_format -> (input/weight format, output format, FloatFunctional format, GELU-input format)
PLACEMENT = {
    "FP32":    (None, None, None, None),
    "FP16_32": (HALF, None, None, None),
    "TF32":    (TF32, None, None, None),
    "FP16_16": (HALF, HALF, HALF, HALF),
}


# After ModelQuantizer.convert() (utils/quantize.py:329-348): torch.quantization.convert drops every hook-based
# quantiser together with its observer hook (QuantStubs, Linear / LayerNorm output observers); only the
# FloatFunctional ones survive.  Pinned by tests/golden/convert_*.npz (oracle/make_golden_convert.py).
CONVERTED = {fmt: (None, None, pl[2], None) for fmt, pl in PLACEMENT.items()}


def convert_params(P, q_format):
    """QLinear.from_float / QLayerNorm.from_float (utils/quantize.py:134-166): Linear weights and LayerNorm gammas
    become their fake-quantised values; biases, tokens and positional embeddings are untouched."""
    fmt = PLACEMENT[str(q_format)][0]
    if fmt is None:
        return dict(P)
    with torch.no_grad():
        return {k: (fq(v, fmt) if k.endswith(".weight") else v) for k, v in P.items()}


def _fq_np(t, exp, man):
    a = t.detach().contiguous().float().numpy()
    if a.size > 4096:
        o = _qo.float_quantize_nearest_np(a, exp, man)
    else:
        o = _qo.float_quantize(a, exp, man, "nearest")
    return torch.from_numpy(o.reshape(a.shape)).to(t.dtype)


class _STEQuant(torch.autograd.Function):
    """forward: float_quantize nearest on x.data; backward: identity (utils/quantize.py:87-89) — or, with a gradient
    format (an option the reference never sets: QPyTorch's quantizer(forward_number, backward_number), whose backward
    is float_quantize(grad_output) with backward_rounding, qtorch/quant/quant_module.py), float_quantize nearest of
    the incoming gradient."""

    @staticmethod
    def forward(ctx, x, exp, man, gfmt):
        ctx.gfmt = gfmt
        return _fq_np(x, exp, man)

    @staticmethod
    def backward(ctx, g):
        if ctx.gfmt is not None:
            g = _fq_np(g, ctx.gfmt[0], ctx.gfmt[1])
        return g, None, None, None


# gradient format of the input / weight quantisers (the QuantStubs in front of Linear / LayerNorm and the weight
# fake-quant); None = the reference's straight-through backward.  Set by vit_forward(grad_format=...).
_GRAD_FMT = None


def fq(x, fmt, gfmt=None):
    if fmt is None:
        return x
    return _STEQuant.apply(x, fmt[0], fmt[1], gfmt)


_WRAPPED = re.compile(r"^(.*)\.1\.(weight|bias)$")


def canonical_key(key):
    """Per-q_format state_dict name -> FP32 name (`a.b.1.weight` -> `a.b.weight`)."""
    m = _WRAPPED.match(key)
    if m is None:
        return key
    return m.group(1) + "." + m.group(2)


def canonical_params(state_dict):
    return {canonical_key(k): v for k, v in state_dict.items()}


def _linear(x, P, name, fin, fout):
    """Sequential(QuantStub, qat.Linear): q(x) @ q(W).T + b, then the output observer."""
    y = F.linear(fq(x, fin, _GRAD_FMT), fq(P[name + ".weight"], fin, _GRAD_FMT), P[name + ".bias"])
    return fq(y, fout)


def _layernorm(x, P, name, fin, fout):
    xq = fq(x, fin, _GRAD_FMT)
    y = F.layer_norm(xq, (xq.shape[-1],), P[name + ".weight"], P[name + ".bias"], 1e-5)
    return fq(y, fout)


def vit_forward(P, img, *, decoder, patch_size=16, heads, q_format="FP32", num_det_tokens=100,
                image_size=None, dim_head=64, converted=False, grad_format=None):
    """P: canonical parameter dict.  Returns what ViT.forward returns.  converted=True: the quantiser set that is
    left after ModelQuantizer.convert() (pass convert_params(P, q_format) as P).  grad_format: (exp, man) applied by
    the input / weight quantisers to the gradient in backward (None: the reference's straight-through)."""
    global _GRAD_FMT
    _GRAD_FMT = grad_format
    fin, fout, ffn, fgelu = (CONVERTED if converted else PLACEMENT)[str(q_format)]
    b, c, h, w = img.shape
    p = patch_size
    gh, gw = h // p, w // p
    # patchify with (ph, pw, c) minor order                           vit.py:271-275
    x = img.reshape(b, c, gh, p, gw, p).permute(0, 2, 4, 3, 5, 1).reshape(b, gh * gw, p * p * c)
    x = _linear(x, P, "patch_to_embedding", fin, fout)              # vit.py:278
    dim = x.shape[-1]
    # cls token; the `decoder == "detection"` test is always False in the reference because
    # self.decoder has been overwritten by a Module (vit.py:196 vs :235-252): det tokens unused.
    cls = P["cls_token"].repeat(b, 1, 1)
    x = fq(torch.cat((cls, x), dim=1), ffn)                          # vit.py:290
    pos = P["pos_embedding"]
    pos_cls, pos_grid = pos[:, 0:1, :], pos[:, 1:, :]
    pos_grid = pos_grid.transpose(1, 2).view(1, -1, 14, 14)
    pos_grid = F.interpolate(pos_grid, size=(gh, gw), mode="bicubic", align_corners=False)
    pos_grid = pos_grid.view(1, -1, gh * gw).transpose(1, 2)
    pos_full = fq(torch.cat((pos_cls, pos_grid), dim=1), ffn)        # vit.py:302
    x = fq(x + pos_full.repeat(b, 1, 1), ffn)                        # vit.py:305-310

    depth = 0
    while "transformer.layers.%d.0.fn.norm.weight" % depth in P:
        depth += 1
    n = x.shape[1]
    dh = dim // heads
    for l in range(depth):
        pre = "transformer.layers.%d." % l
        # Residual(PreNorm(Attention))                                vit.py:84-99
        y = _layernorm(x, P, pre + "0.fn.norm", fin, fout)
        qkv = _linear(y, P, pre + "0.fn.fn.to_qkv", fin, fout)
        qkv = qkv.reshape(b, n, 3, heads, dh).permute(2, 0, 3, 1, 4)
        qh, kh, vh = qkv[0], qkv[1], qkv[2]
        attn = (qh @ kh.transpose(-2, -1)) * dim_head ** -0.5        # vit.py:70,92
        attn = attn.softmax(dim=-1)
        o = (attn @ vh).transpose(1, 2).reshape(b, n, dim)
        o = _linear(o, P, pre + "0.fn.fn.to_out.0", fin, fout)
        x = fq(o + x, ffn)                                           # vit.py:27
        # Residual(PreNorm(FeedForward))                              vit.py:44-56
        y = _layernorm(x, P, pre + "1.fn.norm", fin, fout)
        u = _linear(y, P, pre + "1.fn.fn.net.0", fin, fout)
        g = F.gelu(fq(u, fgelu))              # GELU output is never quantised (Appendix A)
        o = _linear(g, P, pre + "1.fn.fn.net.3", fin, fout)
        x = fq(o + x, ffn)

    if decoder == "classification":                                  # vit.py:335-342
        y = _layernorm(x[:, 0], P, "decoder.norm", fin, fout)
        return _linear(y, P, "decoder.linear", fin, fout)
    if decoder == "segmentation":                                    # vit.py:359-374
        y = _layernorm(x[:, 1:], P, "decoder.norm", fin, fout)
        y = _linear(y, P, "decoder.linear", fin, fout)
        bb, hw, cc = y.shape
        y = y.transpose(1, 2).view(bb, cc, gh, gw)
        size = image_size if image_size is not None else h
        return F.interpolate(y, size=size, mode="bilinear")
    if decoder == "detection":                                       # vit.py:389-396
        t = x[:, -num_det_tokens:, :]
        return {
            "pred_logits": _linear(t, P, "decoder.class_embed", fin, fout),
            "pred_boxes": _linear(t, P, "decoder.bbox_embed", fin, fout).sigmoid(),
        }
    raise ValueError(decoder)


def init_params(*, decoder, num_classes, dim, depth, heads, mlp_dim, patch_size=16, channels=3,
                num_det_tokens=100, seed=1234):
    """Random-init canonical parameters with the reference's distributions (randn tokens /
    pos-embedding, default nn.Linear / nn.LayerNorm init).  Deterministic in `seed`; it does
    not reproduce the reference constructor's RNG consumption order."""
    g = torch.Generator().manual_seed(seed)
    P = {}

    def lin(name, fin, fout):
        bound = 1.0 / np.sqrt(fin)
        P[name + ".weight"] = (torch.rand(fout, fin, generator=g) * 2 - 1) * bound
        P[name + ".bias"] = (torch.rand(fout, generator=g) * 2 - 1) * bound

    def ln(name):
        # the reference initialises gamma=1, beta=0; perturbed here so that tests exercise them
        P[name + ".weight"] = 1.0 + 0.1 * torch.randn(dim, generator=g)
        P[name + ".bias"] = 0.1 * torch.randn(dim, generator=g)

    P["pos_embedding"] = torch.randn(1, 14 * 14 + 1, dim, generator=g)
    P["pos_embedding_det"] = torch.randn(1, num_det_tokens, dim, generator=g)
    P["cls_token"] = torch.randn(1, 1, dim, generator=g)
    P["det_tokens"] = torch.randn(1, num_det_tokens, dim, generator=g)
    lin("patch_to_embedding", channels * patch_size ** 2, dim)
    for l in range(depth):
        pre = "transformer.layers.%d." % l
        ln(pre + "0.fn.norm")
        lin(pre + "0.fn.fn.to_qkv", dim, 3 * dim)
        lin(pre + "0.fn.fn.to_out.0", dim, dim)
        ln(pre + "1.fn.norm")
        lin(pre + "1.fn.fn.net.0", dim, mlp_dim)
        lin(pre + "1.fn.fn.net.3", mlp_dim, dim)
    if decoder in ("classification", "segmentation"):
        ln("decoder.norm")
        lin("decoder.linear", dim, num_classes)
    else:
        lin("decoder.class_embed", dim, num_classes + 1)
        lin("decoder.bbox_embed", dim, 4)
    return P


def train_step(P, img, target, *, decoder, heads, q_format, patch_size=16, num_det_tokens=100, grad_format=None,
               loss_scale=1.0):
    """zero_grad -> forward -> CrossEntropy -> backward (classification/train.py:239-264,
    segmentation/train.py:254-275).  Detection uses a fixed surrogate loss (sum of CE on
    pred_logits vs target['labels'] and L1 on pred_boxes vs target['boxes']) because the
    Hungarian criterion is host-side and out of scope.
    grad_format / loss_scale: the gradient-quantiser option (vit_forward) and the factor the loss is multiplied by
    before backward (the reference's GradScaler, classification/train.py:259-264); the returned gradients are those
    of loss * loss_scale.
    Returns (output, loss, grads dict)."""
    params = {k: v.detach().clone().requires_grad_(True) for k, v in P.items()}
    out = vit_forward(params, img, decoder=decoder, heads=heads, q_format=q_format,
                      patch_size=patch_size, num_det_tokens=num_det_tokens, grad_format=grad_format)
    if decoder == "detection":
        loss = (F.cross_entropy(out["pred_logits"].flatten(0, 1), target["labels"].flatten())
                + (out["pred_boxes"] - target["boxes"]).abs().mean())
    else:
        loss = F.cross_entropy(out, target)
    (loss * loss_scale).backward()
    grads = {k: v.grad for k, v in params.items()}
    return out, loss.detach(), grads
