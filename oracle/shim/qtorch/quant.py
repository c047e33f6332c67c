"""qtorch.quant shim: Quantizer / quantizer / float_quantize / fixed_point_quantize /
block_quantize on CPU tensors, computed by oracle/quant_oracle.c.  TEST
INFRASTRUCTURE ONLY (see qtorch/__init__.py in this directory)."""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.dirname(
    os.path.abspath(__file__))))))
from oracle import quant_oracle as _qo  # noqa: E402

from . import BlockFloatingPoint, FixedPoint, FloatingPoint  # noqa: E402

_SEED = 1234
_offset = 0


def manual_seed(seed, offset=0):
    """(seed, offset) of the explicit Philox stream used for stochastic rounding."""
    global _SEED, _offset
    _SEED, _offset = int(seed), int(offset)


def _next_stream():
    global _offset
    s = (_SEED, _offset)
    _offset += 1
    return s


def _np(x):
    assert not x.is_cuda, "the oracle shim is CPU-only"
    return x.detach().contiguous().float().numpy()


def float_quantize(x, exp, man, rounding="stochastic"):
    a = _np(x)
    if rounding == "nearest":
        o = _qo.float_quantize(a, exp, man, "nearest")
    else:
        seed, off = _next_stream()
        o = _qo.float_quantize(a, exp, man, "stochastic", _qo.philox_bits(a.size, seed, off))
    return torch.from_numpy(o.reshape(a.shape))


def fixed_point_quantize(x, wl, fl, clamp=True, symmetric=False, rounding="stochastic"):
    a = _np(x)
    if rounding == "nearest":
        o = _qo.fixed_point_quantize(a, wl, fl, clamp, symmetric, "nearest")
    else:
        seed, off = _next_stream()
        o = _qo.fixed_point_quantize(a, wl, fl, clamp, symmetric, "stochastic",
                                     _qo.philox_uniform(a.size, seed, off))
    return torch.from_numpy(o.reshape(a.shape))


def block_quantize(x, wl, dim=-1, rounding="stochastic"):
    a = _np(x)
    if rounding == "nearest":
        o = _qo.block_quantize(a, wl, dim, "nearest")
    else:
        seed, off = _next_stream()
        o = _qo.block_quantize(a, wl, dim, "stochastic", _qo.philox_bits(a.size, seed, off))
    return torch.from_numpy(o.reshape(a.shape))


def _make_fn(number, rounding):
    if number is None:
        return lambda x: x
    if isinstance(number, FloatingPoint):
        return lambda x: float_quantize(x, number.exp, number.man, rounding)
    if isinstance(number, FixedPoint):
        return lambda x: fixed_point_quantize(x, number.wl, number.fl, number.clamp,
                                              number.symmetric, rounding)
    if isinstance(number, BlockFloatingPoint):
        return lambda x: block_quantize(x, number.wl, number.dim, rounding)
    raise ValueError("unknown number format {}".format(number))


def quantizer(forward_number=None, backward_number=None, forward_rounding="stochastic",
              backward_rounding="stochastic", clamping_grad_zero=False, backward_hooks=[]):
    for rounding in (forward_rounding, backward_rounding):
        assert rounding in ("stochastic", "nearest"), "invalid rounding type {:s}".format(rounding)
    fwd = _make_fn(forward_number, forward_rounding)
    bwd = _make_fn(backward_number, backward_rounding)
    use_mask = clamping_grad_zero and isinstance(forward_number, FixedPoint)

    class Rounding(torch.autograd.Function):
        @staticmethod
        def forward(ctx, x):
            if forward_number is None:
                return x
            if use_mask:
                a = _np(x)
                runif = None
                if forward_rounding == "stochastic":
                    seed, off = _next_stream()
                    runif = _qo.philox_uniform(a.size, seed, off)
                o, m = _qo.fixed_point_quantize_mask(a, forward_number.wl, forward_number.fl,
                                                     forward_number.symmetric, runif)
                ctx.mask = torch.from_numpy(m.reshape(a.shape).astype(np.bool_))
                return torch.from_numpy(o.reshape(a.shape))
            return fwd(x.contiguous())

        @staticmethod
        def backward(ctx, grad_output):
            if not ctx.needs_input_grad[0]:
                return None
            g = grad_output
            if backward_number is not None:
                g = bwd(g.contiguous())
            if use_mask:
                g = g.masked_fill(ctx.mask, 0.0)
            for hook in backward_hooks:
                g = hook(g)
            return g

    return Rounding.apply


class Quantizer(torch.nn.Module):
    def __init__(self, forward_number=None, backward_number=None, forward_rounding="stochastic",
                 backward_rounding="stochastic"):
        super().__init__()
        self.quantize = quantizer(forward_number, backward_number, forward_rounding,
                                  backward_rounding)

    def forward(self, x):
        return self.quantize(x)
