"""qtorch-shaped facade over the B200 fake-quant kernels (drop-in for QPyTorch 0.3.0's number
classes, which myrtle-vision imports at src/myrtle_vision/utils/quantize.py:4-6).

GPU-only: the functions in qtorch.quant raise on CPU tensors (no CPU fallback).
"""

__all__ = ["Number", "FloatingPoint", "FixedPoint", "BlockFloatingPoint"]


class Number:
    """Base class of the low-precision number formats."""


class FloatingPoint(Number):
    """Low-precision float with `exp` exponent bits and `man` mantissa bits."""

    def __init__(self, exp, man):
        if not 0 < exp <= 8:
            raise AssertionError("invalid bits for exponent:{}".format(exp))
        if not 0 < man <= 23:
            raise AssertionError("invalid bits for mantissa:{}".format(man))
        self.exp, self.man = exp, man

    def __str__(self):
        return "FloatingPoint (exponent={:d}, mantissa={:d})".format(self.exp, self.man)

    __repr__ = __str__


class FixedPoint(Number):
    """Fixed point with word length `wl` and fractional length `fl`."""

    def __init__(self, wl, fl, clamp=True, symmetric=False):
        if wl <= 0 or fl <= 0:
            raise AssertionError("invalid bits for word/fractional length: {}/{}".format(wl, fl))
        self.wl, self.fl, self.clamp, self.symmetric = wl, fl, clamp, symmetric

    def __str__(self):
        return "FixedPoint (wl={:d}, fl={:d})".format(self.wl, self.fl)

    __repr__ = __str__


class BlockFloatingPoint(Number):
    """Shared-exponent block format with word length `wl` over blocks along `dim`."""

    def __init__(self, wl, dim=-1):
        if wl <= 0:
            raise AssertionError("invalid bits for word length:{}".format(wl))
        self.wl, self.dim = wl, dim

    def __str__(self):
        return "BlockFloatingPoint (wl={:d}, dim={:d})".format(self.wl, self.dim)

    __repr__ = __str__
