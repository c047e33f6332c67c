"""qtorch-API-compatible CPU shim backed by oracle/quant_oracle.c.

TEST INFRASTRUCTURE ONLY.  It exists so that the unmodified reference package
(/root/reference/src/myrtle_vision, which does `from qtorch import FixedPoint,
FloatingPoint` and `from qtorch.quant import Quantizer`,
utils/quantize.py:4-6) imports in this container, making the reference Python
itself the model-level oracle and the generator of tests/golden/*.  Number
classes mirror QPyTorch 0.3.0's public constructors.
"""


class Number:
    pass


class FloatingPoint(Number):
    def __init__(self, exp, man):
        assert 8 >= exp > 0, "invalid bits for exponent:{}".format(exp)
        assert 23 >= man > 0, "invalid bits for mantissa:{}".format(man)
        self.exp = exp
        self.man = man

    def __repr__(self):
        return "FloatingPoint (exponent={:d}, mantissa={:d})".format(self.exp, self.man)


class FixedPoint(Number):
    def __init__(self, wl, fl, clamp=True, symmetric=False):
        assert wl > 0 and fl > 0
        self.wl = wl
        self.fl = fl
        self.clamp = clamp
        self.symmetric = symmetric

    def __repr__(self):
        return "FixedPoint (wl={:d}, fl={:d})".format(self.wl, self.fl)


class BlockFloatingPoint(Number):
    def __init__(self, wl, dim=-1):
        assert wl > 0
        self.wl = wl
        self.dim = dim

    def __repr__(self):
        return "BlockFloatingPoint (wl={:d}, dim={:d})".format(self.wl, self.dim)
