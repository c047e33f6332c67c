"""ctypes front-end of oracle/quant_oracle.c (CPU restatement of QPyTorch 0.3.0).

TEST INFRASTRUCTURE ONLY — see the header of quant_oracle.c.  PARITY UNPINNED:
the reference holds no golden vectors for the QPyTorch boundary and qtorch's
source is not available offline.

Reference call sites restated: src/myrtle_vision/utils/quantize.py:47-72
(which formats), :84 (how the quantiser is invoked on X.data.float()).
"""
import ctypes
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "libquant_oracle.so")
_SRC = os.path.join(_HERE, "quant_oracle.c")
_lib = None


def build(force=False):
    """gcc-compile the C restatement next to its source (oracle/libquant_oracle.so)."""
    if (not force and os.path.exists(_SO)
            and os.path.getmtime(_SO) >= os.path.getmtime(_SRC)):
        return _SO
    cmd = ["gcc", "-O2", "-fPIC", "-shared", "-ffp-contract=off", "-fno-fast-math",
           "-o", _SO, _SRC, "-lm"]
    subprocess.check_call(cmd)
    return _SO


def lib():
    global _lib
    if _lib is None:
        build()
        _lib = ctypes.CDLL(_SO)
        _lib.mvo_set_switch.argtypes = [ctypes.c_char_p, ctypes.c_int]
        _lib.mvo_set_switch.restype = ctypes.c_int
    return _lib


def set_switch(name, value):
    if lib().mvo_set_switch(name.encode(), int(value)) != 0:
        raise KeyError(name)


def _f32(x):
    a = np.ascontiguousarray(np.asarray(x, dtype=np.float32))
    return a


def _p(a):
    return ctypes.c_void_p(a.ctypes.data)


def float_quantize(x, exp, man, rounding="nearest", rbits=None):
    """qtorch.quant.float_quantize(x, exp, man, rounding).  For stochastic
    rounding the 32 random bits per element are an explicit input."""
    a = _f32(x)
    o = np.empty_like(a)
    if rounding == "nearest":
        lib().mvo_float_quantize_nearest(_p(a), _p(o), ctypes.c_int64(a.size), int(man), int(exp))
    else:
        r = np.ascontiguousarray(np.asarray(rbits, dtype=np.uint32))
        assert r.size == a.size
        lib().mvo_float_quantize_stochastic(_p(a), _p(r), _p(o), ctypes.c_int64(a.size),
                                            int(man), int(exp))
    return o


def fixed_point_quantize(x, wl, fl, clamp=True, symmetric=False, rounding="nearest", runif=None):
    a = _f32(x)
    o = np.empty_like(a)
    if rounding == "nearest":
        lib().mvo_fixed_point_quantize_nearest(_p(a), _p(o), ctypes.c_int64(a.size), int(wl),
                                               int(fl), int(clamp), int(symmetric))
    else:
        r = _f32(runif)
        assert r.size == a.size
        lib().mvo_fixed_point_quantize_stochastic(_p(a), _p(r), _p(o), ctypes.c_int64(a.size),
                                                  int(wl), int(fl), int(clamp), int(symmetric))
    return o


def fixed_point_quantize_mask(x, wl, fl, symmetric=False, runif=None):
    a = _f32(x)
    o = np.empty_like(a)
    m = np.empty(a.shape, dtype=np.uint8)
    r = None if runif is None else _f32(runif)
    lib().mvo_fixed_point_quantize_mask(_p(a), _p(r) if r is not None else None, _p(o), _p(m),
                                        ctypes.c_int64(a.size), int(wl), int(fl), int(symmetric))
    return o, m


def block_quantize(x, wl, dim=-1, rounding="nearest", rbits=None):
    a = _f32(x)
    o = np.empty_like(a)
    r = None
    if rounding != "nearest":
        r = np.ascontiguousarray(np.asarray(rbits, dtype=np.uint32))
        assert r.size == a.size
    if dim is None or dim < 0:
        outer, dsize, inner, whole = 1, 1, a.size, 1
    else:
        shape = a.shape
        outer = int(np.prod(shape[:dim], dtype=np.int64))
        dsize = int(shape[dim])
        inner = int(np.prod(shape[dim + 1:], dtype=np.int64))
        whole = 0
    lib().mvo_block_quantize(_p(a), _p(r) if r is not None else None, _p(o),
                             ctypes.c_int64(outer), ctypes.c_int64(dsize), ctypes.c_int64(inner),
                             int(whole), int(wl))
    return o


def philox_bits(n, seed, offset=0, half=False):
    """32 random bits per element (word i&3 of philox(i>>2)); half=True: the 16-bit stream the CUDA float_quantize uses
    for man >= 7 (half-word i&7 of philox(i>>3)) — include/mv_b200.h."""
    r = np.empty(int(n), dtype=np.uint32)
    fn = lib().mvo_philox_bits16 if half else lib().mvo_philox_bits
    fn(_p(r), ctypes.c_int64(n), ctypes.c_uint64(seed), ctypes.c_uint64(offset))
    return r


def philox_uniform(n, seed, offset=0):
    r = np.empty(int(n), dtype=np.float32)
    lib().mvo_philox_uniform(_p(r), ctypes.c_int64(n), ctypes.c_uint64(seed), ctypes.c_uint64(offset))
    return r


# --------------------------------------------------------------------------
# Independent numpy restatement of float_quantize nearest (vectorised), used
# by tests to cross-check the C file and as the fast path for big tensors.
def float_quantize_nearest_np(x, exp, man):
    a = _f32(x)
    t = a.view(np.uint32)
    texp = ((t << np.uint32(1)) >> np.uint32(24)).astype(np.int32) - 127
    min_exp = -((1 << (exp - 1)) - 2)
    mask = np.uint32((1 << (23 - man)) - 1)
    half = np.uint32(1 << (23 - man - 1))
    sub = texp < min_exp
    # subnormal branch
    shift = (np.uint32((127 + min_exp) << 23) | (t & np.uint32(0x80000000))).view(np.float32)
    with np.errstate(over="ignore", invalid="ignore"):
        val = (a + shift).astype(np.float32)
        qs = ((val.view(np.uint32) + half) & ~mask).view(np.float32)
        out_sub = (qs - shift).astype(np.float32)
    # normal branch
    q = (t + half) & ~mask
    e = ((q << np.uint32(1)) >> np.uint32(24)).astype(np.int32)
    max_e = (1 << (exp - 1)) - 1 + 127
    max_man = np.uint32((0x007FFFFF >> (23 - man)) << (23 - man))
    sat = (t & np.uint32(0x80000000)) | np.uint32(max_e << 23) | max_man
    q = np.where((e > max_e) & (q != 0), sat, q).astype(np.uint32)
    return np.where(sub, out_sub, q.view(np.float32)).astype(np.float32)
