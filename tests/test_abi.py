"""The C-ABI shared library loads on a CPU-only box and exports every symbol include/*.h declares
(no compute calls here)."""
import ctypes
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def lib():
    import __graft_entry__
    __graft_entry__.build()
    import mv_native
    return mv_native.lib()


def declared_symbols():
    text = open(os.path.join(ROOT, "include", "mv_b200.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(mv_[a-z0-9_]+)\s*\(", text)))


def test_header_declares_the_expected_surface():
    syms = declared_symbols()
    for s in ("mv_float_quantize", "mv_fixed_point_quantize", "mv_block_quantize", "mv_gemm",
              "mv_layernorm_q_fwd", "mv_layernorm_q_bwd", "mv_attention_fwd", "mv_attention_bwd",
              "mv_quantize_weight", "mv_last_error"):
        assert s in syms


def test_every_declared_symbol_is_exported(lib):
    for s in declared_symbols():
        assert hasattr(lib, s), "libmv_b200.so does not export " + s


def test_error_reporting_without_a_gpu(lib):
    lib.mv_last_error.restype = ctypes.c_char_p
    rc = lib.mv_float_quantize(None, None, 0, ctypes.c_int64(8), 9, 10, 0, ctypes.c_uint64(0),
                               ctypes.c_uint64(0), None)
    assert rc != 0
    assert b"exp_bits" in lib.mv_last_error()
    assert lib.mv_version() >= 100


def test_missing_library_fails_loudly(monkeypatch):
    import mv_native
    monkeypatch.setattr(mv_native, "_lib", None)
    monkeypatch.setattr(mv_native, "_SO", "/nonexistent/libmv_b200.so")
    with pytest.raises(mv_native.MvError, match="no CPU fallback"):
        mv_native.lib()
