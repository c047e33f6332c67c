"""Fake-quant kernel bandwidth on 2^28 fp32 values for two input distributions: plain randn (every value in fp16's
normal range) and SURVEY.md §8d's randn * exp(U(-12, 8)) (11 % below 2^-14, some saturating): python tools/probe_quant_inputs.py"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "myrtle-vision_b200")); sys.path.insert(0, ROOT)
import torch
import mv_native as mv
dev, n = "cuda", 1 << 28
g = torch.Generator(device=dev).manual_seed(1234)
x = torch.randn(n, device=dev, generator=g)
wide = x * torch.exp(torch.empty(n, device=dev).uniform_(-12, 8, generator=g))
out = torch.empty_like(x)
def gbs(inp, **kw):
    for _ in range(3): mv.float_quantize(inp, 5, 10, out=out, **kw)
    e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(10): mv.float_quantize(inp, 5, 10, out=out, **kw)
    e1.record(); torch.cuda.synchronize()
    return 8.0 * n * 10 / e0.elapsed_time(e1) / 1e6
for name, inp in (("randn", x), ("randn * exp(U(-12, 8))", wide)):
    print("%-24s nearest %.0f GB/s   stochastic %.0f GB/s   (below 2^-14: %.1f %%)" % (
        name, gbs(inp), gbs(inp, rounding="stochastic", seed=1, offset=0),
        100.0 * float((inp.abs() < 2.0 ** -14).float().mean())))
o16 = torch.empty(n, device=dev, dtype=torch.float16)
def gbs16(inp):
    kw = dict(rounding="stochastic", seed=1, offset=0, out=o16)
    for _ in range(3): mv.float_quantize(inp, 5, 10, **kw)
    e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(10): mv.float_quantize(inp, 5, 10, **kw)
    e1.record(); torch.cuda.synchronize()
    return 6.0 * n * 10 / e0.elapsed_time(e1) / 1e6
print("stochastic -> fp16 container: randn %.0f GB/s, wide %.0f GB/s (6 B/element)" % (gbs16(x), gbs16(wide)))
for _ in range(3): out.copy_(x)
e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(10): out.copy_(x)
e1.record(); torch.cuda.synchronize()
print("torch copy %.0f GB/s" % (8.0 * n * 10 / e0.elapsed_time(e1) / 1e6))
