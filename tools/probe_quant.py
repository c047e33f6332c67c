"""Bring-up probe: quant kernels bit-exactness vs the oracle + achieved bandwidth."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "myrtle-vision_b200")); sys.path.insert(0, ROOT)
import numpy as np, torch
import mv_native as mv
from oracle import quant_oracle as qo
dev = "cuda"
g = torch.Generator(device="cpu").manual_seed(1234)
n = 1 << 20
x = (torch.randn(n, generator=g) * torch.exp(torch.empty(n).uniform_(-12, 8, generator=g)))
x[:8] = torch.tensor([0.0, -0.0, float("inf"), -float("inf"), 65504.0, 65520.0, 2.0**-25, -2.0**-25])
xd = x.to(dev)
for (e, m) in [(5, 10), (8, 10), (4, 3), (5, 2), (8, 7)]:
    got = mv.float_quantize(xd, e, m).cpu().numpy()
    want = qo.float_quantize(x.numpy(), e, m)
    print("float nearest (%d,%d): mismatches %d" % (e, m, int((got.view(np.uint32) != want.view(np.uint32)).sum())))
    r = qo.philox_bits(n, 77, 5, half=m >= 7)
    got = mv.float_quantize(xd, e, m, "stochastic", seed=77, offset=5).cpu().numpy()
    want = qo.float_quantize(x.numpy(), e, m, "stochastic", r)
    print("float stochastic (%d,%d): mismatches %d" % (e, m, int((got.view(np.uint32) != want.view(np.uint32)).sum())))
print("philox stream mismatches", int((mv.philox_bits(1001, 77, 5).cpu().numpy().view(np.uint32) != qo.philox_bits(1001, 77, 5)).sum()))
got = mv.float_quantize(xd, 5, 10, out_dtype=torch.float16).float().cpu().numpy(); want = qo.float_quantize(x.numpy(), 5, 10)
print("float nearest (5,10) -> f16 container mismatches", int((got.view(np.uint32) != want.view(np.uint32)).sum()))
xs = (x * 1e-2).clamp(-10, 10); xsd = xs.to(dev)
for fl in (9, 8, 7):
    got = mv.fixed_point_quantize(xsd, 11, fl).cpu().numpy(); want = qo.fixed_point_quantize(xs.numpy(), 11, fl)
    print("fixed nearest (11,%d): mismatches %d" % (fl, int((got != want).sum())))
    got = mv.fixed_point_quantize(xsd, 11, fl, rounding="stochastic", seed=3, offset=1).cpu().numpy()
    want = qo.fixed_point_quantize(xs.numpy(), 11, fl, rounding="stochastic", runif=qo.philox_uniform(n, 3, 1))
    print("fixed stochastic (11,%d): mismatches %d" % (fl, int((got != want).sum())))
xb = x[8: 8 + 64 * 96 * 40].reshape(64, 96, 40); xbd = xb.to(dev)
for dim in (-1, 0, 1, 2):
    got = mv.block_quantize(xbd, 8, dim).cpu().numpy(); want = qo.block_quantize(xb.numpy(), 8, dim)
    print("block nearest wl8 dim%d: mismatches %d" % (dim, int((got.view(np.uint32) != want.view(np.uint32)).sum())))
w = torch.randn(1152, 384, generator=g); wd = w.to(dev)
q, qt = mv.quantize_weight(wd, 5, 10)
want = qo.float_quantize(w.numpy(), 5, 10)
print("weight quant mismatches", int((q.float().cpu().numpy() != want).sum()), int((qt.float().cpu().numpy() != want.T).sum()))
# bandwidth
for nn, label in [(1 << 28, "1GiB")]:
    big = torch.randn(nn, device=dev); out = torch.empty_like(big)
    for odt, bpe in ((torch.float32, 8), (torch.float16, 6)):
        o = out if odt == torch.float32 else torch.empty(nn, device=dev, dtype=odt)
        for mode in ("nearest", "stochastic"):
            for _ in range(3): mv.float_quantize(big, 5, 10, mode, out=o)
            e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(10): mv.float_quantize(big, 5, 10, mode, out=o)
            e1.record(); torch.cuda.synchronize()
            ms = e0.elapsed_time(e1) / 10
            print("quant %s %s -> %s: %.3f ms  %.0f GB/s" % (label, mode, str(odt)[6:], ms, nn * bpe / ms / 1e6), flush=True)
    e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
    for _ in range(3): out.copy_(big)
    e0.record()
    for _ in range(10): out.copy_(big)
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 10
    print("torch copy %s: %.3f ms %.0f GB/s" % (label, ms, nn * 8 / ms / 1e6), flush=True)
    del big, out
# block quantise at the bench shapes (12 algorithmic bytes per element: max pass + quant pass)
x2 = (torch.randn(16384, 16384, device=dev) * torch.exp(torch.empty(16384, 16384, device=dev).uniform_(-12, 8)))
o2 = torch.empty_like(x2)
for label, kw in [("dim=-1 nearest", dict(dim=-1)), ("dim=0 nearest", dict(dim=0)), ("dim=0 stochastic", dict(dim=0, rounding="stochastic", seed=5, offset=1)),
                  ("dim=1 nearest", dict(dim=1)), ("dim=1 stochastic", dict(dim=1, rounding="stochastic", seed=5, offset=1))]:
    for _ in range(3): mv.block_quantize(x2, 8, out=o2, **kw)
    e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(10): mv.block_quantize(x2, 8, out=o2, **kw)
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 10
    print("block wl=8 %s: %.3f ms  %.0f GB/s" % (label, ms, x2.numel() * 12 / ms / 1e6), flush=True)
