"""fp16-output GEMMs without a second epilogue operand (qkv forward, plain dgrads): correctness incl. ragged M / N, then
timing at the ViT-Small step shapes with rotating buffers.  MV_ALT_LIB selects a variant library."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "myrtle-vision_b200")); sys.path.insert(0, ROOT)
import numpy as np, torch
import mv_native as mv
if os.environ.get("MV_ALT_LIB"): mv._SO = os.path.join(ROOT, "myrtle-vision_b200", "csrc", os.environ["MV_ALT_LIB"])
from oracle import quant_oracle as qo
dev, h = "cuda", torch.float16
torch.manual_seed(0)
def check(M, N, K, bias, quant):
    A = (torch.randn(M, K, device=dev) * 0.5).to(h); B = (torch.randn(N, K, device=dev) * 0.1).to(h)
    b = torch.randn(N, device=dev) if bias else None
    out = torch.full((M, N), float("nan"), device=dev, dtype=h)
    kw = dict(q_out=(5, 10)) if quant else {}
    mv.gemm(A, B, out, bias=b, **kw)
    y = (A.double() @ B.double().t() + (b.double() if bias else 0.0)).float()
    want = torch.from_numpy(qo.float_quantize(y.cpu().numpy(), 5, 10)).to(dev)     # fp16 rounding = (5,10) nearest in range
    bad = int(((out.float() - want).abs() > 2.1e-3 * want.abs().clamp(min=0.5)).sum()) + int(torch.isnan(out.float()).sum())
    print("M%-6d N%-5d K%-5d bias=%d q=%d  max err %.2e  outliers %d" % (M, N, K, bias, quant, (out.float() - want).abs().max().item(), bad),
          "OK" if bad == 0 else "BAD", flush=True)
if "time" not in sys.argv:
    for M in (256, 1000, 65792, 65792):
        check(M, 1152, 384, 1, 0); check(M, 1152, 384, 1, 1); check(M, 384, 384, 0, 0); check(M, 384, 1152, 0, 0); check(M, 384, 1536, 0, 0)
    check(777, 1100, 384, 1, 1); check(300, 128, 64, 1, 0); check(300, 256, 128, 0, 0)
M = 65792
nb = 4
for (N, K, bias) in ((1152, 384, 1), (384, 384, 0), (384, 1152, 0), (384, 1536, 0)):
    A = [(torch.randn(M, K, device=dev) * 0.5).to(h) for _ in range(nb)]
    B = (torch.randn(N, K, device=dev) * 0.1).to(h)
    out = [torch.empty(M, N, device=dev, dtype=h) for _ in range(nb)]
    b = torch.randn(N, device=dev) if bias else None
    for i in range(2 * nb): mv.gemm(A[i % nb], B, out[i % nb], bias=b)
    e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(5 * nb): mv.gemm(A[i % nb], B, out[i % nb], bias=b)
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / (5 * nb)
    print("N=%-5d K=%-5d: %.1f us  %.0f TFLOP/s  %.0f GB/s" % (N, K, ms * 1e3, 2.0 * M * N * K / ms / 1e9, (M * K * 2 + M * N * 2) / ms / 1e6), flush=True)
