"""x + Linear(...) GEMMs (proj / fc2 forward: fp32 residual, fp32 output, FP16_32 and FP16_16 quantiser flags): correctness
incl. ragged M, then timing at the ViT-Small step shapes with rotating buffers."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "myrtle-vision_b200")); sys.path.insert(0, ROOT)
import numpy as np, torch
import mv_native as mv
if os.environ.get("MV_ALT_LIB"): mv._SO = os.path.join(ROOT, "myrtle-vision_b200", "csrc", os.environ["MV_ALT_LIB"])
from oracle import quant_oracle as qo
dev, h = "cuda", torch.float16
torch.manual_seed(0)
def check(M, N, K, quant):
    A = (torch.randn(M, K, device=dev) * 0.5).to(h); B = (torch.randn(N, K, device=dev) * 0.1).to(h)
    bias = torch.randn(N, device=dev); res = torch.randn(M, N, device=dev)
    out = torch.full((M, N), float("nan"), device=dev)
    kw = dict(q_out=(5, 10), q_res=(5, 10)) if quant else {}
    mv.gemm(A, B, out, bias=bias, residual=res, **kw)
    y = (A.double() @ B.double().t() + bias.double()).float()
    if quant:
        y = torch.from_numpy(qo.float_quantize(y.cpu().numpy(), 5, 10)).to(dev)
        want = torch.from_numpy(qo.float_quantize((y + res).cpu().numpy(), 5, 10)).to(dev)
        bad = int(((out - want).abs() > 2e-3 * want.abs().clamp(min=1.0)).sum())     # an fp16 ulp where the fp32 sums straddle a tie
    else:
        want = y + res
        bad = int(((out - want).abs() > 1e-4 * want.abs().clamp(min=1.0)).sum())
    print("M%-6d N%-4d K%-4d quant=%d  max err %.2e  outliers %d" % (M, N, K, quant, (out - want).abs().max().item(), bad),
          "OK" if bad == 0 else "BAD", flush=True)
for M in (256, 1000, 8192, 65792, 65792):
    for K in (384, 1536):
        for q in (0, 1):
            check(M, 384, K, q)
M, N = 65792, 384
nb = 4
for K in (384, 1536):
    A = [(torch.randn(M, K, device=dev) * 0.5).to(h) for _ in range(nb)]
    B = (torch.randn(N, K, device=dev) * 0.1).to(h)
    res = [torch.randn(M, N, device=dev) for _ in range(nb)]
    out = [torch.empty(M, N, device=dev) for _ in range(nb)]
    bias = torch.randn(N, device=dev)
    for i in range(2 * nb): mv.gemm(A[i % nb], B, out[i % nb], bias=bias, residual=res[i % nb])
    e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(5 * nb): mv.gemm(A[i % nb], B, out[i % nb], bias=bias, residual=res[i % nb])
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / (5 * nb)
    byts = M * K * 2 + 2 * M * N * 4
    print("K=%-4d + residual: %.1f us  %.0f GB/s" % (K, ms * 1e3, byts / ms / 1e6), flush=True)
