"""Bring-up probe (not a test): fc1 + GELU epilogue of the CTA-pair GEMM at the flagship shape — error statistics against
fp64 and warm timings (rotating buffers > L2).  Usage: python tools/probe_gelu.py"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "myrtle-vision_b200")); sys.path.insert(0, ROOT)
import torch
import torch.nn.functional as F
import mv_native as mv
if os.environ.get("MV_ALT_LIB"): mv._SO = os.path.join(ROOT, "myrtle-vision_b200", "csrc", os.environ["MV_ALT_LIB"])

torch.manual_seed(0)
dev, h = "cuda", torch.float16
M, N, K = 65792, 1536, 384
nbuf = 3
A = [(torch.randn(M, K, device=dev)).to(h) for _ in range(nbuf)]
SIG = float(sys.argv[1]) if len(sys.argv) > 1 else 0.4          # std of u = fc1's output (0.4: the benchmark's random init)
B = (torch.randn(N, K, device=dev) * (SIG / K ** 0.5)).to(h)
bias = torch.randn(N, device=dev) * 0.02
out = [torch.empty(M, N, device=dev, dtype=h) for _ in range(nbuf)]
aux = [torch.empty(M, N, device=dev, dtype=h) for _ in range(nbuf)]


def rel(got, want):
    return ((got.double() - want.double()).abs().max() / (want.double().abs().max() + 1e-30)).item()


for name, kw in (("FP16_32", dict(q_res=(5, 10))), ("FP16_16", dict(q_out=(5, 10), q_res=(5, 10)))):
    mv.gemm(A[0], B, out[0], bias=bias, aux=aux[0], epilogue=mv.EPI_GELU, **kw)
    rows = slice(0, 8192)
    lin = (A[0][rows].double() @ B.double().t() + bias.double())
    if "q_out" in kw:
        lin = mv.float_quantize(lin.float(), 5, 10).double()
    uu = lin.clone().requires_grad_(True)
    F.gelu(uu).sum().backward()
    want = mv.float_quantize(F.gelu(lin).float(), 5, 10)
    got = out[0][rows].float()
    print("%s  h rel %.2e flips %.5f   gelu' rel %.2e   tail(|h| < 6.1e-5): %d values, max abs err %.2e" % (
        name, rel(got, want), ((got - want).abs() > 0).float().mean().item(), rel(aux[0][rows], uu.grad),
        int((want.abs() < 6.1e-5).sum()), ((got - want).abs() * (want.abs() < 6.1e-5)).max().item()), flush=True)
    for rep in range(2):
        ev = [torch.cuda.Event(enable_timing=True) for _ in range(2)]
        for i in range(6):
            mv.gemm(A[i % nbuf], B, out[i % nbuf], bias=bias, aux=aux[i % nbuf], epilogue=mv.EPI_GELU, **kw)
        ev[0].record()
        n = 30
        for i in range(n):
            mv.gemm(A[i % nbuf], B, out[i % nbuf], bias=bias, aux=aux[i % nbuf], epilogue=mv.EPI_GELU, **kw)
        ev[1].record(); torch.cuda.synchronize()
        print("sigma %.2f %s  %.1f us per launch" % (SIG, name, ev[0].elapsed_time(ev[1]) / n * 1e3), flush=True)
