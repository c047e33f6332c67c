"""gelu' dgrad GEMM (du = (dx W2) * gelu'(u), + fused fc1 bias gradient): correctness incl. ragged M, then timing at the
ViT-Small step shape with rotating buffers (cold L2)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "myrtle-vision_b200")); sys.path.insert(0, ROOT)
import torch
import mv_native as mv
dev, h = "cuda", torch.float16
torch.manual_seed(0)
def check(M, N, K, colsum):
    A = (torch.randn(M, K, device=dev) * 0.5).to(h); B = (torch.randn(N, K, device=dev) * 0.1).to(h)
    aux = torch.rand(M, N, device=dev).to(h)
    out = torch.full((M, N), float("nan"), device=dev, dtype=h)
    cs = torch.zeros(N, device=dev) if colsum else None
    mv.gemm(A, B, out, aux=aux, epilogue=mv.EPI_DGELU, colsum=cs)
    want = (A.double() @ B.double().t()) * aux.double()
    err = ((out.double() - want).abs().max() / want.abs().max()).item()
    msg = "M%-6d N%-5d K%-4d colsum=%d  rel %.2e" % (M, N, K, colsum, err)
    if colsum:
        msg += "  colsum rel %.2e" % ((cs.double() - out.double().sum(0)).abs().max() / out.double().sum(0).abs().max()).item()
    print(msg, "OK" if err < 2e-3 else "BAD", flush=True)
for M in (256, 1000, 4096 + 40, 8192, 16384, 65792, 65792):
    for cs in (0, 1):
        check(M, 1536, 384, cs)
check(777, 1500, 384, 1)        # generic path (ragged N)
M, N, K = 65792, 1536, 384
nb = 4
A = [(torch.randn(M, K, device=dev) * 0.5).to(h) for _ in range(nb)]
B = (torch.randn(N, K, device=dev) * 0.1).to(h)
aux = [torch.rand(M, N, device=dev).to(h) for _ in range(nb)]
out = [torch.empty(M, N, device=dev, dtype=h) for _ in range(nb)]
cs = torch.zeros(N, device=dev)
for mode, kw in (("gelu' + colsum", dict(colsum=cs)), ("gelu'", {})):
    for i in range(2 * nb): mv.gemm(A[i % nb], B, out[i % nb], aux=aux[i % nb], epilogue=mv.EPI_DGELU, **kw)
    e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(5 * nb): mv.gemm(A[i % nb], B, out[i % nb], aux=aux[i % nb], epilogue=mv.EPI_DGELU, **kw)
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / (5 * nb)
    print("%-16s %.1f us  %.0f GB/s of 455 MB" % (mode, ms * 1e3, 455.9e6 / ms / 1e6), flush=True)
