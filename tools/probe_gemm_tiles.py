"""GELU / gelu' GEMM (M=65792, N=1536, K=384) with different N tiles: python tools/probe_gemm_tiles.py"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "myrtle-vision_b200")); sys.path.insert(0, ROOT)
import torch
import mv_native as mv
M, N, K, dev, h = 65792, 1536, 384, "cuda", torch.float16
torch.manual_seed(0)
nb = 4
A = [(torch.randn(M, K, device=dev) * 0.05).to(h) for _ in range(nb)]
B = torch.randn(N, K, device=dev).to(h)
out = [torch.empty(M, N, device=dev, dtype=h) for _ in range(nb)]
aux = [torch.rand(M, N, device=dev).to(h) for _ in range(nb)]
bias = torch.randn(N, device=dev)
def timeit(fn, n=12):
    for i in range(4): fn(i % nb)
    e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(n): fn(i % nb)
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n
for tn in (256, 192, 128):
    g = timeit(lambda j: mv.gemm(A[j], B, out[j], bias=bias, aux=aux[j], epilogue=mv.EPI_GELU, q_res=(5, 10), tile_n=tn))
    d = timeit(lambda j: mv.gemm(A[j], B, out[j], aux=aux[j], epilogue=mv.EPI_DGELU, tile_n=tn))
    p = timeit(lambda j: mv.gemm(A[j], B, out[j], bias=bias, tile_n=tn))
    print("tile_n %3d: gelu %.3f ms  dgelu %.3f ms  plain %.3f ms" % (tn, g, d, p))
