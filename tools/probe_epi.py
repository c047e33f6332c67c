"""Bring-up probe: cost of each GEMM epilogue variant on the FC1 shape (cold operands)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "myrtle-vision_b200")); sys.path.insert(0, ROOT)
import torch
import mv_native as mv
dev = "cuda"; h = torch.float16
M, N, K, nbuf = 65792, 1536, 384, 5
As = [torch.randn(M, K, device=dev).to(h) for _ in range(nbuf)]
Bm = (torch.randn(N, K, device=dev) * 0.05).to(h); bias = torch.randn(N, device=dev)
outs = [torch.empty(M, N, device=dev, dtype=h) for _ in range(nbuf)]
auxs = [torch.empty(M, N, device=dev, dtype=h) for _ in range(nbuf)]
def run(name, fn):
    for i in range(nbuf): fn(i)
    e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(3 * nbuf): fn(i % nbuf)
    e1.record(); torch.cuda.synchronize()
    print("%-40s %.3f ms" % (name, e0.elapsed_time(e1) / (3 * nbuf)), flush=True)
for tn in (0, 128):
    run("tile_n=%d plain" % tn, lambda j: mv.gemm(As[j], Bm, outs[j], tile_n=tn))
    run("tile_n=%d bias" % tn, lambda j: mv.gemm(As[j], Bm, outs[j], bias=bias, tile_n=tn))
    run("tile_n=%d bias+q_res" % tn, lambda j: mv.gemm(As[j], Bm, outs[j], bias=bias, q_res=(5, 10), tile_n=tn))
    run("tile_n=%d bias+q_res(4,3)generic" % tn, lambda j: mv.gemm(As[j], Bm, outs[j], bias=bias, q_res=(4, 3), tile_n=tn))
    run("tile_n=%d gelu (no q)" % tn, lambda j: mv.gemm(As[j], Bm, outs[j], bias=bias, aux=auxs[j], epilogue=mv.EPI_GELU, tile_n=tn))
    run("tile_n=%d gelu + q_res" % tn, lambda j: mv.gemm(As[j], Bm, outs[j], bias=bias, aux=auxs[j], epilogue=mv.EPI_GELU, q_res=(5, 10), tile_n=tn))
    run("tile_n=%d dgelu" % tn, lambda j: mv.gemm(As[j], Bm, outs[j], aux=auxs[j], epilogue=mv.EPI_DGELU, tile_n=tn))
