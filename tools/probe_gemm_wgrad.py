"""Split-K wgrad GEMMs of ViT-Small at batch 256 with different output tile widths: python tools/probe_gemm_wgrad.py"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "myrtle-vision_b200")); sys.path.insert(0, ROOT)
import torch
import mv_native as mv
T, dev, h = 65792, "cuda", torch.float16
torch.manual_seed(0)
def timeit(fn, n=12):
    for i in range(4): fn(i)
    e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(n): fn(i)
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n
for name, no, ki in (("fc1 [1536,384]", 1536, 384), ("qkv [1152,384]", 1152, 384), ("proj [384,384]", 384, 384),
                     ("fc2 [384,1536]", 384, 1536), ("patch [384,768]", 384, 768)):
    dY = [torch.randn(T, no, device=dev).half() for _ in range(3)]
    X = [torch.randn(T, ki, device=dev).half() for _ in range(3)]
    out = torch.zeros(no, ki, device=dev)
    line = []
    for tn in (0, 384, 256, 128):
        try:
            ms = timeit(lambda i: mv.gemm(dY[i % 3], X[i % 3], out, a_major=1, b_major=1, accumulate=True, tile_n=tn))
            line.append("tile_n %3d: %.4f ms (%4.0f TFLOP/s)" % (tn, ms, 2.0 * T * no * ki / ms / 1e9))
        except mv.MvError:
            line.append("tile_n %3d: n/a" % tn)
    out_t = torch.zeros(ki, no, device=dev)
    ms = timeit(lambda i: mv.gemm(X[i % 3], dY[i % 3], out_t, a_major=1, b_major=1, accumulate=True, transpose_out=True))
    line.append("as transpose: %.4f ms (%4.0f TFLOP/s)" % (ms, 2.0 * T * no * ki / ms / 1e9))
    print(name, " | ".join(line))
    del dY, X
