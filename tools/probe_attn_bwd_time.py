"""Attention backward at the benchmark shape, with and without the fused bias gradient, 30 timed launches each."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "myrtle-vision_b200")); sys.path.insert(0, ROOT)
import torch
import mv_native as mv
if os.environ.get("MV_ALT_LIB"):
    mv._SO = os.path.join(ROOT, "myrtle-vision_b200", "csrc", os.environ["MV_ALT_LIB"])
B, H, N = 256, 6, 257
D = H * 64
torch.manual_seed(0)
qkv = torch.randn(B * N, 3 * D, device="cuda").half()
do = torch.randn(B * N, D, device="cuda").half()
out, lse = mv.attention_fwd(qkv, B, H, N, q_out=(5, 10))
dqkv = torch.empty_like(qkv)
dbias = torch.zeros(3 * D, device="cuda")
def timeit(fn, n=30):
    for _ in range(5): fn()
    e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n
print("bwd without dbias: %.4f ms" % timeit(lambda: mv.attention_bwd(qkv, out, do, lse, B, H, N, dqkv=dqkv)))
print("bwd with    dbias: %.4f ms" % timeit(lambda: mv.attention_bwd(qkv, out, do, lse, B, H, N, dqkv=dqkv, dbias=dbias)))
print("fwd              : %.4f ms" % timeit(lambda: mv.attention_fwd(qkv, B, H, N, q_out=(5, 10), out=out, lse=lse)))
