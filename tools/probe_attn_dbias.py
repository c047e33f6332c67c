"""Short-sequence attention backward with / without the fused to_qkv bias gradient: python tools/probe_attn_dbias.py"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "myrtle-vision_b200")); sys.path.insert(0, ROOT)
import torch
import mv_native as mv
B, H, N, dev = 256, 6, 257, "cuda"
D = H * 64
torch.manual_seed(0)
nb = 3
qkv = [torch.randn(B * N, 3 * D, device=dev).half() for _ in range(nb)]
do = [torch.randn(B * N, D, device=dev).half() for _ in range(nb)]
fw = [mv.attention_fwd(q, B, H, N) for q in qkv]
dqkv = torch.empty_like(qkv[0]); delta = torch.empty(B, H, N, device=dev)
dbias = torch.zeros(3 * D, device=dev)
def timeit(fn, n=12):
    for i in range(4): fn(i % nb)
    e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(n): fn(i % nb)
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n
for rep in range(2):
    a = timeit(lambda j: mv.attention_bwd(qkv[j], fw[j][0], do[j], fw[j][1], B, H, N, dqkv=dqkv, delta=delta))
    b = timeit(lambda j: mv.attention_bwd(qkv[j], fw[j][0], do[j], fw[j][1], B, H, N, dqkv=dqkv, delta=delta, dbias=dbias))
    c = timeit(lambda j: mv.colsum(dqkv, dbias))
    print("attn_bwd %.4f ms   with dbias %.4f ms   separate colsum %.4f ms" % (a, b, c))
