"""Profiling driver: a few launches of attention fwd (+bwd) at the flagship shape (or, with `long`, at config 5's:
B = 8, N = 2501).  usage: prof_attn_one.py [bwd] [long]"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "myrtle-vision_b200")); sys.path.insert(0, ROOT)
import torch
import mv_native as mv
B, H, N = (8, 6, 2501) if "long" in sys.argv else (256, 6, 257)
D = H * 64
torch.manual_seed(0)
qkv = torch.randn(B * N, 3 * D, device="cuda").half()
do = torch.randn(B * N, D, device="cuda").half()
dbias = torch.zeros(3 * D, device="cuda")
for _ in range(3):
    out, lse = mv.attention_fwd(qkv, B, H, N, q_out=(5, 10))
    if "bwd" in sys.argv:
        mv.attention_bwd(qkv, out, do, lse, B, H, N, dbias=dbias)     # with the fused to_qkv bias gradient, as in the step
torch.cuda.synchronize()
print("ok")
