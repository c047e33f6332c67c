"""Profiling driver: a few launches of the HBM-bound kernels at the flagship shape.
usage: prof_elementwise_one.py ln_fwd|ln_bwd|quant|colsum"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "myrtle-vision_b200")); sys.path.insert(0, ROOT)
import torch
import mv_native as mv
what = sys.argv[1]
rows, D, dev = 65792, 384, "cuda"
torch.manual_seed(0)
if what == "quant":
    n = 1 << 28
    x = torch.randn(n, device=dev) * torch.exp(torch.empty(n, device=dev).uniform_(-12, 8))
    out = torch.empty_like(x)
    for _ in range(3):
        mv.float_quantize(x, 5, 10, "nearest", out=out)
elif what == "colsum":
    x = torch.randn(rows, 1536, device=dev).half()
    o = torch.zeros(1536, device=dev)
    for _ in range(3):
        mv.colsum(x, o)
else:
    xs = [torch.randn(rows, D, device=dev) for _ in range(3)]
    g = torch.ones(D, device=dev); b = torch.zeros(D, device=dev)
    dy = torch.randn(rows, D, device=dev).half()
    dres = torch.randn(rows, D, device=dev)
    dg = torch.zeros(D, device=dev); db = torch.zeros(D, device=dev); dp = torch.zeros(D, device=dev)
    for x in xs:
        y, mean, rstd = mv.layernorm_q_fwd(x, g, b, q_in=(5, 10), q_post=(5, 10))
        if what == "ln_bwd":
            mv.layernorm_q_bwd(dy, x, g, mean, rstd, dres=dres, q_in=(5, 10), dgamma=dg, dbeta=db, dbias_prev=dp)
torch.cuda.synchronize()
print("ok")
