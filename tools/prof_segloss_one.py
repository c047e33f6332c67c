"""Profiling driver: a few launches of the fused upsample + cross-entropy kernel at config 4's shape."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "myrtle-vision_b200")); sys.path.insert(0, ROOT)
import torch
import mv_native as mv
torch.manual_seed(0)
B, C, g, H = 256, 17, 16, 256
y = torch.randn(B, g * g, C, device="cuda")
lab = torch.randint(0, C, (B, H, H), device="cuda")
for _ in range(4):
    out = mv.upsample_ce(y, lab)
torch.cuda.synchronize()
print("ok")
