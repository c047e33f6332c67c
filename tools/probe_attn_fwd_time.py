"""Warm timing of the attention forward at the flagship shape (B = 256, H = 6, N = 257), rotating inputs > L2.  MV_ALT_LIB picks a variant."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "myrtle-vision_b200")); sys.path.insert(0, ROOT)
import torch
import mv_native as mv
B, H, N = (int(a) for a in sys.argv[1:4]) if len(sys.argv) > 3 else (256, 6, 257)
D = H * 64
torch.manual_seed(0)
qs = [torch.randn(B * N, 3 * D, device="cuda").half() for _ in range(3)]
outs = [torch.empty(B * N, D, device="cuda", dtype=torch.float16) for _ in range(3)]
for rep in range(3):
    for i in range(6): mv.attention_fwd(qs[i % 3], B, H, N, q_out=(5, 10), out=outs[i % 3])
    e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(30): mv.attention_fwd(qs[i % 3], B, H, N, q_out=(5, 10), out=outs[i % 3])
    e1.record(); torch.cuda.synchronize()
    print("attention fwd B%d H%d N%d: %.1f us" % (B, H, N, e0.elapsed_time(e1) / 30 * 1e3), flush=True)
