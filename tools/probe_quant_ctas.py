import os, sys
sys.path.insert(0, "/root/repo/myrtle-vision_b200"); sys.path.insert(0, "/root/repo")
import torch, mv_native as mv
nn = 1 << 28
big = torch.randn(nn, device="cuda")
for odt, bpe in ((torch.float32, 8), (torch.float16, 6)):
    o = torch.empty(nn, device="cuda", dtype=odt)
    for c in (2, 3, 4, 5, 6, 8, 12):
        mv.set_option("quant_ctas", c)
        for mode in ("nearest", "stochastic"):
            for _ in range(3): mv.float_quantize(big, 5, 10, mode, out=o)
            e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(10): mv.float_quantize(big, 5, 10, mode, out=o)
            e1.record(); torch.cuda.synchronize()
            ms = e0.elapsed_time(e1) / 10
            print("ctas/SM %2d %s -> %s: %.0f GB/s" % (c, mode, str(odt)[6:], nn * bpe / ms / 1e6), flush=True)
