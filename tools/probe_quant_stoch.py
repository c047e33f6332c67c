"""Bring-up probe: stochastic float_quantize(5,10) bandwidth on the narrow (randn) and SURVEY 8d wide-range inputs, fp32 and fp16
containers.  MV_ALT_LIB selects a variant library."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "myrtle-vision_b200")); sys.path.insert(0, ROOT)
import torch
import mv_native as mv
dev, nn = "cuda", 1 << 28
torch.manual_seed(1234)
for name, big in (("randn", torch.randn(nn, device=dev)),
                  ("wide ", torch.randn(nn, device=dev) * torch.exp(torch.empty(nn, device=dev).uniform_(-12, 8)))):
    for odt, bpe in ((torch.float32, 8), (torch.float16, 6)):
        o = torch.empty(nn, device=dev, dtype=odt)
        for fmt in ((5, 10), (8, 10)) if odt == torch.float32 else ((5, 10),):
            for _ in range(3): mv.float_quantize(big, fmt[0], fmt[1], "stochastic", seed=3, offset=1, out=o)
            e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(10): mv.float_quantize(big, fmt[0], fmt[1], "stochastic", seed=3, offset=1, out=o)
            e1.record(); torch.cuda.synchronize()
            ms = e0.elapsed_time(e1) / 10
            print("%s stochastic %r -> %s: %.3f ms  %.0f GB/s" % (name, fmt, str(odt)[6:], ms, nn * bpe / ms / 1e6), flush=True)
        del o
    for odt, bpe in ((torch.float32, 8), (torch.float16, 6)):
        o = torch.empty(nn, device=dev, dtype=odt)
        for _ in range(3): mv.float_quantize(big, 5, 10, out=o)
        e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(10): mv.float_quantize(big, 5, 10, out=o)
        e1.record(); torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / 10
        print("%s nearest (5, 10) -> %s: %.3f ms  %.0f GB/s" % (name, str(odt)[6:], ms, nn * bpe / ms / 1e6), flush=True)
        del o
