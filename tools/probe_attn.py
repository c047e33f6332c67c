"""Bring-up probe: fused attention fwd/bwd vs an fp64 torch reference."""
import os, sys, traceback
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "myrtle-vision_b200")); sys.path.insert(0, ROOT)
import torch
import mv_native as mv
if os.environ.get("MV_ALT_LIB"):        # a variant built by tools/build_variant.sh
    mv._SO = os.path.join(ROOT, "myrtle-vision_b200", "csrc", os.environ["MV_ALT_LIB"])
dev = "cuda"
torch.manual_seed(0)
def rel(a, b):
    return ((a.double() - b.double()).abs().max() / (b.double().abs().max() + 1e-30)).item()
def case(B, H, N, gscale=1.0):
    D = H * 64
    qkv = (torch.randn(B * N, 3 * D, device=dev) * 1.0).half()
    out, lse = mv.attention_fwd(qkv, B, H, N)
    torch.cuda.synchronize()
    x = qkv.double().reshape(B, N, 3, H, 64).permute(2, 0, 3, 1, 4)
    q, k, v = [t.clone().requires_grad_(True) for t in (x[0], x[1], x[2])]
    s = (q @ k.transpose(-2, -1)) * 0.125
    pr = s.softmax(-1)
    o = (pr @ v)
    oref = o.transpose(1, 2).reshape(B * N, D)
    lse_ref = torch.logsumexp(s, -1) * 1.4426950408889634
    print("B%d H%d N%d fwd: out rel err %.3e  lse abs err %.3e  nan %d" % (B, H, N, rel(out, oref), (lse.double() - lse_ref).abs().max().item(), int(torch.isnan(out.float()).sum())), flush=True)
    do = (torch.randn(B * N, D, device=dev) * gscale).half()
    oref.backward(do.double())
    dqkv = mv.attention_bwd(qkv, out, do, lse, B, H, N)
    torch.cuda.synchronize()
    g = dqkv.double().reshape(B, N, 3, H, 64).permute(2, 0, 3, 1, 4)
    print("   bwd: dq rel %.3e  dk rel %.3e  dv rel %.3e  nan %d" % (rel(g[0], q.grad), rel(g[1], k.grad), rel(g[2], v.grad), int(torch.isnan(dqkv.float()).sum())), flush=True)
for args in [(1, 1, 128), (1, 1, 64), (2, 2, 257), (40, 6, 257), (2, 2, 258), (3, 3, 260), (2, 1, 261), (1, 1, 300), (2, 3, 197), (1, 2, 1000)]:
    try:
        case(*args)
    except Exception as e:
        print("EXC", args, repr(e)); traceback.print_exc(); break
def timeit(fn, n=10):
    for _ in range(3): fn()
    e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n
for (B, H, N) in [(256, 6, 257), (8, 6, 2501)]:
    D = H * 64
    qkv = torch.randn(B * N, 3 * D, device=dev).half()
    out, lse = mv.attention_fwd(qkv, B, H, N)
    do = torch.randn(B * N, D, device=dev).half()
    dqkv = torch.empty_like(qkv); delta = torch.empty(B, H, N, device=dev)
    ms = timeit(lambda: mv.attention_fwd(qkv, B, H, N, out=out, lse=lse))
    fl = 4.0 * B * H * N * N * 64
    print("attn fwd B%d H%d N%d: %.3f ms  %.1f TFLOP/s (algorithmic)" % (B, H, N, ms, fl / ms / 1e9), flush=True)
    ms = timeit(lambda: mv.attention_bwd(qkv, out, do, lse, B, H, N, dqkv=dqkv, delta=delta))
    print("attn bwd: %.3f ms  %.1f TFLOP/s (algorithmic 2.5x fwd)" % (ms, 2.5 * fl / ms / 1e9), flush=True)
    q4 = qkv.reshape(B, N, 3, H, 64).permute(2, 0, 3, 1, 4)
    ms = timeit(lambda: torch.nn.functional.scaled_dot_product_attention(q4[0], q4[1], q4[2]))
    print("   torch SDPA fwd: %.3f ms" % ms, flush=True)
