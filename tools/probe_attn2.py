"""Bring-up probe: short-sequence attention kernels (attention_sn.cu) vs an fp64 torch reference and
vs the blocked kernels (set_option attn_sn 0).  usage: probe_attn2.py [fwd] [quick]"""
import os, sys, traceback
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "myrtle-vision_b200")); sys.path.insert(0, ROOT)
import torch
import mv_native as mv
dev = "cuda"
torch.manual_seed(0)
fwd_only = "fwd" in sys.argv
quick = "quick" in sys.argv
def rel(a, b):
    return ((a.double() - b.double()).abs().max() / (b.double().abs().max() + 1e-30)).item()
def case(B, H, N, sn):
    mv.set_option("attn_sn", sn)
    D = H * 64
    qkv = (torch.randn(B * N, 3 * D, device=dev) * 1.0).half()
    out, lse = mv.attention_fwd(qkv, B, H, N, q_out=(5, 10))
    torch.cuda.synchronize()
    x = qkv.double().reshape(B, N, 3, H, 64).permute(2, 0, 3, 1, 4)
    q, k, v = [t.clone().requires_grad_(True) for t in (x[0], x[1], x[2])]
    s = (q @ k.transpose(-2, -1)) * 0.125
    o = (s.softmax(-1) @ v)
    oref = o.transpose(1, 2).reshape(B * N, D)
    lse_ref = torch.logsumexp(s, -1) * 1.4426950408889634
    print("sn=%d B%d H%d N%d fwd: out rel err %.3e  lse abs err %.3e  nan %d" % (sn, B, H, N, rel(out, oref), (lse.double() - lse_ref).abs().max().item(), int(torch.isnan(out.float()).sum())), flush=True)
    if fwd_only: return
    do = (torch.randn(B * N, D, device=dev)).half()
    oref.backward(do.double())
    dqkv = mv.attention_bwd(qkv, out, do, lse, B, H, N)
    torch.cuda.synchronize()
    g = dqkv.double().reshape(B, N, 3, H, 64).permute(2, 0, 3, 1, 4)
    print("   bwd: dq rel %.3e  dk rel %.3e  dv rel %.3e  nan %d" % (rel(g[0], q.grad), rel(g[1], k.grad), rel(g[2], v.grad), int(torch.isnan(dqkv.float()).sum())), flush=True)
for args in [(1, 1, 128), (1, 1, 64), (2, 2, 257), (3, 2, 197), (1, 2, 272), (2, 1, 130), (2, 1, 131), (1, 1, 1), (1, 1, 17), (40, 6, 257), (2, 1, 256)]:
    try:
        case(*args, 1)
    except Exception as e:
        print("EXC", args, repr(e)); traceback.print_exc(); break
if quick: sys.exit(0)
def timeit(fn, n=10):
    for _ in range(3): fn()
    e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n
B, H, N = 256, 6, 257
D = H * 64
qkvs = [torch.randn(B * N, 3 * D, device=dev).half() for _ in range(3)]      # 3 x 151 MB > L2
out, lse = mv.attention_fwd(qkvs[0], B, H, N)
do = torch.randn(B * N, D, device=dev).half()
dqkv = torch.empty_like(qkvs[0]); delta = torch.empty(B, H, N, device=dev)
fl = 4.0 * B * H * N * N * 64
for sn in (1, 0):
    mv.set_option("attn_sn", sn)
    it = [0]
    def f():
        it[0] += 1
        mv.attention_fwd(qkvs[it[0] % 3], B, H, N, out=out, lse=lse, q_out=(5, 10))
    ms = timeit(f)
    print("sn=%d attn fwd B%d H%d N%d: %.3f ms  %.1f TFLOP/s (algorithmic)" % (sn, B, H, N, ms, fl / ms / 1e9), flush=True)
    if not fwd_only:
        def g():
            it[0] += 1
            mv.attention_bwd(qkvs[it[0] % 3], out, do, lse, B, H, N, dqkv=dqkv, delta=delta)
        ms = timeit(g)
        print("sn=%d attn bwd: %.3f ms  %.1f TFLOP/s (algorithmic 2.5x fwd)" % (sn, ms, 2.5 * fl / ms / 1e9), flush=True)
