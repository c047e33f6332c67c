"""Bring-up probe (not a test): CTA-pair GEMM kernel (cluster=0) against the single-CTA kernel
(cluster=1) — correctness over layouts / epilogues, then cold-cache timings of the ViT-Small step
shapes.  Prints statistics instead of asserting.  Usage: python tools/probe_gemm2.py [quick]"""
import os, sys, traceback
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "myrtle-vision_b200")); sys.path.insert(0, ROOT)
import torch
import torch.nn.functional as F
import mv_native as mv

torch.manual_seed(0)
dev = "cuda"
h, f32 = torch.float16, torch.float32
quick = "quick" in sys.argv


def rel(got, want):
    return ((got.double() - want.double()).abs().max() / (want.double().abs().max() + 1e-30)).item()


def run(name, fn):
    try:
        fn(); torch.cuda.synchronize()
    except Exception as e:
        print(name, "EXC", repr(e)); traceback.print_exc()


def case(M, N, K, a_major, b_major, **kw):
    def f():
        A = torch.randn((K, M) if a_major else (M, K), device=dev).to(h)
        B = torch.randn((K, N) if b_major else (N, K), device=dev).to(h)
        out = torch.full((M, N), float("nan"), device=dev)
        mv.gemm(A, B, out, a_major=a_major, b_major=b_major, **kw)
        torch.cuda.synchronize()
        Af = A.double().t() if a_major else A.double()
        Bf = B.double().t() if b_major else B.double()
        r = rel(out, Af @ Bf.t())
        print("M%-6d N%-5d K%-5d amaj%d bmaj%d %-22s rel %.2e %s" % (M, N, K, a_major, b_major, kw, r, "OK" if r < 3e-5 else "BAD"), flush=True)
    run("case", f)


for kw in ({}, {"tile_n": 128}, {"tile_n": 192}, {"tile_n": 256}):
    case(128, 128, 64, 0, 0, **kw)
    case(256, 256, 384, 0, 0, **kw)
    case(1000, 384, 1536, 0, 0, **kw)
    case(200, 45, 384, 0, 0, **kw)
    case(520, 1536, 384, 0, 0, **kw)
    if kw.get("tile_n") != 192:
        case(256, 384, 512, 1, 1, **kw)
        case(384, 256, 8192, 1, 1, **kw)
        case(520, 1536, 384, 0, 1, **kw)
    case(520, 1536, 384, 1, 0, **kw)
case(65792, 1152, 384, 0, 0)
case(5002, 1152, 384, 0, 0)


def epi():
    M, N, K = 514, 384, 384
    A = torch.randn(M, K, device=dev).half(); B = (torch.randn(N, K, device=dev) * 0.05).half()
    bias = torch.randn(N, device=dev); res = torch.randn(M, N, device=dev)
    lin = A.double() @ B.double().t() + bias.double()
    out = torch.empty(M, N, device=dev)
    mv.gemm(A, B, out, bias=bias, residual=res)
    print("epi bias+res            rel %.2e" % rel(out, lin + res.double()))
    outq = torch.empty(M, N, device=dev, dtype=h)
    mv.gemm(A, B, outq, bias=bias, q_out=(5, 10))
    want = mv.float_quantize(lin.float(), 5, 10)
    print("epi q_out               rel %.2e  flips %.4f" % (rel(outq, want), ((outq.float() - want).abs() > 0).float().mean().item()))
    gp = torch.empty(M, N, device=dev, dtype=h); hh = torch.empty_like(gp)
    mv.gemm(A, B, hh, bias=bias, aux=gp, epilogue=mv.EPI_GELU, q_res=(5, 10))
    uu = lin.clone().requires_grad_(True)
    F.gelu(uu).sum().backward()
    print("epi gelu                rel %.2e  gelu' %.2e" % (rel(hh, F.gelu(lin)), rel(gp, uu.grad)))
    d = torch.empty(M, N, device=dev, dtype=h)
    mv.gemm(A, B, d, aux=gp, epilogue=mv.EPI_DGELU)
    print("epi dgelu               rel %.2e" % rel(d, (A.double() @ B.double().t()) * uu.grad))
    pos = torch.randn(257, N, device=dev)
    A2 = torch.randn(2 * 257, K, device=dev).half()
    o2 = torch.empty(2 * 257, N, device=dev)
    mv.gemm(A2, B, o2, bias=bias, residual=pos, rows_per_img=257)
    print("epi pos broadcast       rel %.2e" % rel(o2, A2.double() @ B.double().t() + bias.double() + pos.double().repeat(2, 1)))
    T, No, Ki = 8192, 384, 1536
    dY = torch.randn(T, No, device=dev).half(); X = torch.randn(T, Ki, device=dev).half()
    o3 = torch.zeros(No, Ki, device=dev)
    mv.gemm(dY, X, o3, a_major=1, b_major=1, accumulate=True)
    print("wgrad split-K           rel %.2e" % rel(o3, dY.double().t() @ X.double()))
    o4 = torch.zeros(1152, 384, device=dev)
    dY = torch.randn(T, 1152, device=dev).half(); X = torch.randn(T, 384, device=dev).half()
    mv.gemm(dY, X, o4, a_major=1, b_major=1, accumulate=True)
    print("wgrad split-K 1152x384  rel %.2e" % rel(o4, dY.double().t() @ X.double()))
run("epi", epi)

if quick:
    sys.exit(0)

M = 65792


def timeit(call, nbuf, reps=3):
    for i in range(nbuf): call(i)
    e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(reps * nbuf): call(i % nbuf)
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / (reps * nbuf)


def fwd_cold(name, N, K, nbuf=6, gelu=False, res=False, dgelu=False, odt=h):
    As = [torch.randn(M, K, device=dev).to(h) for _ in range(nbuf)]
    Bm = (torch.randn(N, K, device=dev) / K ** 0.5).to(h)      # outputs ~ N(0,1) like the model's pre-activations
    outs = [torch.empty(M, N, device=dev, dtype=odt) for _ in range(nbuf)]
    auxs = [torch.randn(M, N, device=dev).to(h) for _ in range(nbuf)] if (gelu or dgelu) else None
    ress = [torch.randn(M, N, device=dev) for _ in range(nbuf)] if res else None
    bias = torch.randn(N, device=dev)
    for cl in (0, 1):
        def call(j):
            if gelu: mv.gemm(As[j], Bm, outs[j], bias=bias, aux=auxs[j], epilogue=mv.EPI_GELU, q_res=(5, 10), cluster=cl)
            elif dgelu: mv.gemm(As[j], Bm, outs[j], aux=auxs[j], epilogue=mv.EPI_DGELU, cluster=cl)
            elif res: mv.gemm(As[j], Bm, outs[j], bias=bias, residual=ress[j], cluster=cl)
            else: mv.gemm(As[j], Bm, outs[j], bias=bias, cluster=cl)
        ms = timeit(call, nbuf)
        byts = M * K * 2 + M * N * outs[0].element_size() * (2 if gelu else 1) + (M * N * 2 if dgelu else 0) + (M * N * 4 if res else 0)
        print("cold %-26s N%-5d K%-5d cluster=%d: %.3f ms  %6.1f TFLOP/s  %6.1f GB/s" % (name, N, K, cl, ms, 2.0 * M * N * K / ms / 1e9, byts / ms / 1e6), flush=True)


def wgrad_cold(No, Ki, nbuf=4):
    dYs = [torch.randn(M, No, device=dev).to(h) for _ in range(nbuf)]
    Xs = [torch.randn(M, Ki, device=dev).to(h) for _ in range(nbuf)]
    out = torch.zeros(No, Ki, device=dev)
    for cl in (0, 1):
        ms = timeit(lambda j: mv.gemm(dYs[j], Xs[j], out, a_major=1, b_major=1, accumulate=True, cluster=cl), nbuf)
        print("cold wgrad %4dx%-4d cluster=%d: %.3f ms %6.1f TFLOP/s" % (No, Ki, cl, ms, 2.0 * M * No * Ki / ms / 1e9), flush=True)


run("c", lambda: fwd_cold("qkv fwd (f16 out)", 1152, 384))
run("c", lambda: fwd_cold("proj fwd (+res, f32 out)", 384, 384, res=True, odt=f32))
run("c", lambda: fwd_cold("fc1 fwd (gelu)", 1536, 384, gelu=True))
run("c", lambda: fwd_cold("fc2 fwd (+res, f32 out)", 384, 1536, res=True, odt=f32))
run("c", lambda: fwd_cold("dgrad fc2 (dgelu)", 1536, 384, dgelu=True))
run("c", lambda: fwd_cold("dgrad fc1", 384, 1536))
run("c", lambda: fwd_cold("dgrad proj", 384, 384))
run("c", lambda: fwd_cold("dgrad qkv", 384, 1152))
run("c", lambda: fwd_cold("patch embed (+pos)", 384, 768, res=True, odt=f32))
run("w", lambda: wgrad_cold(1536, 384))
run("w", lambda: wgrad_cold(384, 1536))
run("w", lambda: wgrad_cold(1152, 384))
run("w", lambda: wgrad_cold(384, 384))
