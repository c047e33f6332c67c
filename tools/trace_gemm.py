"""Timeline of the CTA-pair GEMM (debug; build with MV_NVCC_FLAGS=-DMV_GEMM_TRACE): CTA 0 stamps clock64() at
the hand-offs between the MMA issuer and two of its epilogue warps.  usage: trace_gemm.py N K mode
(mode: plain | res | gelu | dgelu | wgrad)"""
import ctypes, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "myrtle-vision_b200")); sys.path.insert(0, ROOT)
import torch
import mv_native as mv
N, K, mode = int(sys.argv[1]), int(sys.argv[2]), sys.argv[3]
M, dev, h = 65792, "cuda", torch.float16
torch.manual_seed(0)
A = (torch.randn(M, K, device=dev) * (1.0 if mode == "gelu" else 0.05)).to(h)
B = (torch.randn(N, K, device=dev) * (0.4 / K ** 0.5 if mode == "gelu" else 1.0)).to(h)      # gelu: u with the benchmark's spread
out = torch.empty(M, N, device=dev, dtype=torch.float32 if mode == "res" else h)
aux = (torch.rand(M, N, device=dev)).to(h); res = torch.randn(M, N, device=dev) if mode == "res" else None
bias = torch.randn(N, device=dev) * (0.02 if mode == "gelu" else 1.0)
def run():
    if mode == "gelu": mv.gemm(A, B, out, bias=bias, aux=aux, epilogue=mv.EPI_GELU, q_res=(5, 10))
    elif mode == "dgelu": mv.gemm(A, B, out, aux=aux, epilogue=mv.EPI_DGELU)
    elif mode == "res": mv.gemm(A, B, out, bias=bias, residual=res)
    else: mv.gemm(A, B, out, bias=bias)
for _ in range(3): run()
buf = torch.zeros(3 * 3072, dtype=torch.int64, device=dev)
mv.lib().mv_debug_set_gemm_trace(ctypes.c_void_p(buf.data_ptr()))
run(); torch.cuda.synchronize()
mv.lib().mv_debug_set_gemm_trace(None)
names = {1: "issuer waits accumulator", 2: "accumulator free", 3: "first stage landed", 4: "tile MMAs issued",
         5: "epi waits tile", 6: "tile complete", 7: "epi chunks done", 20: "chunk: loads issued", 21: "chunk: staged", 22: "chunk: rows stored"}
ev = sorted((c, r, e, i) for r in range(3) for e, i, c in buf.cpu().view(3, 1024, 3)[r].tolist() if e)
t0 = ev[0][0]
for c, r, e, i in ev[:150]:
    print("%8d  %-9s %-26s %d" % (c - t0, ["issuer", "epi w2", "epi w17"][r], names[e], i))
