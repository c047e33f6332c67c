"""Timeline of the short-sequence attention backward (debug): CTA 0 stamps clock64() at every hand-off
between the MMA issuer and the two math groups (mv_debug_set_attn_trace).  Prints per-event gaps in cycles."""
import ctypes, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "myrtle-vision_b200")); sys.path.insert(0, ROOT)
import torch
import mv_native as mv
if os.environ.get("MV_ALT_LIB"):        # a variant built by tools/build_variant.sh
    mv._SO = os.path.join(ROOT, "myrtle-vision_b200", "csrc", os.environ["MV_ALT_LIB"])
if os.environ.get("MV_TRACE_LIB"):      # tools/build_variant.sh attention_sn.cu trace -DMV_SN_TRACE
    mv._SO = os.path.join(ROOT, "myrtle-vision_b200", "csrc", "libmv_b200_trace.so")
B, H, N = 256, 6, 257
D = H * 64
torch.manual_seed(0)
qkv = torch.randn(B * N, 3 * D, device="cuda").half()
do = torch.randn(B * N, D, device="cuda").half()
out, lse = mv.attention_fwd(qkv, B, H, N, q_out=(5, 10))
dbias = torch.zeros(3 * D, device="cuda") if os.environ.get("MV_TRACE_DBIAS") else None     # the fused to_qkv bias gradient, as in the step
for _ in range(2):
    mv.attention_bwd(qkv, out, do, lse, B, H, N, dbias=dbias)
buf = torch.zeros(3 * 3072, dtype=torch.int64, device="cuda")
mv.lib().mv_debug_set_attn_trace(ctypes.c_void_p(buf.data_ptr()))
(mv.attention_fwd(qkv, B, H, N, q_out=(5, 10)) if "fwd" in sys.argv else mv.attention_bwd(qkv, out, do, lse, B, H, N, dbias=dbias))
torch.cuda.synchronize()
mv.lib().mv_debug_set_attn_trace(None)
t = buf.cpu().view(3, 1024, 3)
names = {1: "pre issued", 2: "issuer waits math", 3: "issuer sees math done", 4: "post issued", 5: "math waits S", 6: "S ready",
         7: "math done", 12: "max pass done (fwd)", 16: "dq stored", 17: "bar after dq", 18: "pair finished", 19: "stats start", 12: "pre enter", 13: "pre fenced", 14: "pre 2 MMAs issued", 15: "pre 8 MMAs issued", 8: "waits acc", 9: "acc ready", 10: "dq ready", 11: "stats done", 24: "odd keys done", 25: "odd: k,v copied", 26: "odd: A done", 27: "odd: B done", 20: "epi acc in regs", 21: "epi staging free", 22: "epi parked", 23: "blk staging free"}
ev = []
for r in range(3):
    for e, i, c in t[r].tolist():
        if e:
            ev.append((c, r, e, i))
ev.sort()
t0 = ev[0][0]
# first two (image, head) pairs of CTA 0
for c, r, e, i in ev[:260]:
    print("%8d  %-7s %-22s %d" % (c - t0, ["issuer", "math g0", "math g1"][r], names[e], i))
# per-pair summary over everything recorded (math g0): when each pair finished, and the kernel's clock rate
fin = [c - t0 for c, r, e, i in ev if r == 1 and e == 18]
print("pairs finished at:", fin)
print("cycles per pair:", [b - a for a, b in zip(fin, fin[1:])])
