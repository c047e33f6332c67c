import os, sys
sys.path.insert(0, "/root/repo/myrtle-vision_b200")
import torch, mv_native as mv
dev="cuda"
def timeit(fn, n=10):
    for _ in range(3): fn()
    e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n
B,H=256,6
for N in (256, 257, 128, 129, 258):
    D=H*64
    qkvs=[torch.randn(B*N,3*D,device=dev).half() for _ in range(3)]
    out,lse=mv.attention_fwd(qkvs[0],B,H,N)
    it=[0]
    def f():
        it[0]+=1; mv.attention_fwd(qkvs[it[0]%3],B,H,N,out=out,lse=lse,q_out=(5,10))
    print(N, "%.3f ms"%timeit(f), flush=True)
