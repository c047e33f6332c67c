"""Bring-up probe: full model forward/backward on the GPU vs the golden fixtures (reference run)
and vs the CPU oracle run live on the same inputs."""
import json, os, sys, time, traceback
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "myrtle-vision_b200")); sys.path.insert(0, ROOT)
import numpy as np, torch
import torch.nn.functional as F
from myrtle_vision.models.vit import ViT
from oracle import vit_oracle
from oracle.golden_cases import ARCH, CASES, GOLD, make_inputs
dev = "cuda"

def build(decoder, case, fmt, seed):
    m = ViT(decoder=decoder, image_size=case["image_size"], patch_size=16, num_classes=case["num_classes"],
            dim=ARCH["dim"], depth=ARCH["depth"], heads=ARCH["heads"], mlp_dim=ARCH["mlp_dim"], q_format=fmt)
    P = vit_oracle.init_params(decoder=decoder, num_classes=case["num_classes"], dim=ARCH["dim"], depth=ARCH["depth"],
                               heads=ARCH["heads"], mlp_dim=ARCH["mlp_dim"], seed=seed)
    m.load_state_dict({k: P[vit_oracle.canonical_key(k)] for k in m.state_dict()})
    return m.to(dev), P

def loss_fn(decoder, out, tgt):
    if decoder == "detection":
        return (F.cross_entropy(out["pred_logits"].flatten(0, 1), tgt["labels"].flatten())
                + (out["pred_boxes"] - tgt["boxes"]).abs().mean())
    return F.cross_entropy(out, tgt)

def todev(t):
    return {k: v.to(dev) for k, v in t.items()} if isinstance(t, dict) else t.to(dev)

for decoder in ("classification", "segmentation", "detection"):
    for fmt in ("FP16_32", "FP16_16"):
        try:
            case = CASES[decoder]
            meta = json.load(open(os.path.join(GOLD, "vit_%s_%s.json" % (decoder, fmt))))
            m, P = build(decoder, case, fmt, meta["seed"])
            img, tgt = make_inputs(decoder, case, meta["seed"] + 1)
            m.train(); m.zero_grad()
            out = m(img.to(dev)); loss = loss_fn(decoder, out, todev(tgt)); loss.backward()
            torch.cuda.synchronize()
            oo, ol, og = vit_oracle.train_step(P, img, tgt, decoder=decoder, heads=ARCH["heads"], q_format=fmt)
            print("%s %s: loss gpu %.6f  golden %.6f  oracle %.6f" % (decoder, fmt, loss.item(), meta["loss"], float(ol)))
            if decoder == "detection":
                a, b = out["pred_logits"].detach().cpu(), oo["pred_logits"].detach()
            else:
                a, b = out.detach().cpu(), oo.detach()
            print("   out max abs err %.3e (max |out| %.3f)" % ((a - b).abs().max().item(), b.abs().max().item()))
            worst = []
            for name, prm in m.named_parameters():
                ck = vit_oracle.canonical_key(name)
                if og[ck] is None:
                    assert prm.grad is None, name
                    continue
                gg, rr = prm.grad.detach().cpu().double(), og[ck].double()
                rel = ((gg - rr).norm() / (rr.norm() + 1e-30)).item()
                worst.append((rel, ck))
            worst.sort(reverse=True)
            print("   grad rel-L2 err: worst", ["%s %.2e" % (k, r) for r, k in worst[:4]], "median %.2e" % worst[len(worst) // 2][0])
        except Exception as e:
            print("EXC", decoder, fmt, repr(e)); traceback.print_exc()

# flagship timing: ViT-Small 256x256 batch 256, FP16_32, fwd+bwd
def timeit(fn, n=5):
    for _ in range(2): fn()
    torch.cuda.synchronize()
    e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n
try:
    torch.manual_seed(1234)
    m = ViT(decoder="classification", image_size=256, patch_size=16, num_classes=45, dim=384, depth=12, heads=6, mlp_dim=1536, q_format="FP16_32").to(dev)
    B = 256
    img = torch.randn(B, 3, 256, 256, device=dev).clamp(-1, 1); y = torch.randint(0, 45, (B,), device=dev)
    def step():
        m.zero_grad(set_to_none=True)
        F.cross_entropy(m(img), y).backward()
    ms = timeit(step)
    print("ViT-Small 256^2 B=256 FP16_32 fwd+bwd: %.2f ms/step  %.0f img/s  launches/step %d" % (ms, B / ms * 1e3, 0))
    import mv_native
    c0 = mv_native.launch_count(); step(); print("   kernels per step:", mv_native.launch_count() - c0)
    with torch.no_grad():
        ms = timeit(lambda: m(img))
    print("   forward only: %.2f ms" % ms)
    from torch.profiler import profile, ProfilerActivity
    with profile(activities=[ProfilerActivity.CUDA]) as prof:
        step(); torch.cuda.synchronize()
    print(prof.key_averages().table(sort_by="cuda_time_total", row_limit=25, max_name_column_width=60))
except Exception as e:
    print("EXC flagship", repr(e)); traceback.print_exc()
