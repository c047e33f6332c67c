"""Measured parity of every q_format x decoder against the live CPU oracle (bring-up tool; prints the
numbers DESIGN.md quotes).  Run on the GPU box: python tools/probe_formats.py"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "myrtle-vision_b200"), os.path.join(ROOT, "tests")]
import torch  # noqa: E402
from oracle import vit_oracle  # noqa: E402
from oracle.golden_cases import ARCH, CASES, make_inputs  # noqa: E402
import test_gpu_model as T  # noqa: E402

for fmt in ("FP32", "TF32", "FP16_32", "FP16_16"):
    for decoder, case in CASES.items():
        m, P = T.build(decoder, case, fmt, 1234)
        img, tgt = make_inputs(decoder, case, 1235)
        out = m(img.cuda())
        loss = T.loss_fn(decoder, out, T.todev(tgt))
        loss.backward()
        oo, ol, og = vit_oracle.train_step(P, img, tgt, decoder=decoder, heads=ARCH["heads"], q_format=fmt)
        grads = {vit_oracle.canonical_key(n): p.grad for n, p in m.named_parameters()}
        worst = max(float((grads[k].cpu().double() - r.double()).norm() / (r.double().norm() + 1e-30))
                    for k, r in og.items() if r is not None)
        if decoder == "detection":
            oerr = max(float((out[k].detach().cpu() - oo[k].detach()).abs().max() / oo[k].detach().abs().max())
                       for k in ("pred_logits", "pred_boxes"))
        else:
            oerr = float((out.detach().cpu() - oo.detach()).abs().max() / oo.detach().abs().max())
        print("%-8s %-14s loss rel %.2e  out max-rel %.2e  worst grad rel-L2 %.2e"
              % (fmt, decoder, abs(loss.item() - float(ol)) / abs(float(ol)), oerr, worst), flush=True)
