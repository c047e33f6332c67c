"""Generate tests/golden/convert_*.npz by running the UNMODIFIED reference's PTQ path here.

    python oracle/make_golden_convert.py

Flow of classification/test_quantize.py:37-134 of the reference: build the model in `q_format`, load seeded weights,
one calibration forward (a no-op for the QPyTorch formats), `vit.convert()` (utils/quantize.py:329-348 ->
QLinear.from_float / QLayerNorm.from_float, :134-166), evaluation forward.  Stored per q_format: the eval logits
before and after convert(), the state_dict keys after convert(), and the converted values of one Linear weight and
one LayerNorm gamma.  Test infrastructure; the GPU box only reads the committed fixture.
"""
import json
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "oracle", "shim"))
sys.path.insert(0, "/root/reference/src")

from myrtle_vision.models.vit import ViT  # noqa: E402  (the reference)
from oracle import vit_oracle  # noqa: E402
from oracle.golden_cases import ARCH, CASES, GOLD, make_inputs  # noqa: E402

PROBES = ("transformer.layers.0.0.fn.fn.to_qkv.1.weight", "transformer.layers.1.1.fn.norm.1.weight",
          "decoder.norm.1.weight", "decoder.linear.1.weight")


def main():
    decoder, case = "classification", CASES["classification"]
    meta = {}
    for fmt in ("FP16_32", "FP16_16", "TF32"):
        model = ViT(decoder=decoder, image_size=case["image_size"], patch_size=16, num_classes=case["num_classes"],
                    dim=ARCH["dim"], depth=ARCH["depth"], heads=ARCH["heads"], mlp_dim=ARCH["mlp_dim"], q_format=fmt)
        P = vit_oracle.init_params(decoder=decoder, num_classes=case["num_classes"], dim=ARCH["dim"],
                                   depth=ARCH["depth"], heads=ARCH["heads"], mlp_dim=ARCH["mlp_dim"], seed=4321)
        # LayerNorm gammas away from 1 so that their quantisation (convert() only) shows in the output
        g = torch.Generator().manual_seed(99)
        for k in P:
            if k.endswith("norm.weight"):
                P[k] = P[k] + 0.3 * torch.randn(P[k].shape, generator=g)
        model.load_state_dict({k: P[vit_oracle.canonical_key(k)] for k in model.state_dict()})
        img, _ = make_inputs(decoder, case, 77)
        model.eval()
        with torch.no_grad():
            before = model(img)
            model.convert()
            after = model(img)
        sd = model.state_dict()
        out = {"before": before.numpy(), "after": after.numpy()}
        for k in PROBES:
            out["sd/" + k] = sd[k].numpy()
        np.savez_compressed(os.path.join(GOLD, "convert_%s.npz" % fmt), **out)
        meta[fmt] = {"keys": list(sd.keys()), "max_change": float((after - before).abs().max()),
                     "modules": sorted({type(m).__name__ for m in model.modules()})}
        print(fmt, "max |after - before| = %.3e" % meta[fmt]["max_change"], meta[fmt]["modules"])
    with open(os.path.join(GOLD, "convert_meta.json"), "w") as f:
        json.dump(meta, f)


if __name__ == "__main__":
    main()
