"""Generate tests/golden/* by running the UNMODIFIED reference in this container.

    python oracle/make_golden.py

Imports /root/reference/src/myrtle_vision (through oracle/shim/qtorch, because
qtorch==0.3.0 itself is not installable offline), overwrites its randomly
initialised parameters with oracle.vit_oracle.init_params(seed) so the fixture
only has to carry a seed, runs one forward + loss + backward per
(q_format, decoder) on a seeded input and stores logits, loss and a digest of
every parameter gradient.  Also records the reference's state_dict key list per
q_format.  /root/reference does not exist on the GPU box, so
tests only ever read the committed fixtures.
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
import torch.nn.functional as F  # noqa: E402
from oracle import vit_oracle  # noqa: E402

from oracle.golden_cases import ARCH, CASES, FORMATS, GOLD, digest, make_inputs  # noqa: E402


def main():
    os.makedirs(GOLD, exist_ok=True)
    keys = {}
    for decoder, case in CASES.items():
        for fmt in FORMATS:
            seed = 1234
            model = ViT(decoder=decoder, image_size=case["image_size"], patch_size=16,
                        num_classes=case["num_classes"], dim=ARCH["dim"], depth=ARCH["depth"],
                        heads=ARCH["heads"], mlp_dim=ARCH["mlp_dim"], q_format=fmt)
            keys["%s/%s" % (decoder, fmt)] = list(model.state_dict().keys())
            P = vit_oracle.init_params(decoder=decoder, num_classes=case["num_classes"],
                                       dim=ARCH["dim"], depth=ARCH["depth"], heads=ARCH["heads"],
                                       mlp_dim=ARCH["mlp_dim"], seed=seed)
            sd = model.state_dict()
            new_sd = {k: P[vit_oracle.canonical_key(k)] for k in sd}
            assert set(vit_oracle.canonical_key(k) for k in sd) == set(P), "key mismatch"
            model.load_state_dict(new_sd)
            img, tgt = make_inputs(decoder, case, seed + 1)
            model.train()
            model.zero_grad()
            out = model(img)
            if decoder == "detection":
                loss = (F.cross_entropy(out["pred_logits"].flatten(0, 1), tgt["labels"].flatten())
                        + (out["pred_boxes"] - tgt["boxes"]).abs().mean())
                outs = {"pred_logits": out["pred_logits"].detach().numpy(),
                        "pred_boxes": out["pred_boxes"].detach().numpy()}
            else:
                loss = F.cross_entropy(out, tgt)
                o = out.detach()
                if decoder == "segmentation":      # keep the fixture small: 1 image, strided
                    o = o[:1, :, ::5, ::5]
                outs = {"out": o.numpy()}
            loss.backward()
            grads = {}
            for name, prm in model.named_parameters():
                ck = vit_oracle.canonical_key(name)
                grads[ck] = None if prm.grad is None else digest(prm.grad)
            meta = {"decoder": decoder, "q_format": fmt, "seed": seed, "case": case, "arch": ARCH,
                    "loss": float(loss), "grads": grads}
            stem = os.path.join(GOLD, "vit_%s_%s" % (decoder, fmt))
            np.savez_compressed(stem + ".npz", **outs)
            with open(stem + ".json", "w") as f:
                json.dump(meta, f)
            print(decoder, fmt, "loss", float(loss))
    with open(os.path.join(GOLD, "state_dict_keys.json"), "w") as f:
        json.dump(keys, f, indent=0)


if __name__ == "__main__":
    main()
