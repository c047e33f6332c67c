"""Shared definition of the golden cases: architectures, seeded inputs and the gradient
digest.  TEST INFRASTRUCTURE ONLY.  Used by oracle/make_golden.py (which runs the reference,
in the build container only) and by tests/ (which only read the committed fixtures)."""
import os

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GOLD = os.path.join(ROOT, "tests", "golden")
CASES = {
    "classification": dict(image_size=80, num_classes=5, batch=3),
    "segmentation": dict(image_size=80, num_classes=4, batch=2),
    "detection": dict(image_size=176, num_classes=3, batch=2),   # 121 patches >= 100 det slots
}
ARCH = dict(dim=128, depth=2, heads=2, mlp_dim=256, patch_size=16)
FORMATS = ["FP32", "FP16_32", "TF32", "FP16_16"]


def digest(t):
    f = t.detach().flatten().double()
    idx = torch.linspace(0, f.numel() - 1, steps=min(24, f.numel())).long()
    return {"sum": float(f.sum()), "abs": float(f.abs().sum()), "idx": idx.tolist(),
            "val": [float(v) for v in t.detach().flatten()[idx]]}


def make_inputs(decoder, case, seed):
    g = torch.Generator().manual_seed(seed)
    b, s = case["batch"], case["image_size"]
    img = torch.randn(b, 3, s, s, generator=g).clamp(-1, 1)
    if decoder == "classification":
        tgt = torch.randint(0, case["num_classes"], (b,), generator=g)
    elif decoder == "segmentation":
        tgt = torch.randint(0, case["num_classes"], (b, s, s), generator=g)
    else:
        tgt = {"labels": torch.randint(0, case["num_classes"] + 1, (b, 100), generator=g),
               "boxes": torch.rand(b, 100, 4, generator=g)}
    return img, tgt


