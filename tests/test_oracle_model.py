"""oracle/vit_oracle.py (functional restatement) against fixtures produced by running the
unmodified reference model in the build container (oracle/make_golden.py)."""
import json
import os

import numpy as np
import pytest
import torch

from oracle import vit_oracle
from oracle.golden_cases import ARCH, CASES, FORMATS, GOLD, make_inputs


def load(decoder, fmt):
    stem = os.path.join(GOLD, "vit_%s_%s" % (decoder, fmt))
    with open(stem + ".json") as f:
        meta = json.load(f)
    return meta, np.load(stem + ".npz")


def check_digest(g, d, rtol, atol_scale):
    f = g.detach().flatten().double()
    scale = d["abs"] / max(f.numel(), 1) + 1e-30
    assert abs(float(f.sum()) - d["sum"]) <= rtol * d["abs"] + 1e-12
    assert abs(float(f.abs().sum()) - d["abs"]) <= rtol * d["abs"] + 1e-12
    got = f[torch.tensor(d["idx"])]
    want = torch.tensor(d["val"], dtype=torch.float64)
    assert torch.all((got - want).abs() <= rtol * want.abs() + atol_scale * scale)


@pytest.mark.parametrize("decoder", list(CASES))
@pytest.mark.parametrize("fmt", FORMATS)
def test_oracle_matches_reference_fixture(decoder, fmt):
    meta, arrs = load(decoder, fmt)
    case = meta["case"]
    P = vit_oracle.init_params(decoder=decoder, num_classes=case["num_classes"], dim=ARCH["dim"],
                               depth=ARCH["depth"], heads=ARCH["heads"], mlp_dim=ARCH["mlp_dim"],
                               seed=meta["seed"])
    img, tgt = make_inputs(decoder, case, meta["seed"] + 1)
    out, loss, grads = vit_oracle.train_step(P, img, tgt, decoder=decoder, heads=ARCH["heads"],
                                             q_format=fmt)
    # same ops on the same CPU: identical up to op-ordering noise
    assert abs(float(loss) - meta["loss"]) <= 2e-6 * abs(meta["loss"])
    if decoder == "detection":
        np.testing.assert_allclose(out["pred_logits"].detach().numpy(), arrs["pred_logits"],
                                   rtol=1e-5, atol=1e-6)
        np.testing.assert_allclose(out["pred_boxes"].detach().numpy(), arrs["pred_boxes"],
                                   rtol=1e-5, atol=1e-6)
    else:
        o = out.detach()
        if decoder == "segmentation":
            o = o[:1, :, ::5, ::5]
        np.testing.assert_allclose(o.numpy(), arrs["out"], rtol=1e-5, atol=1e-6)
    for k, d in meta["grads"].items():
        if d is None:
            assert grads[k] is None, k          # det tokens get no gradient
        else:
            check_digest(grads[k], d, rtol=1e-4, atol_scale=1e-3)


def test_canonical_key():
    assert vit_oracle.canonical_key("patch_to_embedding.1.weight") == "patch_to_embedding.weight"
    assert (vit_oracle.canonical_key("transformer.layers.3.0.fn.fn.to_out.0.1.bias")
            == "transformer.layers.3.0.fn.fn.to_out.0.bias")
    assert (vit_oracle.canonical_key("transformer.layers.3.1.fn.fn.net.0.weight")
            == "transformer.layers.3.1.fn.fn.net.0.weight")
    assert vit_oracle.canonical_key("pos_embedding") == "pos_embedding"


def _convert_case(seed=4321):
    case = CASES["classification"]
    P = vit_oracle.init_params(decoder="classification", num_classes=case["num_classes"], dim=ARCH["dim"],
                               depth=ARCH["depth"], heads=ARCH["heads"], mlp_dim=ARCH["mlp_dim"], seed=seed)
    g = torch.Generator().manual_seed(99)          # gammas away from 1: their quantisation happens in convert() only
    for k in P:
        if k.endswith("norm.weight"):
            P[k] = P[k] + 0.3 * torch.randn(P[k].shape, generator=g)
    img, _ = make_inputs("classification", case, 77)
    return case, P, img


@pytest.mark.parametrize("fmt", ["FP16_32", "FP16_16", "TF32"])
def test_oracle_convert_matches_the_reference_convert(fmt):
    """PTQ path: fixtures from the unmodified reference's vit.convert() + eval forward
    (oracle/make_golden_convert.py).  Pins WHICH quantisers survive convert(): none of the hook-based ones."""
    z = np.load(os.path.join(GOLD, "convert_%s.npz" % fmt))
    case, P, img = _convert_case()
    with torch.no_grad():
        before = vit_oracle.vit_forward(P, img, decoder="classification", heads=ARCH["heads"], q_format=fmt)
        Q = vit_oracle.convert_params(P, fmt)
        after = vit_oracle.vit_forward(Q, img, decoder="classification", heads=ARCH["heads"], q_format=fmt,
                                       converted=True)
        kept = vit_oracle.vit_forward(Q, img, decoder="classification", heads=ARCH["heads"], q_format=fmt)
    np.testing.assert_allclose(before.numpy(), z["before"], rtol=1e-5, atol=2e-6)
    np.testing.assert_allclose(after.numpy(), z["after"], rtol=1e-5, atol=2e-5)
    # the alternative reading (activation quantisers kept) is measurably not what the reference does
    assert np.abs(kept.numpy() - z["after"]).max() > 10 * np.abs(after.numpy() - z["after"]).max()
    for k in z.files:
        if k.startswith("sd/"):
            assert np.array_equal(Q[vit_oracle.canonical_key(k[3:])].numpy(), z[k]), k
