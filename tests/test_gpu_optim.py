"""Fused multi-tensor AdamW + weight re-quantisation against torch.optim.AdamW and the
standalone weight-quant kernel."""
import copy

import pytest
import torch
import torch.nn.functional as F

from oracle.golden_cases import CASES, make_inputs

pytestmark = pytest.mark.gpu
dev = "cuda"


def test_matches_torch_adamw_on_plain_tensors():
    from myrtle_vision.utils.fused_adamw import FusedAdamW
    torch.manual_seed(0)
    shapes = [(1,), (7,), (1024,), (1025,), (33, 65), (384, 1536), (3, 5, 7)]
    ps = [torch.randn(s, device=dev).requires_grad_(True) for s in shapes]
    qs = [p.detach().clone().requires_grad_(True) for p in ps]
    kw = dict(lr=3e-3, betas=(0.9, 0.95), eps=1e-8)
    ours = FusedAdamW([{"params": ps[:3], "weight_decay": 0.0}, {"params": ps[3:], "weight_decay": 0.05}], **kw)
    ref = torch.optim.AdamW([{"params": qs[:3], "weight_decay": 0.0}, {"params": qs[3:], "weight_decay": 0.05}], **kw)
    for it in range(4):
        for p, q in zip(ps, qs):
            g = torch.randn_like(p) * (10.0 ** (it - 2))
            p.grad, q.grad = g.clone(), g.clone()
        if it == 2:
            for grp in ours.param_groups + ref.param_groups:
                grp["lr"] = 1e-3                       # lr schedulers write param_groups
        ours.step()
        ref.step()
        for p, q in zip(ps, qs):
            assert torch.allclose(p, q, rtol=2e-6, atol=1e-7), (it, p.shape)
    sd = ours.state_dict()
    assert sd["state"][0]["step"] == 4 and set(sd["state"][0]) == {"step", "exp_avg", "exp_avg_sq"}
    assert torch.allclose(sd["state"][5]["exp_avg_sq"], ref.state_dict()["state"][5]["exp_avg_sq"], rtol=1e-5)


def test_grad_scaler_semantics_on_device():
    from myrtle_vision.utils.fused_adamw import FusedAdamW
    p = torch.randn(3000, device=dev).requires_grad_(True)
    q = p.detach().clone().requires_grad_(True)
    ours, ref = FusedAdamW([p], lr=1e-2), torch.optim.AdamW([q], lr=1e-2)
    g = torch.randn_like(p)
    p.grad, q.grad = g * 65536.0, g.clone()
    before = p.detach().clone()
    ours.step(inv_scale=torch.tensor(1.0 / 65536.0, device=dev), found_inf=torch.tensor(1.0, device=dev))
    assert torch.equal(p, before)                      # overflow step skipped entirely
    ours.step(inv_scale=torch.tensor(1.0 / 65536.0, device=dev), found_inf=torch.tensor(0.0, device=dev))
    ref.step()
    assert torch.allclose(p, q, rtol=2e-6, atol=1e-7)  # and it counted as step 1, not 2


@pytest.mark.parametrize("fmt", ["FP16_32", "TF32"])
def test_model_step_emits_the_next_steps_operands(fmt):
    import mv_native as mv
    from myrtle_vision.models.vit import ViT
    from myrtle_vision.utils.fused_adamw import FusedAdamW
    from myrtle_vision.utils.optim import add_weight_decay
    case = CASES["classification"]
    torch.manual_seed(3)
    m = ViT(decoder="classification", image_size=case["image_size"], patch_size=16, num_classes=5,
            dim=128, depth=2, heads=2, mlp_dim=256, q_format=fmt).to(dev).train()
    twin = copy.deepcopy(m)
    img, tgt = make_inputs("classification", case, 11)
    img, tgt = img.to(dev), tgt.to(dev)
    ours = FusedAdamW(add_weight_decay(m, 0.05), lr=1e-3, model=m)
    ref = torch.optim.AdamW(add_weight_decay(twin, 0.05), lr=1e-3)
    for it in range(3):
        m.zero_grad(set_to_none=True)
        F.cross_entropy(m(img), tgt).backward()
        # identical gradients for both optimizers (Adam amplifies last-bit gradient noise to +-lr)
        for p, q in zip(m.parameters(), twin.parameters()):
            q.grad = None if p.grad is None else p.grad.clone()
        ours.step()
        ref.step()
        if it == 0:
            n0 = mv.launch_count()
            m.engine().quantised_weights()
            assert mv.launch_count() == n0             # operands came from the optimizer kernel
    for (n, p), q in zip(m.named_parameters(), twin.parameters()):
        assert torch.allclose(p, q, rtol=1e-4, atol=1e-6), n
    eng = m.engine()
    e, man = eng.fmt
    for i, (qw, qwt) in eng._wq.items():
        want, want_t = mv.quantize_weight(eng.params[i].detach(), e, man, out_dtype=qw.dtype)
        assert torch.equal(qw, want) and torch.equal(qwt, want_t)    # bit-exact operands
    with torch.no_grad():
        assert torch.allclose(m(img), twin(img), rtol=1e-3, atol=1e-3)


def test_fp32_eager_steps_track_torch_adamw():
    """q_format=FP32 (3xTF32 split operands): the fused optimizer must not write q(W) tiles into the split buffers,
    and the engine must re-split the weights it updated through raw pointers — eager mode, several steps."""
    from myrtle_vision.models.vit import ViT
    from myrtle_vision.utils.fused_adamw import FusedAdamW
    from myrtle_vision.utils.optim import add_weight_decay
    case = CASES["classification"]
    torch.manual_seed(5)
    m = ViT(decoder="classification", image_size=case["image_size"], patch_size=16, num_classes=5,
            dim=128, depth=2, heads=2, mlp_dim=256, q_format="FP32").to(dev).train()
    twin = copy.deepcopy(m)
    img, tgt = make_inputs("classification", case, 11)
    img, tgt = img.to(dev), tgt.to(dev)
    ours = FusedAdamW(add_weight_decay(m, 0.05), lr=2e-3, model=m)
    ref = torch.optim.AdamW(add_weight_decay(twin, 0.05), lr=2e-3)
    losses, losses_ref = [], []
    for it in range(4):
        m.zero_grad(set_to_none=True)
        twin.zero_grad(set_to_none=True)
        la, lb = F.cross_entropy(m(img), tgt), F.cross_entropy(twin(img), tgt)
        la.backward()
        lb.backward()
        losses.append(float(la)); losses_ref.append(float(lb))
        ours.step()
        ref.step()
    # both models run the same kernels; the only difference is the optimizer: the trajectories must coincide
    for a, b in zip(losses, losses_ref):
        assert abs(a - b) <= 2e-4 * abs(b), (losses, losses_ref)
    assert losses[-1] < losses[0]
    # parameters: Adam's first steps move every weight by ~lr whatever the gradient's size, so two runs whose
    # gradients differ in the last bits agree to a fraction of lr, not to fp32 rounding
    for (n, p), q in zip(m.named_parameters(), twin.parameters()):
        assert (p - q).abs().max() <= 2e-3 * 4 * 0.25 + 1e-6, n
    with torch.no_grad():
        assert torch.allclose(m(img), twin(img), rtol=2e-3, atol=2e-3)
