"""Host-side mirror of the reference interface: constructors, quantisation config, state_dict
keys, batch-size solver, config schema."""
import json
import os

import pytest
import torch

from oracle.golden_cases import GOLD


def make(decoder="classification", fmt="FP16_32", **kw):
    from myrtle_vision.models.vit import ViT
    args = dict(decoder=decoder, image_size=176 if decoder == "detection" else 80, patch_size=16,
                num_classes=5, dim=128, depth=2, heads=2, mlp_dim=256, q_format=fmt)
    args.update(kw)
    return ViT(**args)


@pytest.mark.parametrize("decoder", ["classification", "segmentation", "detection"])
@pytest.mark.parametrize("fmt", ["FP32", "FP16_32", "TF32", "FP16_16"])
def test_state_dict_keys_equal_the_reference(decoder, fmt):
    with open(os.path.join(GOLD, "state_dict_keys.json")) as f:
        keys = json.load(f)
    assert list(make(decoder, fmt).state_dict().keys()) == keys["%s/%s" % (decoder, fmt)]


def test_qformat_enum_and_errors():
    from myrtle_vision.utils.quantize import ModelQuantizer, QFormat
    assert [f.name for f in QFormat] == ["FP32", "PyTorchINT8", "FP16_16", "FP16_32", "TF32"]
    assert int(QFormat.FP16_32) == 3
    m = make(fmt="FP16_32")
    with pytest.raises(ValueError, match="model already quantized"):
        m.quantizer.prepare_qat("FP16_16")
    m = make(fmt=None)                      # FP32 may be re-prepared, like the reference
    m.quantizer.prepare_qat("FP16_32")
    assert "patch_to_embedding.1.weight" in m.state_dict()
    with pytest.raises(KeyError):
        make(fmt="FP8")
    with pytest.raises(NotImplementedError):
        make(fmt="PyTorchINT8")


def test_constructor_asserts():
    with pytest.raises(AssertionError, match="divisible by the patch size"):
        make(image_size=81)
    with pytest.raises(AssertionError, match="way too small"):
        make(image_size=64)
    with pytest.raises(AssertionError, match="decoder must be"):
        make(decoder="pose")


def test_cpu_forward_fails_loudly():
    with pytest.raises(RuntimeError, match="no CPU"):
        make()(torch.zeros(1, 3, 80, 80))


def test_engine_parameter_order_and_shapes():
    m = make(fmt="FP16_16")
    p = m._engine_params()
    assert len(p) == 2 + 12 * 2
    assert p[0].shape == (128, 768) and p[1].shape == (128,)
    assert p[2 + 2].shape == (384, 128) and p[2 + 8].shape == (256, 128) and p[2 + 10].shape == (128, 256)
    names = {id(v): k for k, v in m.named_parameters()}
    assert names[id(p[2 + 12 + 4])] == "transformer.layers.1.0.fn.fn.to_out.0.1.weight"
    # parameters that the reference leaves without gradient exist but are not engine-owned
    assert "det_tokens" in dict(m.named_parameters())


def test_same_seed_same_init_as_constructor_order():
    torch.manual_seed(7)
    a = make(fmt="FP32")
    torch.manual_seed(7)
    b = make(fmt="FP16_16")
    from oracle.vit_oracle import canonical_key
    sa = a.state_dict()
    for k, v in b.state_dict().items():
        assert torch.equal(v, sa[canonical_key(k)])


def test_get_batch_sizes():
    from myrtle_vision.utils.utils import get_batch_sizes
    assert get_batch_sizes(32, 2, 64) == (32, 1)            # shipped vit_small.json point
    assert get_batch_sizes(256, 8, 2048) == (256, 1)        # BASELINE config 3
    assert get_batch_sizes(32, 1, 128) == (32, 4)
    assert get_batch_sizes(32, 0, 64) == (32, 2)            # CPU: num_gpus == 0
    assert get_batch_sizes(32, 4, 120) == (30, 1)           # best divisor below the target
    assert get_batch_sizes(32, 4, 200) == (25, 2)
    with pytest.raises(ValueError, match="not divisible by the number of GPUs"):
        get_batch_sizes(32, 3, 64)


def test_get_models_reads_the_reference_json_schema(tmp_path):
    from myrtle_vision.utils.models import get_models
    from myrtle_vision.utils.utils import parse_config
    data = tmp_path / "data_config.json"
    data.write_text(json.dumps({"number_of_classes": 45}))
    cfg = {
        "train_config": {"local_batch_size": 32, "global_batch_size": 64, "seed": 1234},
        "data_config_path": str(data),
        "dist_config": {"dist_backend": "nccl", "dist_url": "tcp://localhost:54321"},
        "vit_config": {"decoder": "classification", "image_size": 224, "patch_size": 16,
                       "embed_dim": 192, "depth": 2, "heads": 3, "mlp_dim": 768, "dropout": 0.0,
                       "emb_dropout": 0.0, "q_format": "FP16_32"},
    }
    path = tmp_path / "train.json"
    path.write_text(json.dumps(cfg))
    vit, distiller = get_models(parse_config(str(path)))
    assert distiller is None
    assert vit.decoder.linear[1].weight.shape == (45, 192)
    assert vit.quantizer.q_format.name == "FP16_32"


def test_qtorch_facade_surface():
    import qtorch
    import qtorch.quant as qq
    assert repr(qtorch.FloatingPoint(5, 10)) == "FloatingPoint (exponent=5, mantissa=10)"
    with pytest.raises(AssertionError):
        qtorch.FloatingPoint(9, 10)
    for name in ("float_quantize", "fixed_point_quantize", "block_quantize", "quantizer", "Quantizer"):
        assert hasattr(qq, name)
    with pytest.raises(AssertionError, match="invalid rounding"):
        qq.quantizer(forward_rounding="up")
    import mv_native
    with pytest.raises(mv_native.MvError, match="CUDA tensors"):
        qq.float_quantize(torch.zeros(4), 5, 10, "nearest")


def test_optimizer_groups_and_cosine_schedule():
    """timm-0.5.4-shaped optimizer factory / cosine schedule used by the train.py entry points."""
    import argparse
    from myrtle_vision.utils.optim import add_weight_decay, create_optimizer, create_scheduler
    m = make()
    groups = add_weight_decay(m, 0.05)
    assert groups[0]["weight_decay"] == 0.0 and groups[1]["weight_decay"] == 0.05
    assert all(p.ndim >= 2 for p in groups[1]["params"])              # only matrices / tokens decay
    names = {id(p): n for n, p in m.named_parameters()}
    assert all(names[id(p)].endswith(".bias") or p.ndim <= 1 for p in groups[0]["params"])
    assert sum(len(g["params"]) for g in groups) == len(list(m.parameters()))
    args = argparse.Namespace(opt="adamw", opt_eps=1e-8, opt_betas=None, clip_grad=None, momentum=0.9,
                              weight_decay=0.05, sched="cosine", lr=6.25e-5, warmup_lr=1e-6, min_lr=1e-5,
                              epochs=300, decay_epochs=15, warmup_epochs=5, cooldown_epochs=5,
                              patience_epochs=5, decay_rate=0.1)
    opt = create_optimizer(args, m, fused=False)
    assert isinstance(opt, torch.optim.AdamW) and len(opt.param_groups) == 2
    sched, epochs = create_scheduler(args, opt)
    assert epochs == 305
    assert abs(opt.param_groups[0]["lr"] - 1e-6) < 1e-12              # warm-up start
    sched.step(3)
    assert abs(opt.param_groups[1]["lr"] - (1e-6 + 3 * (6.25e-5 - 1e-6) / 5)) < 1e-12
    sched.step(150)
    assert abs(opt.param_groups[0]["lr"] - (1e-5 + 0.5 * (6.25e-5 - 1e-5))) < 1e-10
    sched.step(300)
    assert opt.param_groups[0]["lr"] == 1e-5
    state = sched.state_dict()
    sched.step(10)
    sched.load_state_dict(state)
    assert opt.param_groups[0]["lr"] == 1e-5
    with pytest.raises(NotImplementedError):
        create_optimizer(argparse.Namespace(opt="lamb", weight_decay=0.0, lr=1e-3, opt_eps=None), m)


@pytest.mark.parametrize("fmt", ["FP32", "FP16_32", "FP16_16"])
def test_rename_timm_state_dict_loads_into_every_key_layout(fmt):
    """pretrained-backbone key mapping incl. the quantised `<name>.1.<param>` layout."""
    from myrtle_vision.utils.models import rename_timm_state_dict
    D, depth, M, P = 128, 2, 256, 16
    g = torch.Generator().manual_seed(0)
    timm_sd = {"cls_token": torch.randn(1, 1, D, generator=g), "pos_embed": torch.randn(1, 197, D, generator=g),
               "patch_embed.proj.weight": torch.randn(D, 3, P, P, generator=g), "patch_embed.proj.bias": torch.randn(D, generator=g),
               "norm.weight": torch.ones(D), "norm.bias": torch.zeros(D),
               "head.weight": torch.randn(5, D, generator=g), "head.bias": torch.zeros(5)}
    for i in range(depth):
        for name, shape in (("norm1.weight", (D,)), ("norm1.bias", (D,)), ("attn.qkv.weight", (3 * D, D)),
                            ("attn.qkv.bias", (3 * D,)), ("attn.proj.weight", (D, D)), ("attn.proj.bias", (D,)),
                            ("norm2.weight", (D,)), ("norm2.bias", (D,)), ("mlp.fc1.weight", (M, D)),
                            ("mlp.fc1.bias", (M,)), ("mlp.fc2.weight", (D, M)), ("mlp.fc2.bias", (D,))):
            timm_sd["blocks.%d.%s" % (i, name)] = torch.randn(*shape, generator=g)
    cfg = {"embed_dim": D, "patch_size": P, "q_format": fmt}
    sd = rename_timm_state_dict(timm_sd, cfg, 5)
    m = make("classification", fmt, image_size=224)
    res = m.load_state_dict(sd, strict=False)
    assert res.unexpected_keys == []
    assert sorted(res.missing_keys) == sorted(k for k in m.state_dict() if k.startswith("decoder.") or "det" in k)
    pe = m.state_dict()["patch_to_embedding.1.weight" if fmt != "FP32" else "patch_to_embedding.weight"]
    w = timm_sd["patch_embed.proj.weight"]
    assert pe.shape == (D, 3 * P * P) and torch.equal(pe[7, (3 * P + 5) * 3 + 2], w[7, 2, 3, 5])   # (O,(H,W,I))
    qkv_key = "transformer.layers.1.0.fn.fn.to_qkv%s.weight" % (".1" if fmt != "FP32" else "")
    assert torch.equal(m.state_dict()[qkv_key], timm_sd["blocks.1.attn.qkv.weight"])


def test_miou_matches_the_histogram_definition():
    """utils/miou.py: confusion-matrix mIoU equals the reference's histc formulation (restated inline)."""
    from myrtle_vision.utils.miou import MIoU
    g = torch.Generator().manual_seed(5)
    C = 6
    m = MIoU(C, "cpu")
    ti, tu = torch.zeros(C, dtype=torch.float64), torch.zeros(C, dtype=torch.float64)
    for _ in range(3):
        pred = torch.randint(0, C, (2, 40, 40), generator=g)
        lab = torch.randint(0, C, (2, 40, 40), generator=g)
        m.add_img(pred, lab)
        inter = torch.histc(pred[pred == lab].float(), bins=C, min=0, max=C - 1)
        ap = torch.histc(pred.float(), bins=C, min=0, max=C - 1)
        al = torch.histc(lab.float(), bins=C, min=0, max=C - 1)
        ti += inter
        tu += ap + al - inter
    assert torch.allclose(m.get_per_class_iou(), ti / tu)
    assert abs(m.get_miou() - float((ti / tu).mean())) < 1e-12


def test_box_average_precision():
    """COCO-style AP of detection/test.py (myrtle_vision/utils/evaluate.py) on hand-checked cases."""
    import torch
    from myrtle_vision.utils.evaluate import box_average_precision, pairwise_iou
    gt = [{"labels": torch.tensor([0, 1, 1]),
           "boxes": torch.tensor([[0., 0, 10, 10], [20, 20, 40, 40], [50, 50, 60, 70]])}]
    perfect = [{"scores": torch.tensor([0.9, 0.8, 0.7, 0.1]), "labels": torch.tensor([0, 1, 1, 1]),
                "boxes": torch.tensor([[0., 0, 10, 10], [20, 20, 40, 40], [50, 50, 60, 70], [0, 0, 5, 5]])}]
    res = box_average_precision(perfect, gt, 3)
    assert res["AP"] == 1.0 and res["AP50"] == 1.0          # the low-score false positive comes after full recall
    # class 0 localised at IoU 0.8 (a hit for 7 of the 10 thresholds), class 1 missed entirely
    loose = [{"scores": torch.tensor([0.9, 0.8]), "labels": torch.tensor([0, 1]),
              "boxes": torch.tensor([[0., 0, 10, 8], [100, 100, 110, 110]])}]
    res = box_average_precision(loose, gt, 3)
    assert abs(res["per_class"][0] - 0.7) < 1e-9 and res["per_class"][1] == 0.0
    assert abs(res["AP"] - 0.35) < 1e-9 and abs(res["AP50"] - 0.5) < 1e-9
    # a confident false positive ahead of the true positive halves the precision at every recall level
    fp_first = [{"scores": torch.tensor([0.9, 0.5]), "labels": torch.tensor([0, 0]),
                 "boxes": torch.tensor([[30., 30, 35, 35], [0, 0, 10, 10]])}]
    assert abs(box_average_precision(fp_first, gt, 3)["per_class"][0] - 0.5) < 1e-9
    assert abs(float(pairwise_iou(torch.tensor([[0., 0, 10, 10]]), torch.tensor([[0., 0, 10, 8]]))) - 0.8) < 1e-6


def test_backward_format_is_an_fp16_32_option_of_prepare_qat():
    """ModelQuantizer.prepare_qat(q_format, backward_format=(exp, man)): the plan carries the gradient format;
    formats with more quantiser sites than FP16_32's are refused rather than half-done."""
    import torch.nn as nn
    from myrtle_vision.utils.quantize import ModelQuantizer, QFormat
    net = nn.Sequential(nn.LayerNorm(8), nn.Linear(8, 8))
    mq = ModelQuantizer(net)
    mq.prepare_qat(QFormat.FP16_32, backward_format=(5, 10))
    assert mq.plan.inp == (5, 10) and mq.plan.grad == (5, 10)
    assert ModelQuantizer(nn.Sequential(nn.Linear(4, 4))).__class__ is ModelQuantizer
    mq2 = ModelQuantizer(nn.Sequential(nn.Linear(4, 4)))
    mq2.prepare_qat("FP16_32")
    assert mq2.plan.grad is None
    for fmt in ("FP16_16", "TF32", "FP32"):
        with pytest.raises(NotImplementedError):
            ModelQuantizer(nn.Sequential(nn.Linear(4, 4))).prepare_qat(fmt, backward_format=(5, 10))
    with pytest.raises(ValueError):
        ModelQuantizer(nn.Sequential(nn.Linear(4, 4))).prepare_qat("FP16_32", backward_format=(9, 3))
