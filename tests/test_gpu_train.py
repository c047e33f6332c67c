"""The train.py entry points (north star: drop-in classification / segmentation / detection training)
run end to end on the fused path: synthetic data -> DataLoader -> (graphed) fwd+loss+bwd -> FusedAdamW."""
import copy
import os

import pytest
import torch

pytestmark = pytest.mark.gpu

TRAIN = {"output_directory": "", "checkpoint_path": "", "epochs": 12, "local_batch_size": 4,
         "global_batch_size": 4, "iters_per_checkpoint": 5, "iters_per_val": 5, "seed": 1234,
         "drop_last_batch": True, "optimizer": "adamw", "opt_eps": 1e-8, "opt_betas": None, "clip_grad": None,
         "momentum": 0.9, "weight_decay": 0.05, "scheduler": "cosine", "lr": 2e-3, "warmup_lr": 2e-3,
         "min_lr": 1e-4, "decay_epochs": 15, "warmup_epochs": 0, "cooldown_epochs": 0, "patience_epochs": 5,
         "decay_rate": 0.1, "distributed": False, "num_workers": 0,
         "loss_ce": 1.0, "class_error": 0.0, "loss_bbox": 5.0, "loss_giou": 2.0, "cardinality_error": 0.0,
         "eos_coef": 0.1}
VIT = {"patch_size": 16, "embed_dim": 128, "depth": 2, "heads": 2, "mlp_dim": 256, "dropout": 0.0,
       "emb_dropout": 0.0}


def config(task, tmp_path, fmt, size, classes):
    return {"train_config": dict(TRAIN, output_directory=str(tmp_path / "ckpt")),
            "dist_config": {"dist_backend": "nccl", "dist_url": "tcp://127.0.0.1:54399"},
            "vit_config": dict(VIT, decoder=task, image_size=size, q_format=fmt),
            "data_config": {"number_of_classes": classes,
                            "synthetic": {"train_length": 4, "val_length": 4, "seed": 7}}}


@pytest.mark.parametrize("task,fmt,size,classes", [("classification", "FP16_32", 80, 5),
                                                   ("classification", "FP16_16", 80, 5),
                                                   ("segmentation", "FP16_32", 96, 4),
                                                   ("detection", "FP16_32", 176, 3)])
def test_entry_point_trains(task, fmt, size, classes, tmp_path):
    import importlib.util
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    spec = importlib.util.spec_from_file_location("train_%s" % task,
                                                  os.path.join(root, "myrtle-vision_b200", task, "train.py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    cfg = config(task, tmp_path, fmt, size, classes)
    history = mod.train_deit(0, 1, copy.deepcopy(cfg))
    assert len(history) == 12 and all(torch.isfinite(torch.tensor(history)))
    assert history[-1] < 0.8 * history[0], history          # memorises its 4 samples
    ckpts = sorted(os.listdir(cfg["train_config"]["output_directory"]))
    assert ckpts[:2] == ["vit_000000", "vit_000005"]
    # resume from a checkpoint: iteration counter and optimizer state come back
    cfg2 = copy.deepcopy(cfg)
    cfg2["train_config"]["checkpoint_path"] = os.path.join(cfg["train_config"]["output_directory"], "vit_000010")
    from myrtle_vision.utils.trainer import train_deit
    more = train_deit(0, 1, cfg2, max_iterations=12)
    assert len(more) == 2 and more[-1] < history[0]


@pytest.mark.parametrize("task,size,classes", [("classification", 80, 5), ("segmentation", 96, 4), ("detection", 176, 3)])
def test_eval_cli_on_a_trained_checkpoint(task, size, classes, tmp_path, capsys):
    """{classification,segmentation,detection}/test.py: checkpoint -> eval forward -> the task's report."""
    import importlib.util
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    cfg = config(task, tmp_path, "FP16_32", size, classes)
    from myrtle_vision.utils.trainer import train_deit
    train_deit(0, 1, copy.deepcopy(cfg), max_iterations=6)
    spec = importlib.util.spec_from_file_location("test_cli_%s" % task,
                                                  os.path.join(root, "myrtle-vision_b200", task, "test.py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    cfg2 = copy.deepcopy(cfg)
    with pytest.raises(AssertionError, match="checkpoint path"):
        mod.test_deit(copy.deepcopy(cfg2))
    cfg2["train_config"]["checkpoint_path"] = os.path.join(cfg["train_config"]["output_directory"], "vit_000005")
    res = mod.test_deit(cfg2)
    out = capsys.readouterr().out
    if task == "classification":
        assert 0.0 <= res["accuracy"] <= 1.0 and "precision" in out
    elif task == "segmentation":
        assert 0.0 <= res["miou"] <= 1.0 and "mIoU is:" in out and res["per_class_iou"].numel() == classes
    else:
        assert "IoU=0.50:0.95" in out and (res["AP"] != res["AP"] or 0.0 <= res["AP"] <= 1.0)


def test_ptq_eval_flow_from_an_fp32_checkpoint(tmp_path):
    """classification/test_quantize.py — FP32 checkpoint -> prepare_qat -> convert -> eval."""
    import importlib.util
    import torch.nn as nn
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    cfg = config("classification", tmp_path, "FP32", 80, 5)
    cfg["train_config"]["epochs"] = 6
    from myrtle_vision.utils.trainer import train_deit
    train_deit(0, 1, copy.deepcopy(cfg), max_iterations=6)
    ckpt = os.path.join(cfg["train_config"]["output_directory"], "vit_000005")
    assert os.path.exists(ckpt)
    spec = importlib.util.spec_from_file_location(
        "test_quantize_cli", os.path.join(root, "myrtle-vision_b200", "classification", "test_quantize.py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    ev = copy.deepcopy(cfg)
    ev["train_config"]["checkpoint_path"] = ckpt
    ev["vit_config"]["q_format"] = "FP16_32"
    acc = mod.test_deit(ev, calib_steps=1, quantized_ckpt=False)
    assert 0.0 <= acc <= 1.0
    # the converted model's Linear weights and LayerNorm gammas sit on the fp16 grid
    from myrtle_vision.utils.models import get_models, prepare_model_and_load_ckpt
    vit, _ = get_models(copy.deepcopy({**ev, "vit_config": {**ev["vit_config"], "q_format": "FP32"}}))
    vit = vit.cuda()
    prepare_model_and_load_ckpt(train_config=ev["train_config"], model=vit)
    vit.quantizer.prepare_qat("FP16_32")
    vit.convert()
    for m in vit.modules():
        if isinstance(m, (nn.Linear, nn.LayerNorm)):
            assert torch.equal(m.weight, m.weight.half().float())


def test_overflow_flag_skips_the_step_and_backs_off():
    """The device-side replacement of the reference's GradScaler (classification/train.py:167, 259-277): an inf in the
    incoming gradient or a saturated fp16 gradient operand raises found_inf without a host sync, FusedAdamW skips the
    update, the operand scale backs off by 16 and recovers after clean steps."""
    import torch.nn.functional as F
    from myrtle_vision.models.vit import ViT
    from myrtle_vision.utils.fused_adamw import FusedAdamW
    from myrtle_vision.utils.optim import add_weight_decay
    dev = "cuda"
    torch.manual_seed(2)
    m = ViT(decoder="classification", image_size=80, patch_size=16, num_classes=5, dim=128, depth=2, heads=2,
            mlp_dim=256, q_format="FP16_32").to(dev).train()
    opt = FusedAdamW(add_weight_decay(m, 0.05), lr=1e-2, model=m)
    img = torch.randn(4, 3, 80, 80, device=dev).clamp(-1, 1)
    tgt = torch.randint(0, 5, (4,), device=dev)
    eng = m.engine()

    def step(scale=1.0):
        m.zero_grad(set_to_none=True)
        (F.cross_entropy(m(img), tgt) * scale).backward()
        found = float(eng.found_inf)
        opt.step(found_inf=eng.found_inf)
        eng.found_inf.zero_()
        return found

    w = dict(m.named_parameters())["transformer.layers.0.0.fn.fn.to_qkv.1.weight"]
    w0 = w.detach().clone()
    assert step() == 0.0 and float(eng.scaler_state[1]) == 1024.0
    w1 = w.detach().clone()
    assert not torch.equal(w1, w0)                                   # a clean step updates
    assert step(float("inf")) == 1.0                                 # inf / NaN in the incoming gradient
    assert torch.equal(w, w1) and float(eng.scaler_state[1]) == 64.0  # skipped, scale backed off by 16
    eng.scaler_state[1] = 2.0 ** 40                                  # operands far outside fp16: saturation
    assert step() == 1.0
    assert torch.equal(w, w1) and float(eng.scaler_state[1]) == 2.0 ** 36
    eng.scaler_state[1] = 1024.0
    assert step() == 0.0 and not torch.equal(w, w1)
    # growth after `growth_interval` clean backwards (mv_overflow_update)
    import mv_native
    eng.scaler_state[1] = 64.0
    eng.scaler_state[2] = 0.0
    for _ in range(3):
        mv_native.overflow_update(eng.overflow.zero_(), eng.scaler_state, growth_interval=3)
    assert float(eng.scaler_state[1]) == 128.0 and float(eng.scaler_state[2]) == 0.0
