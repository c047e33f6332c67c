"""Post-training quantisation + evaluation: `python test_quantize.py -c config.json [--calib-steps N] [--quantized-ckpt]`.

Drop-in for the reference's classification/test_quantize.py:37-134: build the model, load a
checkpoint (an FP32 one is re-prepared for `vit_config.q_format` after loading, a quantised one is loaded as
is), run the calibration batches (a no-op for the float formats, kept for interface parity), `vit.convert()`
— Linear weights and LayerNorm gammas become their quantised values — and report test accuracy.  The forward
is the fused sm_100a path with the weight operands quantised once (they no longer change).
Data: the synthetic RESISC45-shaped sets, or `data_config["dataset_factory"]` (see utils/trainer.py)."""
import argparse
import json
import os
import sys
import tempfile

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))

import numpy as np
import torch
from torch.utils.data import DataLoader

from myrtle_vision.utils.models import get_models, prepare_model_and_load_ckpt
from myrtle_vision.utils.quantize import QFormat
from myrtle_vision.utils.trainer import build_datasets
from myrtle_vision.utils.utils import parse_config


def model_size(model):
    """Size of the model's state_dict in MB."""
    with tempfile.NamedTemporaryFile() as temp:
        torch.save(model.state_dict(), temp.name)
        return os.path.getsize(temp.name) / 1e6


def calibrate(model, loader, calib_steps, device):
    print(f"\nRunning {calib_steps} calibration steps")
    with torch.no_grad():
        for i, (imgs, _) in enumerate(loader):
            if i >= calib_steps:
                break
            model(imgs.to(device))


def test_deit(config, calib_steps, quantized_ckpt):
    train_config = config["train_config"]
    data_config = config.get("data_config") or parse_config(config["data_config_path"])
    config["data_config"] = data_config
    q_format = QFormat[config["vit_config"]["q_format"]]
    if q_format == QFormat.PyTorchINT8:
        raise NotImplementedError("PyTorchINT8 is torch's CPU-only int8 path, outside the B200 hot path")
    device = torch.device("cuda")
    _, testset, collate = build_datasets("classification", data_config, config["vit_config"])
    loader = DataLoader(testset, num_workers=0, batch_size=train_config["local_batch_size"],
                        drop_last=train_config["drop_last_batch"], collate_fn=collate)
    config["vit_config"]["dropout"] = 0.0
    config["vit_config"]["emb_dropout"] = 0.0
    if not quantized_ckpt:
        config["vit_config"]["q_format"] = "FP32"
    vit, _ = get_models(config)
    vit = vit.to(device)
    assert train_config["checkpoint_path"] != "", "Must provide a checkpoint path in the config file"
    prepare_model_and_load_ckpt(train_config=train_config, model=vit)
    if not quantized_ckpt:
        vit.quantizer.prepare_qat(q_format)
    print(f"Pre-quantization model size: {model_size(vit)} MB")
    vit.eval()
    calibrate(vit, loader, calib_steps, device)
    vit.convert()
    print(f"\nPost-quantization model size: {model_size(vit)} MB")
    truth, pred = [], []
    with torch.no_grad():
        for imgs, labels in loader:
            out = vit(imgs.to(device))
            pred.extend(out.argmax(dim=1).cpu().numpy())
            truth.extend(np.asarray(labels))
    accuracy = float(np.mean(np.asarray(truth) == np.asarray(pred))) if truth else float("nan")
    try:
        from sklearn.metrics import classification_report
        print(classification_report(truth, pred, labels=np.arange(data_config["number_of_classes"]), zero_division=0))
    except ImportError:
        print(f"accuracy: {accuracy:.4f}")
    return accuracy


if __name__ == "__main__":
    parser = argparse.ArgumentParser()
    parser.add_argument("-c", "--config", type=str, help="JSON file for configuration")
    parser.add_argument("--calib-steps", type=int, default=10)
    parser.add_argument("--quantized-ckpt", action="store_true",
                        help="the checkpoint was trained with q_format already applied")
    args = parser.parse_args()
    with open(args.config) as f:
        cfg = json.load(f)
    test_deit(cfg, args.calib_steps, args.quantized_ckpt)
