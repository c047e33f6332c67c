"""Model factory and checkpoint helpers (reference: src/myrtle_vision/utils/models.py:25-141).

`get_models(config)` keeps the JSON schema (`vit_config`, `data_config_path`) and returns
`(vit, None)`; the distillation wrapper of the reference is broken upstream and out of scope
(SURVEY.md §2 row 11), so a `distiller_config` raises."""
import torch

from myrtle_vision.models.vit import ViT
from myrtle_vision.utils.quantize import QFormat
from myrtle_vision.utils.utils import parse_config


def get_models(config, profile=False):
    vit_config = config["vit_config"]
    data_config = config.get("data_config")
    if data_config is None:
        data_config = parse_config(config["data_config_path"])
    if "distiller_config" in config:
        raise NotImplementedError("distillation is outside the B200 hot path")
    vit = ViT(
        decoder=vit_config["decoder"],
        image_size=vit_config["image_size"],
        patch_size=vit_config["patch_size"],
        num_classes=data_config["number_of_classes"],
        dim=vit_config["embed_dim"],
        depth=vit_config["depth"],
        heads=vit_config["heads"],
        mlp_dim=vit_config["mlp_dim"],
        dropout=vit_config["dropout"],
        emb_dropout=vit_config["emb_dropout"],
        profile=profile,
        q_format=QFormat[vit_config["q_format"]],
    )
    return vit, None


def save_checkpoint(model, optimizer, lr_scheduler, iteration, filepath):
    torch.save({
        "model": model.state_dict(),
        "optimizer": optimizer.state_dict(),
        "lr_scheduler": lr_scheduler.state_dict() if lr_scheduler is not None else None,
        "iteration": iteration,
    }, filepath)


def load_checkpoint(model, optimizer, lr_scheduler, filepath):
    checkpoint = torch.load(filepath, map_location="cpu")
    model.load_state_dict(checkpoint["model"])
    if optimizer is not None:
        optimizer.load_state_dict(checkpoint["optimizer"])
    if lr_scheduler is not None and checkpoint.get("lr_scheduler") is not None:
        lr_scheduler.load_state_dict(checkpoint["lr_scheduler"])
    return checkpoint["iteration"]


def prepare_model_and_load_ckpt(train_config, model, optimizer=None, lr_scheduler=None):
    if train_config["checkpoint_path"] != "":
        return load_checkpoint(model=model, optimizer=optimizer, lr_scheduler=lr_scheduler,
                               filepath=train_config["checkpoint_path"])
    return 0


def get_optimizer_args(train_config):
    """train_config -> the Namespace the optimizer / scheduler factories read (reference :84-110)."""
    import argparse
    cfg = train_config
    return argparse.Namespace(
        opt=cfg["optimizer"], opt_eps=cfg["opt_eps"], opt_betas=cfg["opt_betas"], clip_grad=cfg["clip_grad"],
        momentum=cfg["momentum"], weight_decay=cfg["weight_decay"], sched=cfg["scheduler"], lr=cfg["lr"],
        lr_noise=cfg.get("lr_noise"), lr_noise_pct=cfg.get("lr_noise_pct"), lr_noise_std=cfg.get("lr_noise_std"),
        warmup_lr=cfg["warmup_lr"], min_lr=cfg["min_lr"], epochs=cfg["epochs"], decay_epochs=cfg["decay_epochs"],
        warmup_epochs=cfg["warmup_epochs"], cooldown_epochs=cfg["cooldown_epochs"],
        patience_epochs=cfg["patience_epochs"], decay_rate=cfg["decay_rate"])
