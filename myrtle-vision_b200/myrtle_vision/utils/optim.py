"""Optimizer and learning-rate schedule of the train.py entry points.

The reference builds both with timm 0.5.4 (`timm.optim.create_optimizer`, `timm.scheduler.
create_scheduler`; classification/train.py:161-166) from the Namespace made by
`get_optimizer_args` (utils/models.py:84-110).  timm is not a dependency here; the two pieces the
shipped configs use are restated [recall of timm 0.5.4 — optim_factory.py / cosine_lr.py]:

* `adamw`: AdamW where parameters with ndim <= 1, names ending in ".bias" and names in
  `model.no_weight_decay()` get weight_decay 0, everything else `args.weight_decay`.
* `cosine`: per-epoch schedule, linear warm-up from `warmup_lr` over `warmup_epochs`, then
  `min_lr + 0.5 (lr - min_lr)(1 + cos(pi t / epochs))` (warm-up epochs are NOT a prefix), `min_lr`
  after `epochs`; training runs `epochs + cooldown_epochs` epochs.

`FusedAdamW` (utils/fused_adamw.py) is the B200-side optimizer the entry points use when the model
is on a CUDA device; the torch.optim path below is its parity reference.
"""
import math

import torch


def add_weight_decay(model, weight_decay=1e-5, skip_list=()):
    decay, no_decay = [], []
    for name, param in model.named_parameters():
        if not param.requires_grad:
            continue
        if param.ndim <= 1 or name.endswith(".bias") or name in skip_list:
            no_decay.append(param)
        else:
            decay.append(param)
    return [{"params": no_decay, "weight_decay": 0.0},
            {"params": decay, "weight_decay": weight_decay}]


def optimizer_kwargs(args):
    kw = dict(lr=args.lr, eps=args.opt_eps if args.opt_eps is not None else 1e-8)
    if getattr(args, "opt_betas", None) is not None:
        kw["betas"] = tuple(args.opt_betas)
    return kw


def create_optimizer(args, model, filter_bias_and_bn=True, fused=None):
    """args: the Namespace of get_optimizer_args.  fused=None picks FusedAdamW for CUDA models."""
    opt = args.opt.lower()
    weight_decay = args.weight_decay
    if weight_decay and filter_bias_and_bn:
        skip = model.no_weight_decay() if hasattr(model, "no_weight_decay") else ()
        params = add_weight_decay(model, weight_decay, skip)
        weight_decay = 0.0
    else:
        params = [{"params": [p for p in model.parameters() if p.requires_grad]}]
    if opt == "adamw":
        on_cuda = all(p.is_cuda for g in params for p in g["params"])
        if fused is None:
            fused = on_cuda
        if fused:
            from myrtle_vision.utils.fused_adamw import FusedAdamW
            return FusedAdamW(params, weight_decay=weight_decay, model=model, **optimizer_kwargs(args))
        return torch.optim.AdamW(params, weight_decay=weight_decay, **optimizer_kwargs(args))
    if opt in ("sgd", "nesterov", "momentum"):
        return torch.optim.SGD(params, lr=args.lr, momentum=args.momentum, weight_decay=weight_decay,
                               nesterov=opt != "momentum")
    raise NotImplementedError("optimizer %r (the shipped configs use adamw)" % args.opt)


class CosineSchedule:
    """Epoch-indexed cosine decay with linear warm-up; `step(epoch)` sets the lr of every group."""

    def __init__(self, optimizer, t_initial, lr_min=0.0, warmup_t=0, warmup_lr_init=0.0):
        self.optimizer = optimizer
        self.t_initial, self.lr_min = t_initial, lr_min
        self.warmup_t, self.warmup_lr_init = warmup_t, warmup_lr_init
        for group in optimizer.param_groups:
            group.setdefault("initial_lr", group["lr"])
        self.base = [g["initial_lr"] for g in optimizer.param_groups]
        self.last_epoch = None
        if warmup_t > 0:
            self._apply([warmup_lr_init] * len(self.base))

    def values(self, t):
        if t < self.warmup_t:
            return [self.warmup_lr_init + t * (b - self.warmup_lr_init) / self.warmup_t for b in self.base]
        if t < self.t_initial:
            return [self.lr_min + 0.5 * (b - self.lr_min) * (1 + math.cos(math.pi * t / self.t_initial))
                    for b in self.base]
        return [self.lr_min] * len(self.base)

    def _apply(self, values):
        for group, lr in zip(self.optimizer.param_groups, values):
            group["lr"] = lr

    def step(self, epoch, metric=None):
        self.last_epoch = epoch
        self._apply(self.values(epoch))

    def get_cycle_length(self):
        return self.t_initial

    def state_dict(self):
        return {"last_epoch": self.last_epoch, "base": self.base}

    def load_state_dict(self, state):
        self.last_epoch, self.base = state["last_epoch"], state["base"]
        if self.last_epoch is not None:
            self._apply(self.values(self.last_epoch))


def create_scheduler(args, optimizer):
    """-> (scheduler, total epochs incl. cool-down), like timm.scheduler.create_scheduler."""
    if args.sched != "cosine":
        raise NotImplementedError("scheduler %r (the shipped configs use cosine)" % args.sched)
    sched = CosineSchedule(optimizer, t_initial=args.epochs, lr_min=args.min_lr,
                           warmup_t=args.warmup_epochs, warmup_lr_init=args.warmup_lr)
    return sched, sched.get_cycle_length() + args.cooldown_epochs
