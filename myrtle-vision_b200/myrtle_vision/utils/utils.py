"""Config parsing, seeding, batch-size solving and process-group setup — the data-parallel
contract of the reference (src/myrtle_vision/utils/utils.py:70-147), same names and behaviour."""
import json
import os
import random

import numpy as np
import torch
import torch.distributed as dist


def parse_config(config_path):
    with open(config_path) as f:
        return json.loads(f.read())


def seed_everything(seed):
    random.seed(seed)
    os.environ["PYTHONHASHSEED"] = str(seed)
    np.random.seed(seed)
    torch.manual_seed(seed)
    if torch.cuda.is_available():
        torch.cuda.manual_seed(seed)
        torch.cuda.manual_seed_all(seed)
    torch.backends.cudnn.deterministic = True


def get_batch_sizes(target_batch, num_gpus, global_batch, verbose=False):
    """(local batch, gradient-accumulation steps) for a target local batch and a global batch.

    Mirrors the reference's three cases (utils.py:86-125): exact multiple; divisible by the GPU
    count only (largest divisor of the per-GPU share below the target); otherwise ValueError."""
    per_step = num_gpus * target_batch if num_gpus > 0 else target_batch
    if global_batch % per_step == 0:
        return target_batch, global_batch // per_step
    if num_gpus > 0 and global_batch % num_gpus == 0:
        share = global_batch // num_gpus
        local = target_batch - 1
        while share % local != 0:
            local -= 1
        if verbose:
            print("WARNING: Did not select preferred max local batch size "
                  f"{target_batch}; using a local batch size of {local} instead")
        return local, share // local
    raise ValueError(
        "WARNING: Could not fulfill the desired global batch size of "
        f"{global_batch} as it is not divisible by the number of GPUs "
        f" available ({num_gpus})\nPlease update the global_batch_size "
        "parameter in your config file or change the number of GPUs "
        "available (e.g. with CUDA_VISIBLE_DEVICES)")


def init_distributed(rank, num_gpus, dist_backend, dist_url, group_name=None):
    """One process per GPU; NCCL over NVLink for `dist_backend == "nccl"`."""
    if dist_backend == "nccl":
        assert torch.cuda.is_available(), "Distributed mode requires CUDA."
        torch.cuda.set_device(rank)
    if rank == 0:
        print("Initializing Distributed")
    dist.init_process_group(dist_backend, init_method=dist_url, world_size=num_gpus, rank=rank,
                            group_name=group_name or "")


def cleanup_distributed():
    dist.destroy_process_group()


# ---- rank helpers used by the detection criterion / evaluation (reference utils.py:153-260)
def is_dist_avail_and_initialized():
    return dist.is_available() and dist.is_initialized()


def get_world_size():
    return dist.get_world_size() if is_dist_avail_and_initialized() else 1


def get_rank():
    return dist.get_rank() if is_dist_avail_and_initialized() else 0


def all_gather(data):
    """Gather any picklable object from every rank -> list (one entry per rank)."""
    if get_world_size() == 1:
        return [data]
    out = [None] * get_world_size()
    dist.all_gather_object(out, data)
    return out


def reduce_dict(input_dict, average=True):
    """Sum (or mean) a dict of scalar tensors over ranks; keys are sorted so ranks agree on order."""
    world = get_world_size()
    if world < 2:
        return input_dict
    with torch.no_grad():
        names = sorted(input_dict.keys())
        values = torch.stack([input_dict[k] for k in names], dim=0)
        dist.all_reduce(values)
        if average:
            values /= world
        return dict(zip(names, values))


@torch.no_grad()
def accuracy(output, target, topk=(1,)):
    """precision@k in percent, one 0-dim tensor per k; zeros for an empty target."""
    if target.numel() == 0:
        return [torch.zeros([], device=output.device)]
    top = output.topk(max(topk), dim=1).indices                       # [n, maxk]
    hit = top.eq(target.view(-1, 1))
    return [hit[:, :k].any(dim=1).float().sum() * (100.0 / target.size(0)) for k in topk]
