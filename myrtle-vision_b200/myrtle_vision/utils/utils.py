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
