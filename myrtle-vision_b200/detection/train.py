"""detection training entry point: `python train.py -c train_configs/vit_small.json`.

Drop-in for the reference's detection/train.py (same CLI, same config schema, same
`train_deit(rank, num_gpus, config)`); the loop itself lives in myrtle_vision/utils/trainer.py."""
import argparse
import json
import os
import sys
from datetime import datetime

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))

import torch
import torch.multiprocessing as mp

from myrtle_vision.utils.trainer import train_deit as _train


def train_deit(rank, num_gpus, config):
    return _train(rank, num_gpus, config, task="detection")


if __name__ == "__main__":
    parser = argparse.ArgumentParser()
    parser.add_argument("-c", "--config", type=str, help="JSON file for configuration")
    args = parser.parse_args()
    with open(args.config) as f:
        config = json.load(f)
    base = os.path.dirname(os.path.abspath(args.config))
    if "data_config_path" in config and not os.path.isabs(config["data_config_path"]):
        here = os.path.join(os.path.dirname(base), config["data_config_path"])
        if not os.path.exists(config["data_config_path"]) and os.path.exists(here):
            config["data_config_path"] = here
    config["train_config"]["output_directory"] += datetime.now().strftime("_%m_%d_%Y_%H_%M_%S")
    num_gpus = torch.cuda.device_count()
    if config["train_config"]["distributed"]:
        if num_gpus <= 1:
            print(f"WARNING: tried to enable distributed training but only found {num_gpus} GPU(s)")
    elif num_gpus > 1:
        print("INFO: you have multiple GPUs available but did not enable distributed training")
        num_gpus = 1
    try:
        if num_gpus > 1:
            mp.spawn(train_deit, args=(num_gpus, config), nprocs=num_gpus, join=True)
        else:
            train_deit(0, num_gpus, config)
    except KeyboardInterrupt:
        print("Ctrl-c pressed; cleaning up and ending training early...")
