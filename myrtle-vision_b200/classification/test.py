"""Evaluate a checkpoint on the test split: `python test.py -c config.json` (reference classification/test.py; the flow
is myrtle_vision/utils/evaluate.py:test_deit)."""
import argparse
import json
import os
import sys

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))

from myrtle_vision.utils.evaluate import test_deit as _test_deit


def test_deit(config):
    return _test_deit(config, "classification")


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
    test_deit(config)
