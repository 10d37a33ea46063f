"""CLI (reference src/train.py): python train.py -f CONFIG [-o OVERRIDE] [--output_dir ...]"""
import argparse

import algorithms
from utils.config import load_config


def parse():
    p = argparse.ArgumentParser("SemiSegECG B200 training")
    p.add_argument("-f", "--config_path", required=True, type=str, help="YAML config file path")
    p.add_argument("-o", "--override_config_path", default=None, type=str, help="YAML override")
    p.add_argument("--output_dir", default="", type=str)
    p.add_argument("--exp_name", default="", type=str)
    p.add_argument("--resume", default="", type=str)
    p.add_argument("--start_epoch", default=0, type=int)
    p.add_argument("--test", action="store_true")
    a = p.parse_args()
    cli = {k: getattr(a, k) for k in ("output_dir", "exp_name", "resume", "start_epoch")}
    return load_config(a.config_path, a.override_config_path, cli), a.test


if __name__ == "__main__":
    config, run_test = parse()
    assert config["algorithm"] in algorithms.__dict__, f"Unsupported algorithm: {config['algorithm']}"
    algo = algorithms.__dict__[config["algorithm"]]
    algo.train(config)
    if run_test:
        algo.test(config)
