"""CLI (reference src/inference.py): python inference.py -f CONFIG [-o OVERRIDE] [--output_dir ...] [--model_path ...]

Loads the checkpoint (config['test']['model_path'] or <output_dir>/<exp_name>/best-<target_metric>.pth), runs the
eval-mode forward + soft-max over the test split and writes test_outputs.npy, as the reference does
(inference.py:75-126); the per-batch work is one CUDA graph (semiseg_b200.evaluate.predict_loader)."""
import argparse
import os

import numpy as np
import torch

from algorithms.base import init_model_from_cfg
from semiseg_b200.evaluate import predict_loader
from utils.config import load_config
from utils.semi_dataset import build_seg_dataset, get_dataloader


def parse() -> dict:
    p = argparse.ArgumentParser("ECG segmentation inference")
    p.add_argument("-f", "--config_path", required=True, type=str, metavar="FILE", help="YAML config file path")
    p.add_argument("-o", "--override_config_path", default=None, type=str, metavar="FILE", help="YAML override")
    p.add_argument("--output_dir", default="", type=str, metavar="DIR", help="path where to save")
    p.add_argument("--exp_name", default="", type=str, help="experiment name")
    p.add_argument("--model_path", default="", type=str, metavar="PATH", help="saved from checkpoint")
    a = p.parse_args()
    config = load_config(a.config_path, a.override_config_path, {"output_dir": a.output_dir, "exp_name": a.exp_name})
    if a.model_path:
        config.setdefault("test", {})["model_path"] = a.model_path
    return config


def inference(config):
    output_dir = os.path.join(config["output_dir"], config["exp_name"])
    os.makedirs(output_dir, exist_ok=True)
    device = torch.device(config["device"])
    ds = build_seg_dataset(config["dataset"], split="test")
    loader = get_dataloader(ds, is_distributed=False, mode="test", **config["dataloader"])
    model = init_model_from_cfg(config, train=False)
    if config["test"].get("model_path", None):
        path = config["test"]["model_path"]
    else:
        path = os.path.join(output_dir, f"best-{config['test'].get('target_metric', 'loss')}.pth")
    assert os.path.exists(path), f"Checkpoint not found: {path}"
    state_dict = torch.load(path, map_location="cpu", weights_only=False)["model"]
    for k in list(state_dict.keys()):          # drop the auxiliary head
        if k.startswith("auxiliary_head"):
            del state_dict[k]
    print(model.load_state_dict(state_dict))
    model.to(device)
    outputs = predict_loader(model, loader, device, use_amp=config["test"].get("use_amp", False)).numpy()
    np.save(os.path.join(output_dir, "test_outputs.npy"), outputs)
    print("Done!")
    return outputs


if __name__ == "__main__":
    inference(parse())
