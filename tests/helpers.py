"""Shared test helpers (oracle-side)."""
import os
import sys

import numpy as np
import torch

REPO = os.path.abspath(os.path.join(os.path.dirname(__file__), ".."))
sys.path.insert(0, REPO)
sys.path.insert(0, os.path.join(REPO, "semi-seg-ecg_b200", "src"))

from oracle import segnet_oracle as O  # noqa: E402
from semiseg_b200 import synthetic  # noqa: E402

TINY_ARCH = O.Arch(num_leads=2, stem_channels=8, base_channels=8, head_channels=16, dropout_ratio=0.0)
TRAIN_CFG = {"epochs": 100, "accum_iter": 1, "warmup_epochs": 10, "min_lr": 0.0001, "blr": None, "lr": 0.001,
             "weight_decay": 0.05, "max_norm": None, "layer_decay": None, "optimizer": "adamw",
             "optimizer_kwargs": {"betas": [0.9, 0.999]}}


def group(g, prefix):
    pre = prefix + "/"
    return {k[len(pre):]: g[k] for k in g.files if k.startswith(pre)}


def sd_from(g, prefix):
    return {k: torch.from_numpy(np.asarray(v)) for k, v in group(g, prefix).items()}


def batches(seed0, n, Bl, Bu, C, L):
    out = []
    for i in range(n):
        a, b = synthetic.make_batch(seed0 + i, Bl, Bu, C, L)
        out.append(({k: torch.from_numpy(v) for k, v in a.items()}, {k: torch.from_numpy(v) for k, v in b.items()}))
    return out


def rel_err(a, b):
    a = torch.as_tensor(a).detach().double().flatten().cpu()
    b = torch.as_tensor(b).detach().double().flatten().cpu()
    return float((a - b).norm() / (b.norm() + 1e-30))


def model_cfg(num_leads, stem, base, head_ch, dropout):
    return {
        "backbone": {"resnet18": {"num_leads": num_leads, "num_stages": 4, "out_indices": [0, 1, 2, 3],
                                  "dilations": [1, 1, 1, 1], "strides": [1, 2, 2, 2], "deep_stem": False,
                                  "avg_down": False, "contract_dilation": False, "stem_channels": stem,
                                  "base_channels": base}},
        "decode_head": {"FCNHead": {"in_channels": base * 8, "in_index": 3, "channels": head_ch, "num_convs": 1,
                                    "concat_input": False, "dropout_ratio": dropout, "num_classes": 4,
                                    "align_corners": False}},
    }


def eval_batches(g):
    """the batches of golden case H (tests/golden/make_golden_eval.py): sizes 5, 5, 2"""
    out = []
    for i, n in enumerate(g["H/sizes"].tolist()):
        lab, _ = synthetic.make_batch(int(g["H/data_seed"]) + i, int(n), 1, 2, 300)
        out.append({k: torch.from_numpy(v) for k, v in lab.items()})
    return out
