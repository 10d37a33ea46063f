"""Optimizer factory (reference src/utils/optimizer.py:8-37).  'adamw' returns `FusedAdamW`, a
`torch.optim.Optimizer` whose moments live in the model's flat arenas and whose step is ONE
multi-tensor kernel (ssb_adamw_ema); 'sgd' is not on the accelerated path."""
from typing import Dict, Iterable

import torch

from semiseg_b200.optim import FusedAdamW


def get_optimizer_from_config(config: dict, param_groups) -> torch.optim.Optimizer:
    name = config["optimizer"]
    kwargs = config.get("optimizer_kwargs", {}) or {}
    if name == "adamw":
        betas = kwargs.get("betas", (0.9, 0.999))
        return FusedAdamW(param_groups, lr=config["lr"], betas=tuple(betas), eps=kwargs.get("eps", 1e-8),
                          weight_decay=config["weight_decay"])
    if name == "sgd":
        raise NotImplementedError("optimizer 'sgd' is not part of the accelerated hot path (every shipped config "
                                  "uses adamw); use torch.optim.SGD with the module-level API if needed")
    raise ValueError(f"Unknown optimizer: {name}")
