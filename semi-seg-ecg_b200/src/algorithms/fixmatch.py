"""FixMatch trainer (reference src/algorithms/fixmatch.py).

train_one_epoch keeps the reference signature and returned keys (fixmatch.py:28-39,188-192:
lr, loss_total, loss_x, loss_u_s, mask_ratio); the step itself -- self-eval pseudo-labels,
concatenated student forward, supervised CE + confidence-masked CE, backward, AdamW -- runs as
one CUDA graph of libsemiseg_b200 kernels (semiseg_b200.engine.StepEngine)."""
from typing import Iterable, Optional

import torch

from algorithms.base import _setup, build_model_and_optimizer, evaluate, init_model_from_cfg, test, train_loop  # noqa: F401
from semiseg_b200.trainer import run_epoch
from utils.semi_dataset import build_seg_dataset, get_dataloader


def train_one_epoch(model: torch.nn.Module, labeled_data_loader: Iterable, unlabeled_data_loader: Iterable,
                    optimizer: torch.optim.Optimizer, device: torch.device, epoch: int, loss_scaler,
                    log_writer=None, use_amp=True, config: Optional[dict] = None):
    """FixMatch training; `config` is config['train'] (needs conf_thresh and the LR schedule keys)."""
    return run_epoch("fixmatch", model, None, labeled_data_loader, unlabeled_data_loader, optimizer, device, epoch,
                     loss_scaler, log_writer, use_amp, config)


def train(config):
    device, seed = _setup(config)
    ds_u = build_seg_dataset(config["dataset"], split="train_unlabeled")
    ds_l = build_seg_dataset(config["dataset"], split="train_labeled", num_unlabeled=len(ds_u))
    ds_v = build_seg_dataset(config["dataset"], split="valid")
    dist_on = config["ddp"]["distributed"]
    ld_l = get_dataloader(ds_l, is_distributed=dist_on, mode="train", **config["dataloader"])
    ld_u = get_dataloader(ds_u, is_distributed=dist_on, mode="train", **config["dataloader"])
    ld_v = get_dataloader(ds_v, is_distributed=dist_on, mode="valid", **config["dataloader"])
    print(f"Labeled: {len(ds_l)} samples / {len(ld_l)} batches; Unlabeled: {len(ds_u)} samples / {len(ld_u)} batches")
    model, optimizer, scaler = build_model_and_optimizer(config, device, seed)

    def epoch_fn(epoch, log_writer, use_amp):
        return train_one_epoch(model, ld_l, ld_u, optimizer, device, epoch, scaler, log_writer, use_amp,
                               config["train"])

    train_loop(config, epoch_fn, model, optimizer, scaler, {"train": [ld_l, ld_u], "valid": ld_v}, device)
