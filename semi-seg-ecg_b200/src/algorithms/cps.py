"""Cross Pseudo Supervision trainer (reference src/algorithms/cps.py).

train_one_epoch(model_1, model_2, ..., optimizer_1, optimizer_2, ...) keeps the reference signature and returned keys
(cps.py:28-41,213-217: lr, loss_total, loss_x, loss_u_s -- the losses are the means of the two models').  One step
(cps.py:96-160): both models label the weak views in eval mode BEFORE either is updated; each then trains on
cat(ecg_x, ecg_u_w) against the labels and the OTHER model's argmax labels, CE over every position, (loss_x+loss_u)/2,
AdamW.  Runs as two hard-teacher StepEngines over the two models' arenas (semiseg_b200.engine.CpsEngine): two small
pseudo-label graphs, then the two training graphs side by side."""
from typing import Iterable, Optional

import torch

from algorithms.base import _setup, build_model_and_optimizer, evaluate, init_model_from_cfg, test, train_loop  # noqa: F401
from semiseg_b200.trainer import run_epoch_cps
from utils.semi_dataset import build_seg_dataset, get_dataloader


def train_one_epoch(model_1: torch.nn.Module, model_2: torch.nn.Module, labeled_data_loader: Iterable,
                    unlabeled_data_loader: Iterable, optimizer_1: torch.optim.Optimizer,
                    optimizer_2: torch.optim.Optimizer, device: torch.device, epoch: int, loss_scaler,
                    log_writer=None, use_amp=True, config: Optional[dict] = None):
    """Cross Pseudo Supervision (CPS) training; `config` is config['train']."""
    return run_epoch_cps(model_1, model_2, labeled_data_loader, unlabeled_data_loader, optimizer_1, optimizer_2,
                         device, epoch, loss_scaler, log_writer, use_amp, config)


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
    # two models drawn one after the other from the same RNG stream (cps.py:271-272): different initialisations
    model_1, optimizer_1, scaler = build_model_and_optimizer(config, device, seed)
    model_2, optimizer_2, _ = build_model_and_optimizer(config, device, seed)
    model_2.seed = seed + 1      # the two networks draw independent dropout masks (two nn.Dropout modules in the reference)

    def epoch_fn(epoch, log_writer, use_amp):
        return train_one_epoch(model_1, model_2, ld_l, ld_u, optimizer_1, optimizer_2, device, epoch, scaler,
                               log_writer, use_amp, config["train"])

    # validation and the checkpoints follow model_1 (cps.py:343-384)
    train_loop(config, epoch_fn, model_1, optimizer_1, scaler, {"train": [ld_l, ld_u], "valid": ld_v}, device)
