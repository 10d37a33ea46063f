"""Mean-Teacher trainer (reference src/algorithms/mean_teacher.py).

train_one_epoch(model_student, model_teacher, ...) keeps the reference signature and returned
keys (mean_teacher.py:28-40,192-196: lr, loss_total, loss_x, loss_u_s).  Teacher soft targets,
soft-target CE, AdamW and the EMA over parameters AND buffers (mean_teacher.py:139-149) are
kernels of the same captured step.  The teacher/student aliasing quirk of the reference's
teacher init (mean_teacher.py:285-290) is reproduced for step parity:
the first EMA sees k_old == q_new."""
from typing import Iterable, Optional

import torch

from algorithms.base import _setup, build_model_and_optimizer, evaluate, init_model_from_cfg, test, train_loop  # noqa: F401
from semiseg_b200.trainer import run_epoch
from utils.semi_dataset import build_seg_dataset, get_dataloader


def train_one_epoch(model_student: torch.nn.Module, model_teacher: torch.nn.Module, labeled_data_loader: Iterable,
                    unlabeled_data_loader: Iterable, optimizer: torch.optim.Optimizer, device: torch.device,
                    epoch: int, loss_scaler, log_writer=None, use_amp=True, config: Optional[dict] = None):
    """Mean Teacher training; `config` is config['train'] (ema_decay default 0.999)."""
    return run_epoch("mean_teacher", model_student, model_teacher, labeled_data_loader, unlabeled_data_loader,
                     optimizer, device, epoch, loss_scaler, log_writer, use_amp, config)


def init_teacher(config, student, device):
    """Teacher = second model from the same config, parameters taken from the student, buffers
    fresh, frozen (mean_teacher.py:281-290)."""
    teacher = init_model_from_cfg(config)
    teacher.to(device)
    for p in teacher.parameters():
        p.requires_grad = False
    with torch.no_grad():
        for q, k in zip(student.parameters(), teacher.parameters()):
            k.data.copy_(q.data)
    teacher.eval()
    return teacher


def train(config):
    device, seed = _setup(config)
    ds_u = build_seg_dataset(config["dataset"], split="train_unlabeled")
    ds_l = build_seg_dataset(config["dataset"], split="train_labeled", num_unlabeled=len(ds_u))
    ds_v = build_seg_dataset(config["dataset"], split="valid")
    dist_on = config["ddp"]["distributed"]
    ld_l = get_dataloader(ds_l, is_distributed=dist_on, mode="train", **config["dataloader"])
    ld_u = get_dataloader(ds_u, is_distributed=dist_on, mode="train", **config["dataloader"])
    ld_v = get_dataloader(ds_v, is_distributed=dist_on, mode="valid", **config["dataloader"])
    model, optimizer, scaler = build_model_and_optimizer(config, device, seed)
    teacher = init_teacher(config, model, device)

    def epoch_fn(epoch, log_writer, use_amp):
        return train_one_epoch(model, teacher, ld_l, ld_u, optimizer, device, epoch, scaler, log_writer, use_amp,
                               config["train"])

    train_loop(config, epoch_fn, model, optimizer, scaler, {"train": [ld_l, ld_u], "valid": ld_v}, device,
               model_ema=teacher)
