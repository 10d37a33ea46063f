"""ST++ self-training step (reference src/algorithms/stpp.py).

train_one_epoch(model_student, model_teacher, ...) keeps the reference signature and returned keys (stpp.py:91-103,
241-245: lr, loss_total, loss_x, loss_u_s).  One step (stpp.py:140-197): the FROZEN teacher labels the weak views in
eval mode (argmax), the student trains on cat(ecg_x, ecg_u_w) against the labels and those pseudo-labels, CE over every
position, (loss_x+loss_u)/2, AdamW -- the hard-teacher mode of semiseg_b200.engine.StepEngine, one CUDA graph.

calculate_miou / select_reliable (stpp.py:32-88) are mirrored for the reliability split.  The three-stage pipeline
around the step (stpp.py:248-760: supervised stage, checkpoint reload, reliable-subset loaders, two retraining stages)
is checkpoint / dataset orchestration, outside the accelerated hot path: `train` says so."""
from typing import Iterable, Optional

import numpy as np
import torch

from algorithms.base import evaluate, init_model_from_cfg, test  # noqa: F401
from semiseg_b200.trainer import run_epoch


def calculate_miou(onehot_preds, onehot_labels, ignore_background=False):
    """Mean over classes of |A & B| / |A | B| on one-hot [N, K, L] arrays; a class absent from both counts 0."""
    if ignore_background:
        onehot_preds, onehot_labels = onehot_preds[:, 1:], onehot_labels[:, 1:]
    ious = []
    for c in range(onehot_preds.shape[1]):
        inter = (onehot_preds[:, c] * onehot_labels[:, c]).sum()
        union = onehot_preds[:, c].sum() + onehot_labels[:, c].sum() - inter
        ious.append(inter / union if union > 0 else 0.0)
    return np.mean(ious)


@torch.no_grad()
def select_reliable(models, dataloader, device):
    """Rank unlabeled strips by the agreement (mIoU) of the earlier checkpoints' predictions with the last one's;
    returns (reliable ids = top half, unreliable ids = the rest)."""
    for m in models:
        m.eval()
    scored = []
    for i, data in enumerate(dataloader):
        ecg = data["ecg"].to(device, non_blocking=True)
        assert ecg.shape[0] == 1, "Batch size should be 1 for reliability estimation"
        onehots = []
        for m in models:
            logits = m(ecg, return_loss=False)["seg_logits"]
            pred = torch.argmax(logits, dim=1)
            onehots.append(torch.nn.functional.one_hot(pred, num_classes=logits.shape[1]).movedim(-1, 1).cpu().numpy())
        mious = [calculate_miou(onehots[j], onehots[-1]) for j in range(len(onehots) - 1)]
        scored.append((i, sum(mious) / len(mious)))
    scored.sort(key=lambda e: e[1], reverse=True)
    half = len(scored) // 2
    return [e[0] for e in scored[:half]], [e[0] for e in scored[half:]]


def train_one_epoch(model_student: torch.nn.Module, model_teacher: torch.nn.Module, labeled_data_loader: Iterable,
                    unlabeled_data_loader: Iterable, optimizer: torch.optim.Optimizer, device: torch.device,
                    epoch: int, loss_scaler, log_writer=None, use_amp=True, config: Optional[dict] = None):
    """Self-training with the frozen teacher's hard pseudo-labels; `config` is config['train']."""
    return run_epoch("stpp", model_student, model_teacher, labeled_data_loader, unlabeled_data_loader, optimizer,
                     device, epoch, loss_scaler, log_writer, use_amp, config)


def train(config):
    raise NotImplementedError(
        "stpp.train: the three-stage ST++ pipeline (supervised stage, checkpoint reload, reliability split, two "
        "retraining stages; reference stpp.py:248-760) is dataset / checkpoint orchestration outside the accelerated "
        "path.  Drive it with algorithms.base.train for stage 1, stpp.select_reliable for the split and "
        "stpp.train_one_epoch (the accelerated step) for stages 2 and 3.")
