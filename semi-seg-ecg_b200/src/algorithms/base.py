"""Model factory + supervised trainer (reference src/algorithms/base.py:32-181, 184-245, 248-439)."""
import json
import os
import time
from typing import Iterable, Optional

import numpy as np
import torch
import torch.nn as nn

import models.backbones as backbones
import models.decode_heads as decode_heads
import utils.misc as misc
from models.encoder_decoder import EncoderDecoder
from semiseg_b200.trainer import run_epoch
from utils.misc import NativeScalerWithGradNormCount as NativeScaler
from utils.optimizer import get_optimizer_from_config
from utils.semi_dataset import build_seg_dataset, get_dataloader


def init_model_from_cfg(config, train=True):
    """YAML dict -> EncoderDecoder via the by-name registries (reference base.py:32-80).
    Construction order (backbone, then head) matches the reference so a seeded init is identical."""
    backbone_name, backbone_kwargs = list(config["backbone"].items())[0]
    assert backbone_name in backbones.__dict__, f"Unsupported model name: {backbone_name}"
    backbone = backbones.__dict__[backbone_name](**backbone_kwargs)
    decoder_name, decoder_kwargs = list(config["decode_head"].items())[0]
    assert decoder_name in decode_heads.__dict__, f"Unsupported decode head name: {decoder_name}"
    decoder = decode_heads.__dict__[decoder_name](**decoder_kwargs)
    if config.get("auxiliary_heads", None) and train:
        raise NotImplementedError("auxiliary_heads: unused by every shipped config; not on the hot path")
    model = EncoderDecoder(backbone=backbone, decode_head=decoder, decode_head_loss=nn.CrossEntropyLoss(),
                           use_latent_projection=config.get("use_latent_projection", False),
                           projection_in_dim=config.get("projection_in_dim", None),
                           projection_out_dim=config.get("projection_out_dim", None))
    prec = (config.get("train", {}) or {}).get("precision", None)
    if prec is not None:
        model.precision = prec
    return model


def train_one_epoch(model: torch.nn.Module, data_loader: Iterable, optimizer: torch.optim.Optimizer,
                    device: torch.device, epoch: int, loss_scaler, log_writer=None, use_amp=True,
                    config: Optional[dict] = None):
    """Supervised epoch (reference base.py:83-181).  Returns {'lr', 'loss'} epoch means."""
    return run_epoch("supervised", model, None, data_loader, None, optimizer, device, epoch, loss_scaler,
                     log_writer, use_amp, config)


@torch.no_grad()
def evaluate(model, data_loader, device, metric_fn=None, use_amp=True, return_outputs=False):
    """Validation loss + MeanIoU (reference base.py:184-245): eval forward and the whole metric tail run as one CUDA
    graph per batch, the metric (torchmetrics 1.5.2 MeanIoU semantics) is finalised on the device with a single
    read-back (semiseg_b200.evaluate).  `metric_fn`: None, or a dict-like with the reference's MeanIoU options
    (include_background, per_class).  The reference also returns every soft-max output and one-hot label on the
    CPU; that is opt-in here (return_outputs=True -- `test` uses it)."""
    from semiseg_b200.evaluate import evaluate_loader
    opts = metric_fn if isinstance(metric_fn, dict) else {}
    stats, metrics, outputs, labels = evaluate_loader(
        _unwrap_model(model), data_loader, device, use_amp=use_amp,
        include_background=bool(opts.get("include_background", True)), per_class=bool(opts.get("per_class", False)),
        want_outputs=return_outputs)
    print("* " + "  ".join(f"{k}: {v:.3f}" for k, v in metrics.items()) + f"  loss: {stats['loss']:.3f}")
    return stats, metrics, outputs, labels


def _unwrap_model(model):
    return model.module if hasattr(model, "module") and not hasattr(model, "runtime") else model


def _setup(config):
    misc.init_distributed_mode(config["ddp"])
    device = torch.device(config["device"])
    seed = config["seed"] + misc.get_rank()
    torch.manual_seed(seed)
    np.random.seed(seed)
    return device, seed


def build_model_and_optimizer(config, device, seed):
    model = init_model_from_cfg(config)
    if config.get("mode", "scratch") != "scratch":
        ckpt = torch.load(config["pretrained_backbone"], map_location="cpu", weights_only=False)
        print(f"Load backbone from {config['pretrained_backbone']}")
        msg = model.backbone.load_state_dict(ckpt["model"], strict=False)
        print(msg)
        if config["mode"] == "freeze_backbone":
            raise NotImplementedError("freeze_backbone is not supported by the fused step")
    model.to(device)
    model.sync_bn = bool(config["ddp"].get("distributed", False) and config["ddp"].get("sync_bn", True))
    model.seed = seed
    eff = config["dataloader"]["batch_size"] * config["train"]["accum_iter"] * misc.get_world_size()
    if config["train"]["lr"] is None:
        config["train"]["lr"] = config["train"]["blr"] * eff / 256
    print(f"actual lr: {config['train']['lr']}  effective batch size: {eff}")
    if config["train"].get("layer_decay", None):
        raise NotImplementedError("layer_decay is ViT-only (out of scope)")
    if misc.get_world_size() > 1:   # what DDP does at wrap time: rank 0's initial weights everywhere
        model.runtime().ensure()
        torch.distributed.broadcast(model.runtime().weights.params, src=0)
        torch.distributed.broadcast(model.runtime().weights.bufs, src=0)
    optimizer = get_optimizer_from_config(config["train"], model.parameters())
    return model, optimizer, NativeScaler()


def train_loop(config, algorithm_epoch_fn, model, optimizer, loss_scaler, loaders, device, model_ema=None):
    """Epoch loop, validation, best-checkpoint bookkeeping and log.txt (reference fixmatch.py:318-408)."""
    output_dir = None
    log_writer = None
    if misc.is_main_process() and config.get("output_dir"):
        output_dir = os.path.join(config["output_dir"], str(config.get("exp_name", "exp")))
        os.makedirs(output_dir, exist_ok=True)
        try:
            from torch.utils.tensorboard import SummaryWriter
            log_writer = SummaryWriter(log_dir=output_dir)
        except Exception:  # tensorboard is optional
            log_writer = None
    misc.load_model(config, model, optimizer, loss_scaler, model_ema)
    best_loss, best_miou = float("inf"), -1.0
    use_amp = config.get("use_amp", True)
    start = time.time()
    for epoch in range(config.get("start_epoch", 0), config["train"]["epochs"]):
        for ld in loaders["train"]:
            if hasattr(getattr(ld, "sampler", None), "set_epoch"):
                ld.sampler.set_epoch(epoch)
        train_stats = algorithm_epoch_fn(epoch, log_writer, use_amp)
        valid_stats, metrics, _, _ = evaluate(model, loaders["valid"], device, config.get("metric"), use_amp=use_amp)
        if "MeanIoU" not in metrics:          # per_class: true -> MeanIoU_<c>; the checkpoint criterion is their mean
            metrics = {**metrics, "MeanIoU": float(np.mean(list(metrics.values())))}
        if output_dir and valid_stats["loss"] < best_loss:
            best_loss = valid_stats["loss"]
            misc.save_model(config, os.path.join(output_dir, "best-loss.pth"), epoch, model, optimizer, loss_scaler,
                            metrics={"loss": best_loss, **metrics}, model_ema=model_ema)
        if output_dir and metrics["MeanIoU"] > best_miou:
            best_miou = metrics["MeanIoU"]
            misc.save_model(config, os.path.join(output_dir, "best-MeanIoU.pth"), epoch, model, optimizer, loss_scaler,
                            metrics={"loss": valid_stats["loss"], **metrics}, model_ema=model_ema)
        print(f"MeanIoU: {metrics['MeanIoU']:.3f}  Best MeanIoU: {best_miou:.3f}")
        if log_writer is not None:
            log_writer.add_scalar("perf/valid_loss", valid_stats["loss"], epoch)
            for k, v in metrics.items():
                log_writer.add_scalar(f"perf/{k}", v, epoch)
            log_writer.flush()
        if output_dir and misc.is_main_process():
            log_stats = {**{f"train_{k}": v for k, v in train_stats.items()},
                         **{f"valid_{k}": v for k, v in valid_stats.items()}, **metrics, "epoch": epoch}
            with open(os.path.join(output_dir, "log.txt"), mode="a", encoding="utf-8") as f:
                f.write(json.dumps(log_stats) + "\n")
    print(f"Training time {time.time() - start:.0f}s")
    if log_writer is not None:
        log_writer.close()


def train(config):
    """Supervised training entry (reference base.py:248-439)."""
    device, seed = _setup(config)
    ds_train = build_seg_dataset(config["dataset"], split="train_labeled")
    ds_valid = build_seg_dataset(config["dataset"], split="valid")
    dist_on = config["ddp"]["distributed"]
    ld_train = get_dataloader(ds_train, is_distributed=dist_on, mode="train", **config["dataloader"])
    ld_valid = get_dataloader(ds_valid, is_distributed=dist_on, mode="valid", **config["dataloader"])
    model, optimizer, scaler = build_model_and_optimizer(config, device, seed)

    def epoch_fn(epoch, log_writer, use_amp):
        return train_one_epoch(model, ld_train, optimizer, device, epoch, scaler, log_writer, use_amp, config["train"])

    train_loop(config, epoch_fn, model, optimizer, scaler, {"train": [ld_train], "valid": ld_valid}, device)


def test(config):
    """Evaluate a trained checkpoint on the test split (reference base.py:442-499): `test.model_path`, else
    `<output_dir>/<exp_name>/best-<test.target_metric>.pth`; the checkpoint must exist (a freshly initialised model is
    never scored); rank 0 only; writes test_metrics.csv, test_outputs.npy and test_labels.npy like the reference."""
    if not misc.is_main_process():
        return None
    device = torch.device(config["device"])
    output_dir = os.path.join(config["output_dir"], str(config.get("exp_name", "exp")))
    os.makedirs(output_dir, exist_ok=True)
    ds = build_seg_dataset(config["dataset"], split="test")
    ld = get_dataloader(ds, is_distributed=False, mode="test", **config["dataloader"])
    model = init_model_from_cfg(config, train=False)
    tcfg = config.get("test") or {}
    if tcfg.get("model_path", None):
        path = tcfg["model_path"]
    else:
        path = os.path.join(output_dir, f"best-{tcfg.get('target_metric', 'loss')}.pth")
    assert os.path.exists(path), f"Checkpoint not found: {path}"
    state_dict = torch.load(path, map_location="cpu", weights_only=False)["model"]
    for k in list(state_dict.keys()):          # drop the auxiliary head
        if k.startswith("auxiliary_head"):
            del state_dict[k]
    print(model.load_state_dict(state_dict))
    model.to(device)
    stats, metrics, outputs, labels = evaluate(model, ld, device, config.get("metric"), use_amp=config.get("use_amp", True),
                                               return_outputs=True)
    metrics = dict(metrics, loss=stats["loss"])
    import csv
    with open(os.path.join(output_dir, "test_metrics.csv"), "w", newline="") as f:
        w = csv.writer(f)
        w.writerow(list(metrics.keys()))
        w.writerow([f"{float(v):.4f}" for v in metrics.values()])
    if outputs is not None:
        np.save(os.path.join(output_dir, "test_outputs.npy"), outputs.cpu().numpy())
        np.save(os.path.join(output_dir, "test_labels.npy"), labels.cpu().numpy())
    print("Done!")
    return stats, metrics
