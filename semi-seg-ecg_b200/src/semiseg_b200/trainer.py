"""Shared epoch loop behind algorithms.{base,fixmatch,mean_teacher}.train_one_epoch.

Keeps the reference's observable behaviour (per-iteration LR schedule, returned epoch means,
non-finite-loss abort, tensorboard scalar names) while replacing the per-step host work:
no `.item()` per step, no `cuda.synchronize()` per step, no separate logging all-reduces --
loss sums are read back asynchronously every `print_freq` steps.
"""
from __future__ import annotations

import math
import os
import sys
import time
from typing import Dict, Iterable, Optional

import torch
import torch.distributed as dist

from . import _lib
from .engine import CpsEngine, StepEngine

PRINT_FREQ = 20


def _unwrap(model):
    return model.module if hasattr(model, "module") and not hasattr(model, "runtime") else model


def _lr_at(epoch: float, cfg: dict) -> float:
    warm = cfg["warmup_epochs"]
    if epoch < warm:
        return cfg["lr"] * epoch / warm
    return cfg["min_lr"] + (cfg["lr"] - cfg["min_lr"]) * 0.5 * (
        1.0 + math.cos(math.pi * (epoch - warm) / (cfg["epochs"] - warm)))


def get_engine(algorithm: str, model, teacher, Bl: int, Bu: int, L: int, dtype: int, config: dict,
               optimizer=None, use_graph: bool = True, algo: Optional[int] = None,
               external_pseudo: bool = False) -> StepEngine:
    """Engine cache keyed by shapes; engines share the model's arenas."""
    rt = model.runtime()
    if not rt.quick_ok():        # (per-step path: an O(1) check; the full parameter walk only when something moved)
        rt.ensure()
    rt_t = None
    if algorithm == "mean_teacher":
        rt_t = teacher.runtime(nbt_float=True)
        if not rt_t.quick_ok():
            rt_t.ensure()
    elif algorithm in ("cps", "stpp"):      # the peer model / the frozen teacher: a plain second weight set
        rt_t = teacher.runtime()
        if not rt_t.quick_ok():
            rt_t.ensure()
    key = (algorithm, Bl, Bu, L, dtype, use_graph, algo, id(rt_t), bool(getattr(model, "sync_bn", False)), external_pseudo,
           config.get("max_norm", None))
    eng = rt.engines.get(key)
    if eng is None:
        pg = dist.group.WORLD if (dist.is_available() and dist.is_initialized() and
                                  (dist.get_world_size() > 1 or os.environ.get("SSB_FORCE_COLLECTIVES"))) else None
        cfg = dict(config)
        if optimizer is not None:
            g = optimizer.param_groups[0]
            cfg["optimizer"] = "adamw"
            cfg["weight_decay"] = g.get("weight_decay", cfg.get("weight_decay", 0.0))
            cfg["optimizer_kwargs"] = {"betas": tuple(g.get("betas", (0.9, 0.999))), "eps": g.get("eps", 1e-8)}
        eng = StepEngine(rt.weights, rt.state, dtype, algorithm, Bl, Bu, L, cfg,
                         teacher=rt_t.weights if rt_t is not None else None, algo=algo, use_graph=use_graph,
                         process_group=pg, sync_bn=bool(getattr(model, "sync_bn", False)),
                         seed=int(getattr(model, "seed", 0)), external_pseudo=external_pseudo)
        if rt_t is not None and algorithm == "mean_teacher":
            # fresh run: the teacher's parameters alias the student's until the first EMA (mean_teacher.py:281-290);
            # a teacher restored from a checkpoint (misc.load_model) or already averaged keeps its own history
            eng.ema_first = not (rt_t.ema_started or getattr(teacher, "ema_restored", False))
        rt.engines[key] = eng
    return eng


def bind_optimizer_state(optimizer, model) -> None:
    """Make `optimizer.state` (any torch AdamW-like optimizer over model.parameters()) alias the flat
    moment arenas, importing existing state first."""
    rt = model.runtime()
    rt.ensure()
    st = rt.state
    bound = getattr(optimizer, "_ssb_bound", None)
    if bound is not None and bound[0] is st and len(optimizer.state) == bound[1]:
        return      # this optimizer's state already aliases these arenas (every epoch after the first)
    mv, vv = rt.weights.param_views(st.exp_avg), rt.weights.param_views(st.exp_avg_sq)
    with torch.no_grad():
        for n, p in model.named_parameters():
            s = optimizer.state.get(p, None)
            if s and "exp_avg" in s and s["exp_avg"].data_ptr() != mv[n].data_ptr():
                mv[n].copy_(s["exp_avg"])
                vv[n].copy_(s["exp_avg_sq"])
                st.step = max(st.step, int(s["step"]))
            optimizer.state[p] = {"step": torch.tensor(float(st.step)), "exp_avg": mv[n], "exp_avg_sq": vv[n]}
    optimizer._ssb_bound = (st, len(optimizer.state))


def run_epoch(algorithm: str, model, teacher, labeled_loader: Iterable, unlabeled_loader: Optional[Iterable],
              optimizer, device, epoch: int, loss_scaler=None, log_writer=None, use_amp: bool = True,
              config: Optional[dict] = None) -> Dict[str, float]:
    model = _unwrap(model)
    teacher = _unwrap(teacher) if teacher is not None else None
    config = config or {}
    if config.get("accum_iter", 1) != 1:
        raise NotImplementedError("accum_iter > 1 is not supported by the fused step (all shipped configs use 1)")
    if torch.device(device).type != "cuda":
        raise RuntimeError("train_one_epoch: the B200 hot path needs device='cuda' (no CPU fallback)")
    model.train()
    if teacher is not None:
        teacher.eval()
    precision = getattr(model, "precision", None)
    dtype = {"fp32": _lib.F32, "bf16": _lib.BF16}[precision] if precision else (_lib.BF16 if use_amp else _lib.F32)
    num_steps = len(labeled_loader)
    if unlabeled_loader is not None:
        assert len(unlabeled_loader) == num_steps, "The number of labeled and unlabeled data should be the same"
    if log_writer is not None:
        print("log_dir: {}".format(log_writer.log_dir))
    bind_optimizer_state(optimizer, model)

    sums: Dict[str, float] = {}
    count = 0
    last_lr = 0.0
    lr_sum = 0.0
    t0 = time.time()
    eng = None
    pairs = zip(labeled_loader, unlabeled_loader) if unlabeled_loader is not None else ((b, None) for b in labeled_loader)

    def drain(step_idx, block=True):
        nonlocal count
        for s in eng.read_stats(block=block):
            total = s.get("loss_total", s.get("loss"))
            if not math.isfinite(total):
                print(f"Loss is {total}, stopping training")
                sys.exit(1)
            for k, v in s.items():
                sums[k] = sums.get(k, 0.0) + v
            count += 1
            if log_writer is not None:
                x = int((epoch + (count - 1) / num_steps) * 1000)
                for k, v in s.items():
                    log_writer.add_scalar(k, v, x)
        if log_writer is not None:
            log_writer.add_scalar("lr", last_lr, int((epoch + step_idx / num_steps) * 1000))

    for it, (lab, unl) in enumerate(pairs):
        lr = _lr_at(it / num_steps + epoch, config)
        for g in optimizer.param_groups:
            g["lr"] = lr * g["lr_scale"] if "lr_scale" in g else lr
        last_lr = lr
        lr_sum += lr
        ecg_x, mask_x = lab["ecg"], lab["target"]
        Bl, _, L = ecg_x.shape
        Bu = unl["ecg"].shape[0] if unl is not None else 0
        e = get_engine(algorithm, model, teacher, Bl, Bu, L, dtype, config, optimizer)
        if e is not eng and eng is not None:
            drain(it)
        eng = e
        if unl is not None:
            eng.load_batch(ecg_x, mask_x, unl["ecg"], unl.get("ecg_aug") if algorithm != "stpp" else None)
        else:
            eng.load_batch(ecg_x, mask_x)
        eng.step(lr)
        if teacher is not None and algorithm == "mean_teacher":
            teacher.runtime().ema_started = True
        if (it + 1) % PRINT_FREQ == 0 or it + 1 == num_steps:
            # mid-epoch log lines use the steps whose loss sums have ALREADY arrived (the reference blocks on .item() every
            # step; blocking here every PRINT_FREQ steps would still drain the GPU's queue once per line)
            drain(it, block=it + 1 == num_steps)
            dt = time.time() - t0
            print(f"Epoch: [{epoch}]  [{it + 1}/{num_steps}]  lr: {lr:.6f}  " +
                  "  ".join(f"{k}: {v / max(count, 1):.4f}" for k, v in sums.items()) +
                  f"  time: {dt / (it + 1):.4f}  max mem: {torch.cuda.max_memory_allocated() / 2 ** 20:.0f}")
    # (host-side bookkeeping first, the blocking read of the last steps' loss sums after it: the GPU is still working)
    step_t = torch.tensor(float(model.runtime().state.step))
    for p in optimizer.state.values():
        if isinstance(p, dict) and "step" in p:
            p["step"] = step_t
    if eng is not None:
        drain(num_steps - 1)
    stats = {k: v / max(count, 1) for k, v in sums.items()}
    stats["lr"] = lr_sum / max(num_steps, 1)      # the reference returns each meter's global average (fixmatch.py:188-192)
    if dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1:
        keys = sorted(k for k in stats if k != "lr")
        t = torch.tensor([stats[k] for k in keys], dtype=torch.float64, device=device)
        dist.all_reduce(t)
        t /= dist.get_world_size()
        for k, v in zip(keys, t.tolist()):
            stats[k] = v
    print("Averaged stats:", "  ".join(f"{k}: {v:.6f}" for k, v in stats.items()))
    return stats


def run_epoch_cps(model_1, model_2, labeled_loader: Iterable, unlabeled_loader: Iterable, optimizer_1, optimizer_2,
                  device, epoch: int, loss_scaler=None, log_writer=None, use_amp: bool = True,
                  config: Optional[dict] = None) -> Dict[str, float]:
    """Cross Pseudo Supervision epoch (reference cps.py:28-217): the same loop as run_epoch over a CpsEngine -- two
    models, two optimizers, one staged batch; returns lr, loss_total, loss_x, loss_u_s (means of the two models)."""
    model_1, model_2 = _unwrap(model_1), _unwrap(model_2)
    config = config or {}
    if config.get("accum_iter", 1) != 1:
        raise NotImplementedError("accum_iter > 1 is not supported by the fused step (all shipped configs use 1)")
    if torch.device(device).type != "cuda":
        raise RuntimeError("train_one_epoch: the B200 hot path needs device='cuda' (no CPU fallback)")
    model_1.train()
    model_2.train()
    precision = getattr(model_1, "precision", None)
    dtype = {"fp32": _lib.F32, "bf16": _lib.BF16}[precision] if precision else (_lib.BF16 if use_amp else _lib.F32)
    num_steps = len(unlabeled_loader)
    assert len(labeled_loader) == num_steps, "The number of labeled and unlabeled data should be the same"
    bind_optimizer_state(optimizer_1, model_1)
    bind_optimizer_state(optimizer_2, model_2)
    sums: Dict[str, float] = {}
    count = 0
    lr_sum = 0.0
    t0 = time.time()
    cps: Optional[CpsEngine] = None
    engines = {}

    def drain():
        nonlocal count
        for s in cps.read_stats():
            if not math.isfinite(s["loss_total"]):
                print(f"Loss is {s['loss_total']}, stopping training")
                sys.exit(1)
            for k, v in s.items():
                sums[k] = sums.get(k, 0.0) + v
            count += 1
            if log_writer is not None:
                x = int((epoch + (count - 1) / num_steps) * 1000)
                for k, v in s.items():
                    log_writer.add_scalar(k, v, x)

    for it, (lab, unl) in enumerate(zip(labeled_loader, unlabeled_loader)):
        lr = _lr_at(it / num_steps + epoch, config)
        for opt in (optimizer_1, optimizer_2):
            for g in opt.param_groups:
                g["lr"] = lr * g["lr_scale"] if "lr_scale" in g else lr
        lr_sum += lr
        ecg_x, mask_x, ecg_u_w = lab["ecg"], lab["target"], unl["ecg"]
        Bl, _, L = ecg_x.shape
        Bu = ecg_u_w.shape[0]
        key = (Bl, Bu, L)
        if key not in engines:
            e1 = get_engine("cps", model_1, model_2, Bl, Bu, L, dtype, config, optimizer_1, external_pseudo=True)
            e2 = get_engine("cps", model_2, model_1, Bl, Bu, L, dtype, config, optimizer_2, external_pseudo=True)
            engines[key] = CpsEngine(e1, e2)
        if cps is not None and engines[key] is not cps:
            drain()
        cps = engines[key]
        cps.load_batch(ecg_x, mask_x, ecg_u_w)
        cps.step(lr)
        if (it + 1) % PRINT_FREQ == 0 or it + 1 == num_steps:
            drain()
            print(f"Epoch: [{epoch}]  [{it + 1}/{num_steps}]  lr: {lr:.6f}  " +
                  "  ".join(f"{k}: {v / max(count, 1):.4f}" for k, v in sums.items()) +
                  f"  time: {(time.time() - t0) / (it + 1):.4f}")
    if cps is not None:
        drain()
    for opt, m in ((optimizer_1, model_1), (optimizer_2, model_2)):
        for p in opt.state.values():
            if isinstance(p, dict) and "step" in p:
                p["step"] = torch.tensor(float(m.runtime().state.step))
    stats = {k: v / max(count, 1) for k, v in sums.items()}
    stats["lr"] = lr_sum / max(num_steps, 1)
    if dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1:
        keys = sorted(k for k in stats if k != "lr")
        t = torch.tensor([stats[k] for k in keys], dtype=torch.float64, device=device)
        dist.all_reduce(t)
        t /= dist.get_world_size()
        for k, v in zip(keys, t.tolist()):
            stats[k] = v
    print("Averaged stats:", "  ".join(f"{k}: {v:.6f}" for k, v in stats.items()))
    return stats
