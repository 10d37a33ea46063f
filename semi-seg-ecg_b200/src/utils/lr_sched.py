"""Per-iteration LR schedule (reference src/utils/lr_sched.py:6-18): linear warm-up over
`warmup_epochs`, then half-cosine down to `min_lr`; honours a per-group `lr_scale`."""
import math


def lr_at(epoch: float, config: dict) -> float:
    warm = config["warmup_epochs"]
    if epoch < warm:
        return config["lr"] * epoch / warm
    span = config["epochs"] - warm
    return config["min_lr"] + (config["lr"] - config["min_lr"]) * 0.5 * (1.0 + math.cos(math.pi * (epoch - warm) / span))


def adjust_learning_rate(optimizer, epoch, config):
    lr = lr_at(epoch, config)
    for group in optimizer.param_groups:
        group["lr"] = lr * group["lr_scale"] if "lr_scale" in group else lr
    return lr
