"""Runtime helpers with the reference's call shapes (src/utils/misc.py), hot-path subset:
distributed init (209-233), loss-scaler facade (236-262), grad norm (265-278), checkpoint
dict format (281-321), scalar all-reduce (324-332), all-gather (335-350), meters (14-159)."""
import datetime
import math
import os
import time
from collections import defaultdict, deque

import torch
import torch.distributed as dist


# ---- distributed ---------------------------------------------------------------------
def is_dist_avail_and_initialized():
    return dist.is_available() and dist.is_initialized()


def get_world_size():
    return dist.get_world_size() if is_dist_avail_and_initialized() else 1


def get_rank():
    return dist.get_rank() if is_dist_avail_and_initialized() else 0


def is_main_process():
    return get_rank() == 0


def save_on_master(*args, **kwargs):
    if is_main_process():
        torch.save(*args, **kwargs)


def setup_for_distributed(is_master, with_time=True):
    """Rank-0-only print, optionally time-stamped (reference misc.py:162-177)."""
    import builtins
    if getattr(builtins.print, "_ssb_wrapped", False):
        return
    raw = builtins.print

    def rank0_print(*args, **kwargs):
        force = kwargs.pop("force", False)
        if is_master or force:
            if with_time:
                raw("[{}] ".format(datetime.datetime.now().time()), end="")
            raw(*args, **kwargs)

    rank0_print._ssb_wrapped = True
    builtins.print = rank0_print


def init_distributed_mode(config, with_time=True):
    """Same env contract as the reference: RANK / WORLD_SIZE / LOCAL_RANK from torchrun, NCCL
    backend, env:// rendezvous; one process per GPU."""
    if "RANK" in os.environ and "WORLD_SIZE" in os.environ:
        config["rank"] = int(os.environ["RANK"])
        config["world_size"] = int(os.environ["WORLD_SIZE"])
        config["gpu"] = int(os.environ.get("LOCAL_RANK", 0))
    elif "SLURM_PROCID" in os.environ:
        config["rank"] = int(os.environ["SLURM_PROCID"])
        config["gpu"] = config["rank"] % torch.cuda.device_count()
    else:
        print("Not using distributed mode")
        setup_for_distributed(is_master=True, with_time=with_time)
        config["distributed"] = False
        return
    config["distributed"] = True
    torch.cuda.set_device(config["gpu"])
    config["dist_backend"] = "nccl"
    print(f"| distributed init (rank {config['rank']}): {config.get('dist_url', 'env://')}, gpu {config['gpu']}", flush=True)
    if not dist.is_initialized():
        dist.init_process_group(backend="nccl", init_method=config.get("dist_url", "env://"),
                                world_size=config["world_size"], rank=config["rank"],
                                device_id=torch.device("cuda", config["gpu"]))
    dist.barrier()
    setup_for_distributed(config["rank"] == 0, with_time=with_time)


def all_reduce_mean(x):
    world_size = get_world_size()
    if world_size > 1:
        t = torch.tensor(x, device="cuda" if dist.get_backend() == "nccl" else "cpu")
        dist.all_reduce(t)
        return (t / world_size).item()
    return x


@torch.no_grad()
def concat_all_gather(tensor):
    world_size = get_world_size()
    if world_size == 1:
        return tensor
    out = [torch.ones_like(tensor) for _ in range(world_size)]
    dist.all_gather(out, tensor, async_op=False)
    return torch.cat(out, dim=0)


# ---- loss scaler facade ---------------------------------------------------------------
def get_grad_norm_(parameters, norm_type: float = 2.0) -> torch.Tensor:
    if isinstance(parameters, torch.Tensor):
        parameters = [parameters]
    grads = [p.grad.detach() for p in parameters if p.grad is not None]
    if not grads:
        return torch.tensor(0.0)
    if norm_type == math.inf:
        return max(g.abs().max() for g in grads)
    return torch.norm(torch.stack([torch.norm(g, norm_type) for g in grads]), norm_type)


class NativeScalerWithGradNormCount:
    """`loss_scaler(loss, optimizer, clip_grad, parameters, create_graph, update_grad)`.

    The B200 path computes in bf16 or fp32, neither of which needs dynamic loss scaling, so the
    scale is the constant 1.0 and no step is ever skipped; `state_dict()` keeps the GradScaler
    keys so checkpoints interchange (reference misc.py:236-262)."""
    state_dict_key = "amp_scaler"

    def __init__(self):
        self._state = {"scale": 1.0, "growth_factor": 2.0, "backoff_factor": 0.5, "growth_interval": 2000,
                       "_growth_tracker": 0}

    def __call__(self, loss, optimizer, clip_grad=None, parameters=None, create_graph=False, update_grad=True):
        loss.backward(create_graph=create_graph)
        norm = None
        if update_grad:
            if clip_grad is not None:
                assert parameters is not None
                norm = torch.nn.utils.clip_grad_norm_(parameters, clip_grad)
            elif parameters is not None:
                norm = get_grad_norm_(parameters)
            optimizer.step()
        return norm

    def state_dict(self):
        return dict(self._state)

    def load_state_dict(self, state_dict):
        self._state.update(state_dict or {})


# ---- checkpoints -----------------------------------------------------------------------
def save_model(config, checkpoint_path, epoch, model_without_ddp, optimizer=None, loss_scaler=None, metrics=None,
               model_ema=None):
    to_save = {
        "epoch": epoch,
        "model": model_without_ddp.state_dict(),
        "optimizer": optimizer.state_dict() if optimizer is not None else None,
        "scaler": loss_scaler.state_dict() if loss_scaler is not None else None,
        "config": config,
    }
    if metrics is not None:
        to_save["metrics"] = metrics
    if model_ema is not None:
        to_save["model_ema"] = model_ema.state_dict()
    save_on_master(to_save, checkpoint_path)


def load_model(config, model_without_ddp, optimizer, loss_scaler, model_ema=None):
    if not config.get("resume"):
        return
    checkpoint = torch.load(config["resume"], map_location="cpu", weights_only=False)
    model_without_ddp.load_state_dict(checkpoint["model"])
    if hasattr(model_without_ddp, "runtime") and next(model_without_ddp.parameters()).is_cuda:
        model_without_ddp.runtime().ensure()    # the flat arenas the fused optimizer's state lives in
    if model_ema is not None and "model_ema" in checkpoint:
        # NOTE: the reference loads model_ema into storage still aliased with the student and thereby
        # overwrites the student (SURVEY.md Appendix A, quirk ii); here the two models own separate arenas.
        if hasattr(model_ema, "runtime") and next(model_ema.parameters()).is_cuda:
            model_ema.runtime(nbt_float=True).ensure()   # float32 num_batches_tracked, as the reference's EMA leaves them
        model_ema.load_state_dict(checkpoint["model_ema"])
        # the restored teacher carries EMA history: the first step after the resume must CONTINUE the average
        # (trainer.get_engine: ema_first = False), not re-initialise the teacher from the student
        model_ema.ema_restored = True
    print("Resume checkpoint %s" % config["resume"])
    if "optimizer" in checkpoint and "epoch" in checkpoint and not config.get("eval"):
        optimizer.load_state_dict(checkpoint["optimizer"])
        config["start_epoch"] = checkpoint["epoch"] + 1
        if "scaler" in checkpoint and checkpoint["scaler"] is not None:
            loss_scaler.load_state_dict(checkpoint["scaler"])
        print("With optim & sched!")


# ---- meters ----------------------------------------------------------------------------
class SmoothedValue:
    def __init__(self, window_size=20, fmt=None):
        self.deque = deque(maxlen=window_size)
        self.total, self.count = 0.0, 0
        self.fmt = fmt or "{median:.4f} ({global_avg:.4f})"

    def update(self, value, n=1):
        self.deque.append(value)
        self.count += n
        self.total += value * n

    def synchronize_between_processes(self):
        if not is_dist_avail_and_initialized():
            return
        t = torch.tensor([self.count, self.total], dtype=torch.float64,
                         device="cuda" if dist.get_backend() == "nccl" else "cpu")
        dist.barrier()
        dist.all_reduce(t)
        self.count, self.total = int(t[0].item()), t[1].item()

    @property
    def median(self):
        return torch.tensor(list(self.deque)).median().item()

    @property
    def avg(self):
        return torch.tensor(list(self.deque), dtype=torch.float32).mean().item()

    @property
    def global_avg(self):
        return self.total / max(self.count, 1)

    @property
    def max(self):
        return max(self.deque)

    @property
    def value(self):
        return self.deque[-1]

    def __str__(self):
        return self.fmt.format(median=self.median, avg=self.avg, global_avg=self.global_avg, max=self.max,
                               value=self.value)


class MetricLogger:
    def __init__(self, delimiter="\t"):
        self.meters = defaultdict(SmoothedValue)
        self.delimiter = delimiter

    def update(self, **kwargs):
        for k, v in kwargs.items():
            if v is None:
                continue
            if isinstance(v, torch.Tensor):
                v = v.item()
            self.meters[k].update(float(v))

    def add_meter(self, name, meter):
        self.meters[name] = meter

    def synchronize_between_processes(self):
        for meter in self.meters.values():
            meter.synchronize_between_processes()

    def __str__(self):
        return self.delimiter.join(f"{name}: {meter}" for name, meter in self.meters.items())
