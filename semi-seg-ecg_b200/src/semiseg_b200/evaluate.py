"""Evaluation path: replaces the body of the reference's `evaluate`
(src/algorithms/base.py:184-245) -- eval-mode forward, soft-max, arg-max, one-hot encoding of predictions and labels,
torchmetrics MeanIoU on the CPU, a `.item()` per batch -- by one CUDA graph per batch shape:

  eval forward (BatchNorm folded, with residual add and ReLU, into the conv epilogues: ssb_conv1d_bn_act_fwd)
  -> ssb_eval_metrics (upsample + softmax + argmax + CE sum + per-sample class intersections / marginals)

Nothing is read back until the loader is exhausted; the metric is finalised on the device from the count arrays.

MeanIoU follows torchmetrics 1.5.2 (requirements.txt:12; the package is not installable in this environment, so its
published algorithm is restated -- torchmetrics/functional/segmentation/mean_iou.py and segmentation/mean_iou.py):
per sample and class IoU = |P & T| / (|P| + |T| - |P & T|), 0 where the union is empty; mean over classes (without class
0 when include_background is false); update() adds the BATCH mean of those to `score` and 1 to `num_batches`;
compute() = score / num_batches.  Under DDP the reference gathers every rank's batch first (misc.concat_all_gather),
so a "batch" is the global batch."""
from __future__ import annotations

from typing import Dict, List, Optional, Tuple

import torch
import torch.distributed as dist

from . import _lib
from ._lib import call
from .net import NetPlan, WeightSet


def mean_iou_from_counts(counts: torch.Tensor, include_background: bool = True, per_class: bool = False) -> torch.Tensor:
    """counts [N, ncls, 3] = (intersection, |pred|, |target|) per sample and class -> per-sample score
    ([N] or [N, C'] when per_class), torchmetrics 1.5.2 `_mean_iou_compute` semantics."""
    c = counts.double()
    if not include_background:
        c = c[:, 1:]
    inter = c[..., 0]
    union = c[..., 1] + c[..., 2] - inter
    iou = torch.where(union > 0, inter / union.clamp(min=1.0), torch.zeros_like(inter))
    return iou if per_class else iou.mean(dim=1)


def aggregate_eval(per_batch: List[Tuple[torch.Tensor, torch.Tensor, int]], include_background: bool = True,
                   per_class: bool = False, group=None) -> Tuple[Dict[str, float], Dict[str, float]]:
    """Finalise an evaluation from its per-batch results [(sums fp64[2] = {CE sum, labelled positions}, counts
    int32 [n, ncls, 3], n)], on whatever device they live, with ONE read-back.
      loss    = sum_b loss_b * n_b / sum_b n_b                  (metric_logger.meters['loss'].update(loss, n), base.py:219,
                                                                 then synchronize_between_processes: totals over ranks)
      MeanIoU = mean over batches of the batch mean of the per-sample scores (torchmetrics update()/compute())
    Under torch.distributed every rank holds its shard of each batch; the reference gathers the global batch before
    each update (base.py:207-217), i.e. the batch mean runs over all ranks' samples: per-batch score sums and sample
    counts are all-reduced (one collective for the whole evaluation), then divided."""
    dev = per_batch[0][0].device
    world = dist.get_world_size(group) if dist.is_available() and dist.is_initialized() else 1
    loss_w = torch.stack([s_[0] / s_[1].clamp(min=1.0) * n for s_, _, n in per_batch]).sum()        # sum_b loss_b * n_b
    n_tot = torch.tensor(float(sum(n for _, _, n in per_batch)), dtype=torch.float64, device=dev)
    score_sum = torch.stack([mean_iou_from_counts(c, include_background, per_class).sum(dim=0) for _, c, _ in per_batch])
    n_b = torch.tensor([float(n) for _, _, n in per_batch], dtype=torch.float64, device=dev)
    if world > 1:
        pack = torch.cat([loss_w.reshape(1), n_tot.reshape(1), n_b, score_sum.reshape(-1)])
        dist.all_reduce(pack, group=group)
        nb = len(per_batch)
        loss_w, n_tot, n_b = pack[0], pack[1], pack[2:2 + nb]
        score_sum = pack[2 + nb:].reshape(score_sum.shape)
    score = (score_sum / n_b.reshape(-1, *([1] * (score_sum.dim() - 1)))).mean(dim=0)              # mean over batches
    loss = float(loss_w / n_tot)
    if per_class:
        metrics = {f"MeanIoU_{i}": float(v) for i, v in enumerate(score.tolist())}
    else:
        metrics = {"MeanIoU": float(score)}
    return {"loss": loss}, metrics


class EvalEngine:
    """Static buffers + one captured graph for batches of one shape."""

    def __init__(self, weights: WeightSet, dtype: int, B: int, L: int, want_outputs: bool = False,
                 use_graph: bool = True):
        if weights.device.type != "cuda":
            raise RuntimeError("EvalEngine needs CUDA tensors: the hot path has no CPU fallback")
        _lib.check(_lib.load().ssb_device_check(), "ssb_device_check")
        self.w, self.dtype, self.B, self.L = weights, dtype, B, L
        self.spec = weights.layout.spec
        dev = weights.device
        self.plan = NetPlan(weights, dtype, B, L, False, None)
        ncls = self.spec.num_classes
        self.x = torch.zeros(B, self.spec.num_leads, L, dtype=torch.float32, device=dev)
        self.y = torch.zeros(B, L, dtype=torch.int64, device=dev)
        self.sums = torch.zeros(2, dtype=torch.float64, device=dev)
        self.counts = torch.zeros(B, ncls, 3, dtype=torch.int32, device=dev)
        self.probs = torch.zeros(B, ncls, L, dtype=torch.float32, device=dev) if want_outputs else None
        self.pred = torch.zeros(B, L, dtype=torch.int64, device=dev) if want_outputs else None
        self.use_graph = use_graph
        self.graph: Optional[torch.cuda.CUDAGraph] = None
        self.launches = 0

    def refresh_weights(self) -> None:
        """storage-dtype copy of the (possibly just trained) master weights; once per evaluation, not per batch"""
        self.plan.sh.refresh(torch.cuda.current_stream().cuda_stream)

    def _enqueue(self) -> None:
        st = torch.cuda.current_stream().cuda_stream
        n0 = _lib.load().ssb_launch_count()
        call("ssb_memset_zero", self.sums.data_ptr(), 16, st)
        call("ssb_memset_zero", self.counts.data_ptr(), self.counts.numel() * 4, st)
        low = self.plan.forward(self.x, st, train_mode=False)
        call("ssb_eval_metrics", low.data_ptr(), self.y.data_ptr(), self.sums.data_ptr(), self.counts.data_ptr(),
             self.probs.data_ptr() if self.probs is not None else None,
             self.pred.data_ptr() if self.pred is not None else None, self.B, self.plan.Lh, self.L,
             self.spec.num_classes, 1 if self.spec.align_corners else 0, st)
        self.launches = int(_lib.load().ssb_launch_count() - n0)

    def run(self, ecg: torch.Tensor, target: Optional[torch.Tensor] = None) -> None:
        """One batch (host or device tensors): afterwards self.sums / self.counts (/ probs / pred) hold its results
        on the device, in stream order.  target None (inference): the loss / count outputs refer to the stale labels
        and are meaningless."""
        self.x.copy_(ecg, non_blocking=True)
        if target is not None:
            self.y.copy_(target, non_blocking=True)
        if self.use_graph:
            if self.graph is None:
                g = torch.cuda.CUDAGraph()
                with torch.cuda.graph(g, capture_error_mode="thread_local"):
                    self._enqueue()
                self.graph = g
            self.graph.replay()
        else:
            self._enqueue()


def evaluate_loader(model, data_loader, device, use_amp: bool = True, include_background: bool = True,
                    per_class: bool = False, want_outputs: bool = False
                    ) -> Tuple[Dict[str, float], Dict[str, float], Optional[torch.Tensor], Optional[torch.Tensor]]:
    """The reference's evaluate() contract: ({'loss': sample-weighted mean of the batch losses}, {'MeanIoU': ...} (or
    MeanIoU_<c> per class), outputs, labels).  outputs / labels (soft-max probabilities [N, ncls, L] and one-hot labels
    [N, ncls, L], on the CPU, as the reference returns them) only when want_outputs."""
    if torch.device(device).type != "cuda":
        raise RuntimeError("evaluate: the B200 path needs device='cuda' (no CPU fallback)")
    model.eval()
    rt = model.runtime()
    rt.ensure()
    precision = getattr(model, "precision", None)
    dtype = {"fp32": _lib.F32, "bf16": _lib.BF16}[precision] if precision else (_lib.BF16 if use_amp else _lib.F32)
    engines = rt.__dict__.setdefault("eval_engines", {})
    ncls = model.decode_head.num_classes
    per_batch: List[Tuple[torch.Tensor, torch.Tensor, int]] = []
    outs, labs = [], []
    fresh = set()
    for samples in data_loader:
        ecg, target = samples["ecg"], samples["target"]
        B, _, L = ecg.shape
        key = (dtype, B, L, want_outputs)
        if key not in engines:
            engines[key] = EvalEngine(rt.weights, dtype, B, L, want_outputs=want_outputs)
        eng = engines[key]
        if key not in fresh:
            eng.refresh_weights()
            fresh.add(key)
        eng.run(ecg, target)
        per_batch.append((eng.sums.clone(), eng.counts.clone(), B))
        if want_outputs:
            outs.append(eng.probs.to("cpu", non_blocking=False))
            labs.append(torch.nn.functional.one_hot(eng.y, num_classes=ncls).movedim(-1, 1).to("cpu"))
    if not per_batch:
        return {"loss": float("nan")}, {"MeanIoU": float("nan")}, None, None
    stats, metrics = aggregate_eval(per_batch, include_background, per_class)
    outputs = torch.cat(outs, dim=0) if want_outputs else None
    labels = torch.cat(labs, dim=0) if want_outputs else None
    return stats, metrics, outputs, labels


@torch.no_grad()
def predict_loader(model, data_loader, device, use_amp: bool = False) -> torch.Tensor:
    """The loop of the reference's inference script (src/inference.py:108-119): soft-max outputs [N, ncls, L] of every
    batch of the loader, on the CPU.  Same graph as evaluation (the probabilities are an output of ssb_eval_metrics);
    device-to-host copies go through a pinned double buffer and overlap the next batch."""
    if torch.device(device).type != "cuda":
        raise RuntimeError("inference: the B200 path needs device='cuda' (no CPU fallback)")
    model.eval()
    rt = model.runtime()
    rt.ensure()
    precision = getattr(model, "precision", None)
    dtype = {"fp32": _lib.F32, "bf16": _lib.BF16}[precision] if precision else (_lib.BF16 if use_amp else _lib.F32)
    engines = rt.__dict__.setdefault("eval_engines", {})
    fresh = set()
    outs: List[torch.Tensor] = []
    pending: List[Tuple[torch.Tensor, torch.cuda.Event]] = []
    for samples in data_loader:
        ecg = samples["ecg"]
        B, _, L = ecg.shape
        key = (dtype, B, L, True)
        if key not in engines:
            engines[key] = EvalEngine(rt.weights, dtype, B, L, want_outputs=True)
        eng = engines[key]
        if key not in fresh:
            eng.refresh_weights()
            fresh.add(key)
        eng.run(ecg, None)
        host = torch.empty(eng.probs.shape, dtype=torch.float32, pin_memory=True)
        host.copy_(eng.probs, non_blocking=True)
        ev = torch.cuda.Event()
        ev.record()
        pending.append((host, ev))
        if len(pending) > 1:              # the previous batch's copy has had a whole batch of compute to finish
            h, e = pending.pop(0)
            e.synchronize()
            outs.append(h)
        # the next run() overwrites eng.probs: stream order keeps the copy above ahead of it
    for h, e in pending:
        e.synchronize()
        outs.append(h)
    return torch.cat(outs, dim=0) if outs else torch.empty(0)
