"""CPU restatement of the reference's evaluation tail -- TEST INFRASTRUCTURE ONLY (imported by tests/ and by
tests/golden/make_golden_eval.py, never by the product path).

  evaluate                      src/algorithms/base.py:184-245
  metric construction           src/utils/perf_metrics.py:9-47, configs/base/resnet18/fixmatch.yaml:110-119
  MeanIoU                       torchmetrics==1.5.2 (requirements.txt:12) -- a third-party dependency that is NOT in the
                                reference tree and not installable here (no network, not in the wheelhouse)

PINNING: loss, soft-max outputs, predictions and one-hot labels are pinned against the unmodified reference
`evaluate` run on CPU (tests/golden/eval_vectors.npz).  The MeanIoU AGGREGATION is "parity unpinned": it restates the
published torchmetrics 1.5.2 algorithm
  functional/segmentation/mean_iou.py  _mean_iou_update: intersection = sum(preds & target) over the positions of each
                                       sample and class, union = sum(target) + sum(preds) - intersection
                                       _mean_iou_compute: _safe_divide(intersection, union) (0 where union == 0), mean
                                       over the class axis unless per_class
  segmentation/mean_iou.py             update(): score += batch mean of that; num_batches += 1
                                       compute(): score / num_batches
and the golden generator feeds THIS class to the reference's evaluate as its `metric_fn`.
"""
from __future__ import annotations

from typing import Dict, Iterable, List

import numpy as np
import torch

from . import segnet_oracle as O


class MeanIoU:
    """torchmetrics.segmentation.MeanIoU(num_classes, include_background, per_class, input_format='one-hot') restated."""
    higher_is_better = True

    def __init__(self, num_classes: int, include_background: bool = True, per_class: bool = False,
                 input_format: str = "one-hot", **_unused):
        assert input_format == "one-hot"
        self.num_classes, self.include_background, self.per_class = num_classes, include_background, per_class
        self.reset()

    def reset(self):
        n = self.num_classes - (0 if self.include_background else 1)
        self.score = torch.zeros(n if self.per_class else 1, dtype=torch.float64)
        self.num_batches = 0

    def to(self, device):
        return self

    def update(self, preds: torch.Tensor, target: torch.Tensor):
        """preds, target: one-hot [N, C, L] integer tensors"""
        assert preds.shape == target.shape and preds.shape[1] == self.num_classes
        preds, target = preds.bool(), target.bool()
        if not self.include_background:
            preds, target = preds[:, 1:], target[:, 1:]
        axes = list(range(2, preds.ndim))
        inter = (preds & target).sum(dim=axes).double()
        union = target.sum(dim=axes).double() + preds.sum(dim=axes).double() - inter
        iou = torch.where(union > 0, inter / union.clamp(min=1.0), torch.zeros_like(inter))
        score = iou if self.per_class else iou.mean(dim=1)
        self.score += score.mean(dim=0) if self.per_class else score.mean()
        self.num_batches += 1

    def compute(self):
        out = self.score / self.num_batches
        return out if self.per_class else out.squeeze(0)


class MetricCollection(dict):
    """The two calls the reference makes on torchmetrics.MetricCollection (base.py:218, 228, 243)."""

    def __init__(self, metrics: List[MeanIoU]):
        super().__init__({m.__class__.__name__: m for m in metrics})

    def to(self, device):
        return self

    def update(self, preds, target):
        for m in self.values():
            m.update(preds, target)

    def compute(self):
        return {k: m.compute() for k, m in self.items()}

    def reset(self):
        for m in self.values():
            m.reset()


def evaluate(sd: Dict[str, torch.Tensor], arch: O.Arch, batches: Iterable[Dict[str, torch.Tensor]],
             include_background: bool = True, per_class: bool = False, dtype=torch.float64):
    """base.py:184-245 on the oracle's forward: returns (stats, metric_dict, outputs [N, C, L], preds [N, L])."""
    sdd = {k: (v.to(dtype) if v.is_floating_point() else v) for k, v in sd.items()}
    metric = MeanIoU(arch.num_classes, include_background, per_class)
    tot, cnt = 0.0, 0
    outs, preds = [], []
    for b in batches:
        x, y = b["ecg"], b["target"]
        with torch.no_grad():
            logits = O.forward(sdd, x.to(dtype), arch, False)["seg_logits"]
        loss = float(O.ce_hard(logits, y))
        prob = logits.softmax(dim=1)
        pred = prob.argmax(dim=1)
        oh = lambda t: torch.nn.functional.one_hot(t, num_classes=arch.num_classes).movedim(-1, 1)   # noqa: E731
        metric.update(oh(pred), oh(y))
        tot += loss * x.shape[0]
        cnt += x.shape[0]
        outs.append(prob)
        preds.append(pred)
    m = metric.compute()
    md = {f"MeanIoU_{i}": float(v) for i, v in enumerate(m.tolist())} if per_class else {"MeanIoU": float(m)}
    return {"loss": tot / cnt}, md, torch.cat(outs), torch.cat(preds)
