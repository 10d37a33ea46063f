"""Shimmed import of the UNMODIFIED reference (TEST INFRASTRUCTURE ONLY).

Used only in the build container (where /root/reference exists) by
tests/golden/make_golden.py and by the optional live cross-check in
tests/test_oracle_golden.py.  Never imported by the product path, never available on
the GPU box.  Recipe = SURVEY.md Appendix B: three out-of-tree shims
(`torch._six`, a `torchmetrics` stub, no-op `torch.cuda.synchronize` on CPU hosts).
No reference source is copied; the modules are imported from where they lie.
"""
from __future__ import annotations

import math
import os
import sys
import types

REF_CANDIDATES = [os.environ.get("SEMISEG_REF", ""), "/root/reference"]


def find_reference() -> str | None:
    for c in REF_CANDIDATES:
        if c and os.path.isdir(os.path.join(c, "src", "algorithms")):
            return c
    return None


def import_reference():
    """Returns a namespace with the reference's modules.  Must run in a process that
    has NOT imported this repo's own `models` / `algorithms` / `utils` packages (same
    top-level names)."""
    root = find_reference()
    if root is None:
        raise RuntimeError("reference tree not found (set SEMISEG_REF)")
    import torch

    if "torch._six" not in sys.modules:
        six = types.ModuleType("torch._six")
        six.inf = math.inf
        sys.modules["torch._six"] = six
    if "torchmetrics" not in sys.modules:
        tm = types.ModuleType("torchmetrics")
        tm.Metric = object
        tm.MetricCollection = dict
        seg = types.ModuleType("torchmetrics.segmentation")
        tm.segmentation = seg
        sys.modules["torchmetrics"] = tm
        sys.modules["torchmetrics.segmentation"] = seg
    if "mergedeep" not in sys.modules:
        md = types.ModuleType("mergedeep")
        md.merge = lambda a, *bs: a
        sys.modules["mergedeep"] = md
    if not torch.cuda.is_available():
        torch.cuda.synchronize = lambda *a, **k: None
    src = os.path.join(root, "src")
    if src not in sys.path:
        sys.path.insert(0, src)
    import algorithms.base as base
    import algorithms.fixmatch as fixmatch
    import algorithms.mean_teacher as mean_teacher
    import models.backbones as backbones
    import models.decode_heads as decode_heads
    import utils.lr_sched as lr_sched
    import utils.misc as misc
    import utils.optimizer as optimizer
    from models.encoder_decoder import EncoderDecoder

    return types.SimpleNamespace(
        root=root, base=base, fixmatch=fixmatch, mean_teacher=mean_teacher, backbones=backbones,
        decode_heads=decode_heads, lr_sched=lr_sched, misc=misc, optimizer=optimizer,
        EncoderDecoder=EncoderDecoder)


class ListLoader(list):
    """Sized iterable of batch dicts -- all `train_one_epoch` needs (fixmatch.py:58-71)."""
