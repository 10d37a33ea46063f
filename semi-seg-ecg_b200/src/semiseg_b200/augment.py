"""GPU-resident augmentation + batch assembly.

Replaces the per-item numpy/scipy work of the reference's DataLoader workers
(file:line relative to the reference tree):
  weak   RandomResizeCrop                 src/utils/transforms.py:93-127
  strong RandAugment(4 ops, 3 layers)     src/utils/transforms.py:340-351, 480-546, 574-583, 647-657
  Standardize + ToTensor                  src/utils/transforms.py:301-310, 603-617
  item order                              src/utils/semi_dataset.py:193-197, 235-242
  parameters                              configs/base/resnet18/fixmatch.yaml:56-83

The host only makes the scalar DRAWS (a handful per strip, in the reference's `np.random` call order, so a
seeded run consumes the same stream as the reference does for these scalars); everything that touches
samples runs on the device and lands directly in the step engine's static input arena.  The bulk noise
arrays (AmplitudeScaling factors, white noise) come from a counter-based device RNG unless explicit arrays
are injected (parity tests).  There is no CPU path: without the CUDA library this module raises.
"""
from __future__ import annotations

import ctypes as C
from dataclasses import dataclass, field
from typing import Dict, List, Optional, Sequence

import numpy as np
import torch

from . import _lib
from ._lib import AugOp, call

OP_KINDS = {"amplitude_scaling": _lib.AUG_AMPLITUDE, "powerline": _lib.AUG_POWERLINE,
            "partial_white": _lib.AUG_PARTIAL_WHITE, "partial_sine": _lib.AUG_PARTIAL_SINE}
OP_NAMES = ("amplitude_scaling", "powerline", "partial_white", "partial_sine")   # YAML op order


@dataclass
class AugConfig:
    """The augmentation block of the shipped YAML (fixmatch.yaml:56-83)."""
    target_length: int = 2500
    scale_min: float = 0.5
    scale_max: float = 2.0
    level: int = 10
    num_layers: int = 3
    prob: float = 0.5
    fs: int = 250

    @classmethod
    def from_config(cls, cfg: dict) -> "AugConfig":
        ds = cfg.get("dataset", cfg)
        out = cls(target_length=int(ds.get("signal_length", 2500)))
        for a in ds.get("augmentations", []) or []:
            if "random_resize_crop" in a:
                p = a["random_resize_crop"]
                out.target_length = int(p.get("target_length", out.target_length))
                out.scale_min, out.scale_max = float(p.get("scale_min", 0.5)), float(p.get("scale_max", 2.0))
        for a in ds.get("strong_augmentations", []) or []:
            if "RandAugment" in a:
                p = a["RandAugment"]
                out.level, out.num_layers, out.prob = int(p.get("level", 10)), int(p.get("num_layers", 2)), float(p.get("prob", 0.5))
                for op in p.get("ops", []):
                    if "AdaptivePowerlineNoise" in op:
                        out.fs = int(op["AdaptivePowerlineNoise"].get("fs", 500))
        return out


def draw_weak(L: int, cfg: AugConfig) -> Dict:
    """RandomResizeCrop's two draws (transforms.py:96-97, 120)."""
    ratio = np.random.uniform(cfg.scale_min, cfg.scale_max)
    size = int(L * ratio)
    start = int(np.random.randint(0, max(size, cfg.target_length) - cfg.target_length + 1))
    return {"size": size, "start": start}


def draw_strong(C_: int, L: int, cfg: AugConfig, bulk: bool = False) -> Dict:
    """RandAugment's draws (transforms.py:647-657 -> 575 -> per-op).  bulk=True also draws the per-sample
    noise arrays from numpy exactly where the reference does (parity mode); otherwise they are left to the
    device RNG and numpy's stream is NOT advanced for them."""
    lv = cfg.level / 10.0
    order = np.random.choice(len(OP_NAMES), cfg.num_layers, replace=False)
    ops = []
    for oi in order:
        name = OP_NAMES[int(oi)]
        d = {"op": name, "apply": bool(np.random.rand() < cfg.prob), "a": 0, "b": 0}
        if d["apply"]:
            if name == "amplitude_scaling":
                if bulk:
                    d["scales"] = np.random.normal(1, lv * 0.5, size=(C_, L))
            elif name == "powerline":
                d["a"] = 50 if np.random.rand() < 0.5 else 60
            else:
                if name == "partial_white" and bulk:
                    d["noise"] = np.random.randn(C_, L)
                d["a"] = int(np.random.uniform(0, lv * 0.5) * L)
                d["b"] = int(np.random.randint(0, L - d["a"]))
        ops.append(d)
    return {"ops": ops}


class GpuAugmenter:
    """Batched weak / strong / standardise on the device for strips of one shape [B, C, L]."""

    def __init__(self, cfg: AugConfig, B: int, num_leads: int, L: int, device, seed: int = 0):
        _lib.prepare()
        if torch.device(device).type != "cuda":
            raise RuntimeError("GpuAugmenter runs on CUDA only (no CPU fallback)")
        if L != cfg.target_length:
            raise ValueError(f"strip length {L} != target_length {cfg.target_length} (resample first, semi_dataset.py:183-186)")
        self.cfg, self.B, self.C, self.L, self.device, self.seed = cfg, B, num_leads, L, torch.device(device), seed
        self.spec = torch.zeros(B * num_leads, L // 2 + 1, 2, dtype=torch.float32, device=device)
        self.weak = torch.zeros(B, num_leads, L, dtype=torch.float32, device=device)
        self.size_d = torch.zeros(B, dtype=torch.int32, device=device)
        self.start_d = torch.zeros(B, dtype=torch.int32, device=device)
        self.ops_d = torch.zeros(B * max(cfg.num_layers, 1) * 4, dtype=torch.int32, device=device)
        self.size_h = torch.zeros(B, dtype=torch.int32).pin_memory()
        self.start_h = torch.zeros(B, dtype=torch.int32).pin_memory()
        self.ops_h = torch.zeros(B * max(cfg.num_layers, 1) * 4, dtype=torch.int32).pin_memory()
        self.calls = 0

    def _st(self) -> int:
        return torch.cuda.current_stream().cuda_stream

    # ---- weak ---------------------------------------------------------------------
    def weak_resize_crop(self, x: torch.Tensor, labels: Optional[torch.Tensor], draws: Sequence[Dict],
                         out: Optional[torch.Tensor] = None, labels_out: Optional[torch.Tensor] = None) -> torch.Tensor:
        """x [B, C, L] fp32 cuda, labels [B, L] int64 or None; draws: per strip {'size', 'start'}."""
        B, Cn, L = self.B, self.C, self.L
        assert x.is_cuda and x.dtype == torch.float32 and x.is_contiguous() and tuple(x.shape) == (B, Cn, L)
        out = self.weak if out is None else out
        for i, d in enumerate(draws):
            if not (1 <= d["size"] <= 2 * L):
                raise ValueError(f"resize target {d['size']} outside [1, 2L]")
            self.size_h[i], self.start_h[i] = d["size"], d["start"]
        self.size_d.copy_(self.size_h, non_blocking=True)
        self.start_d.copy_(self.start_h, non_blocking=True)
        st = self._st()
        call("ssb_aug_spectrum", x.data_ptr(), self.spec.data_ptr(), self.size_d.data_ptr(), B, Cn, L, st)
        if labels is not None:
            assert labels.dtype == torch.int64 and labels.is_contiguous() and tuple(labels.shape) == (B, L)
            assert labels_out is not None and labels_out.dtype == torch.int64 and labels_out.is_contiguous()
        call("ssb_aug_resize_crop", self.spec.data_ptr(), labels.data_ptr() if labels is not None else None, out.data_ptr(),
             labels_out.data_ptr() if labels is not None else None, self.size_d.data_ptr(), self.start_d.data_ptr(), B, Cn, L,
             int(max(d["size"] for d in draws)), st)
        return out

    # ---- strong + standardise -------------------------------------------------------
    def strong_standardize(self, x: torch.Tensor, out: torch.Tensor, draws: Optional[Sequence[Dict]] = None,
                           scales: Optional[torch.Tensor] = None, white: Optional[torch.Tensor] = None) -> torch.Tensor:
        """out = Standardize(RandAugment(x)) (draws given) or Standardize(x) (draws None)."""
        B, Cn, L = self.B, self.C, self.L
        assert x.is_cuda and x.dtype == torch.float32 and x.is_contiguous() and tuple(x.shape) == (B, Cn, L)
        assert out.is_cuda and out.dtype == torch.float32 and out.is_contiguous() and tuple(out.shape) == (B, Cn, L)
        n_ops = 0
        if draws is not None:
            n_ops = len(draws[0]["ops"])
            flat = self.ops_h.view(-1, 4)
            for i, d in enumerate(draws):
                for j, op in enumerate(d["ops"]):
                    flat[i * n_ops + j, 0] = OP_KINDS[op["op"]]
                    flat[i * n_ops + j, 1] = 1 if op["apply"] else 0
                    flat[i * n_ops + j, 2] = op.get("a", 0)
                    flat[i * n_ops + j, 3] = op.get("b", 0)
            self.ops_d.copy_(self.ops_h, non_blocking=True)
        self.calls += 1
        call("ssb_aug_strong_standardize", x.data_ptr(), out.data_ptr(), self.ops_d.data_ptr() if n_ops else None, n_ops,
             scales.data_ptr() if scales is not None else None, white.data_ptr() if white is not None else None,
             (self.seed * 1000003 + self.calls) & 0xFFFFFFFF, B, Cn, L, self.cfg.fs, self.cfg.level / 10.0, self._st())
        return out


class FixMatchBatcher:
    """Raw strips -> the step engine's input arena, entirely on the device:
       labeled    ecg = Standardize(weak(x)), target = weak(labels)
       unlabeled  ecg = Standardize(weak(x)),  ecg_aug = Standardize(RandAugment(weak(x)))
    (semi_dataset.py:234-242; the weak crop of an unlabeled strip is shared by both views)."""

    def __init__(self, engine, cfg: AugConfig, seed: int = 0):
        self.eng = engine
        dev, Cn, L = engine.device, engine.spec.num_leads, engine.L
        self.aug_l = GpuAugmenter(cfg, engine.Bl, Cn, L, dev, seed)
        self.aug_u = GpuAugmenter(cfg, engine.Bu, Cn, L, dev, seed + 1) if engine.Bu else None
        self.cfg = cfg

    def load(self, raw_l: torch.Tensor, lab_l: torch.Tensor, raw_u: Optional[torch.Tensor] = None) -> None:
        eng, cfg = self.eng, self.cfg
        Cn, L = eng.spec.num_leads, eng.L
        dl = [draw_weak(L, cfg) for _ in range(eng.Bl)]
        xw = self.aug_l.weak_resize_crop(raw_l, lab_l, dl, labels_out=eng.y_l)
        self.aug_l.strong_standardize(xw, eng.x_s[: eng.Bl])
        if self.aug_u is not None:
            du, ds = [], []
            for _ in range(eng.Bu):       # per item: weak draws, then strong draws (dataset order)
                du.append(draw_weak(L, cfg))
                ds.append(draw_strong(Cn, L, cfg))
            xw = self.aug_u.weak_resize_crop(raw_u, None, du)
            self.aug_u.strong_standardize(xw, eng.x_uw)
            self.aug_u.strong_standardize(xw, eng.x_s[eng.Bl:], ds)
