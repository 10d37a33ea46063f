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


def weak_table(draws: Sequence[Dict]):
    """per-strip draw dicts -> (sizes, starts) int32 arrays"""
    return (np.asarray([d["size"] for d in draws], dtype=np.int32), np.asarray([d["start"] for d in draws], dtype=np.int32))


def ops_table(draws: Sequence[Dict]) -> np.ndarray:
    """per-strip RandAugment draw dicts -> int32 [B, n_ops, 4] rows (kind, apply, a, b)"""
    return np.asarray([[[OP_KINDS[op["op"]], 1 if op["apply"] else 0, int(op.get("a", 0)), int(op.get("b", 0))]
                        for op in d["ops"]] for d in draws], dtype=np.int32)


def draw_batch(rng: np.random.Generator, B: int, L: int, cfg: AugConfig, strong: bool):
    """Vectorised draws for a whole batch (production path): the same distributions as draw_weak /
    draw_strong, one numpy call per quantity instead of ~8 Python-level RNG calls per strip."""
    ratio = rng.uniform(cfg.scale_min, cfg.scale_max, B)
    sizes = (L * ratio).astype(np.int64)
    starts = rng.integers(0, np.maximum(sizes, cfg.target_length) - cfg.target_length + 1)
    if not strong:
        return sizes.astype(np.int32), starts.astype(np.int32), None
    lv = cfg.level / 10.0
    n = cfg.num_layers
    kinds = rng.permuted(np.tile(np.arange(len(OP_NAMES)), (B, 1)), axis=1)[:, :n]
    apply = rng.random((B, n)) < cfg.prob
    freq = np.where(rng.random((B, n)) < 0.5, 50, 60)
    count = (rng.uniform(0, lv * 0.5, (B, n)) * L).astype(np.int64)
    start = rng.integers(0, L - count)
    partial = (kinds == _lib.AUG_PARTIAL_WHITE) | (kinds == _lib.AUG_PARTIAL_SINE)
    a = np.where(kinds == _lib.AUG_POWERLINE, freq, np.where(partial, count, 0))
    b = np.where(partial, start, 0)
    ops = np.stack([kinds, apply.astype(np.int64), a, b], axis=2).astype(np.int32)
    return sizes.astype(np.int32), starts.astype(np.int32), ops


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
        # pinned staging ring for the per-call draw tables: a slot is rewritten by the host only after the
        # asynchronous H2D copy that read it has completed (calls are enqueued back to back without syncs)
        self.ring = [{"size": torch.zeros(B, dtype=torch.int32).pin_memory(),
                      "start": torch.zeros(B, dtype=torch.int32).pin_memory(),
                      "ops": torch.zeros(B * max(cfg.num_layers, 1) * 4, dtype=torch.int32).pin_memory(),
                      "event": torch.cuda.Event(), "used": False} for _ in range(8)]
        for r in self.ring:
            r["size_np"], r["start_np"], r["ops_np"] = r["size"].numpy(), r["start"].numpy(), r["ops"].numpy()
        self.ring_pos = 0
        self.calls = 0

    def _slot(self) -> Dict:
        s = self.ring[self.ring_pos % len(self.ring)]
        self.ring_pos += 1
        if s["used"]:
            s["event"].synchronize()
        s["used"] = True
        return s

    def _st(self) -> int:
        return torch.cuda.current_stream().cuda_stream

    # ---- weak ---------------------------------------------------------------------
    def weak_resize_crop(self, x: torch.Tensor, labels: Optional[torch.Tensor], sizes: np.ndarray, starts: np.ndarray,
                         out: Optional[torch.Tensor] = None, labels_out: Optional[torch.Tensor] = None) -> torch.Tensor:
        """x [B, C, L] fp32 cuda, labels [B, L] int64 or None; sizes / starts: the two draws per strip."""
        B, Cn, L = self.B, self.C, self.L
        assert x.is_cuda and x.dtype == torch.float32 and x.is_contiguous() and tuple(x.shape) == (B, Cn, L)
        out = self.weak if out is None else out
        sizes, starts = np.asarray(sizes), np.asarray(starts)
        if sizes.shape != (B,) or starts.shape != (B,):
            raise ValueError("one (size, start) draw per strip is required")
        if sizes.min() < 1 or sizes.max() > 2 * L or starts.min() < 0 or (starts > np.maximum(sizes, L) - L).any():
            raise ValueError("resize draws outside their range (size in [1, 2L], start in [0, max(size, L) - L])")
        slot = self._slot()
        slot["size_np"][:] = sizes
        slot["start_np"][:] = starts
        self.size_d.copy_(slot["size"], non_blocking=True)
        self.start_d.copy_(slot["start"], non_blocking=True)
        slot["event"].record()
        st = self._st()
        call("ssb_aug_spectrum", x.data_ptr(), self.spec.data_ptr(), self.size_d.data_ptr(), B, Cn, L, st)
        if labels is not None:
            assert labels.dtype == torch.int64 and labels.is_contiguous() and tuple(labels.shape) == (B, L)
            assert labels_out is not None and labels_out.dtype == torch.int64 and labels_out.is_contiguous()
        call("ssb_aug_resize_crop", self.spec.data_ptr(), labels.data_ptr() if labels is not None else None, out.data_ptr(),
             labels_out.data_ptr() if labels is not None else None, self.size_d.data_ptr(), self.start_d.data_ptr(), B, Cn, L,
             int(sizes.max()), st)
        return out

    # ---- strong + standardise -------------------------------------------------------
    def strong_standardize(self, x: torch.Tensor, out: torch.Tensor, ops: Optional[np.ndarray] = None,
                           scales: Optional[torch.Tensor] = None, white: Optional[torch.Tensor] = None) -> torch.Tensor:
        """out = Standardize(RandAugment(x)) (ops: int32 [B, n_ops, 4] rows kind/apply/a/b) or Standardize(x) (ops None)."""
        B, Cn, L = self.B, self.C, self.L
        assert x.is_cuda and x.dtype == torch.float32 and x.is_contiguous() and tuple(x.shape) == (B, Cn, L)
        assert out.is_cuda and out.dtype == torch.float32 and out.is_contiguous() and tuple(out.shape) == (B, Cn, L)
        n_ops = 0
        if ops is not None:
            ops = np.asarray(ops, dtype=np.int32)
            n_ops = ops.shape[1]
            if ops.shape != (B, n_ops, 4) or n_ops > max(self.cfg.num_layers, 1):
                raise ValueError(f"ops table must be [B, <= {self.cfg.num_layers}, 4], got {ops.shape}")
            part = (ops[..., 1] != 0) & ((ops[..., 0] == _lib.AUG_PARTIAL_WHITE) | (ops[..., 0] == _lib.AUG_PARTIAL_SINE))
            if (ops[..., 0].min() < 0 or ops[..., 0].max() > 3 or
                    (part & ((ops[..., 2] < 0) | (ops[..., 3] < 0) | (ops[..., 2] + ops[..., 3] > L))).any()):
                raise ValueError("RandAugment draws outside their range")
            slot = self._slot()
            slot["ops_np"][: ops.size] = ops.reshape(-1)
            self.ops_d.copy_(slot["ops"], non_blocking=True)
            slot["event"].record()
        self.calls += 1
        call("ssb_aug_strong_standardize", x.data_ptr(), out.data_ptr(), self.ops_d.data_ptr() if n_ops else None, n_ops,
             scales.data_ptr() if scales is not None else None, white.data_ptr() if white is not None else None,
             (self.seed * 1000003 + self.calls) & 0xFFFFFFFF, None, B, Cn, L, self.cfg.fs, self.cfg.level / 10.0, self._st())
        return out


class FixMatchBatcher:
    """Raw strips -> the step engine's input arena, entirely on the device:
       labeled    ecg = Standardize(weak(x)), target = weak(labels)
       unlabeled  ecg = Standardize(weak(x)),  ecg_aug = Standardize(RandAugment(weak(x)))
    (semi_dataset.py:234-242; the weak crop of an unlabeled strip is shared by both views)."""

    def __init__(self, engine, cfg: AugConfig, seed: int = 0, exact_stream: bool = False):
        """exact_stream=True: per-strip draws on numpy's GLOBAL RNG in the reference's call order (a seeded run
        consumes the stream like the reference's dataset does for these scalars); False (default): vectorised
        draws from a private Generator -- same distributions, ~50x less host time per batch."""
        self.eng = engine
        dev, Cn, L = engine.device, engine.spec.num_leads, engine.L
        self.aug_l = GpuAugmenter(cfg, engine.Bl, Cn, L, dev, seed)
        self.aug_u = GpuAugmenter(cfg, engine.Bu, Cn, L, dev, seed + 1) if engine.Bu else None
        self.cfg = cfg
        self.exact_stream = exact_stream
        self.rng = np.random.Generator(np.random.PCG64(seed))

    def load(self, raw_l: torch.Tensor, lab_l: torch.Tensor, raw_u: Optional[torch.Tensor] = None) -> None:
        """Augment one batch on the current stream straight into the engine's input arena."""
        eng = self.eng
        self._augment_into(raw_l, lab_l, raw_u, eng.x_s[: eng.Bl], eng.y_l, eng.x_uw, eng.x_s[eng.Bl:])

    def _augment_into(self, raw_l, lab_l, raw_u, ecg_x, target, ecg_u_w, ecg_u_s) -> None:
        eng, cfg = self.eng, self.cfg
        Cn, L = eng.spec.num_leads, eng.L
        if self.exact_stream:
            sz, st = weak_table([draw_weak(L, cfg) for _ in range(eng.Bl)])
        else:
            sz, st, _ = draw_batch(self.rng, eng.Bl, L, cfg, False)
        xw = self.aug_l.weak_resize_crop(raw_l, lab_l, sz, st, labels_out=target)
        self.aug_l.strong_standardize(xw, ecg_x)
        if self.aug_u is not None:
            if self.exact_stream:
                du, ds = [], []
                for _ in range(eng.Bu):       # per item: weak draws, then strong draws (dataset order)
                    du.append(draw_weak(L, cfg))
                    ds.append(draw_strong(Cn, L, cfg))
                (sz, st), ops = weak_table(du), ops_table(ds)
            else:
                sz, st, ops = draw_batch(self.rng, eng.Bu, L, cfg, True)
            xw = self.aug_u.weak_resize_crop(raw_u, None, sz, st)
            self.aug_u.strong_standardize(xw, ecg_u_w)
            self.aug_u.strong_standardize(xw, ecg_u_s, ops)

    # ---- overlapped mode: augment batch i+1 on a side stream while step i runs ----------------
    def prefetch(self, raw_l: torch.Tensor, lab_l: torch.Tensor, raw_u: Optional[torch.Tensor] = None,
                 inputs_ready: Optional[torch.cuda.Event] = None) -> None:
        """Enqueue the augmentation of the NEXT batch on the augmentation stream into a staging set
        (two sets, used alternately); `commit()` hands the oldest prefetched batch to the engine.
        inputs_ready: event after which the raw strips may be read (host tensors, or device tensors that
        are already complete, need none).  The augmentation stream deliberately does NOT wait for the
        caller's stream: that would serialise it behind the training step it is meant to overlap."""
        eng = self.eng
        if not hasattr(self, "stream"):
            dev = eng.device
            self.stream = torch.cuda.Stream(device=dev)
            self.stage = [{"x": torch.empty_like(eng.x_s), "y": torch.empty_like(eng.y_l), "uw": torch.empty_like(eng.x_uw),
                           "ready": torch.cuda.Event(), "free": None} for _ in range(2)]
            self.queue: List[int] = []
            self.next = 0
        k = self.next
        self.next ^= 1
        sg = self.stage[k]
        if inputs_ready is not None:
            self.stream.wait_event(inputs_ready)
        if sg["free"] is not None:
            self.stream.wait_event(sg["free"])                   # the engine has copied this set out
        with torch.cuda.stream(self.stream):
            self._augment_into(raw_l, lab_l, raw_u, sg["x"][: eng.Bl], sg["y"], sg["uw"], sg["x"][eng.Bl:])
            sg["ready"].record()
        self.queue.append(k)

    def commit(self) -> None:
        """Make the oldest prefetched batch the engine's current batch (three device-to-device copies)."""
        eng = self.eng
        sg = self.stage[self.queue.pop(0)]
        cur = torch.cuda.current_stream()
        cur.wait_event(sg["ready"])
        eng.x_s.copy_(sg["x"], non_blocking=True)
        eng.y_l.copy_(sg["y"], non_blocking=True)
        if self.aug_u is not None:
            eng.x_uw.copy_(sg["uw"], non_blocking=True)
        sg["free"] = torch.cuda.Event()
        sg["free"].record()

    # ---- captured mode: the whole augmentation of a batch as ONE CUDA-graph launch ------------------
    def _build_graphs(self) -> None:
        eng, cfg = self.eng, self.cfg
        dev, Cn, L, Bl, Bu, n = eng.device, eng.spec.num_leads, eng.L, eng.Bl, eng.Bu, max(self.cfg.num_layers, 1)
        self.raw_l = torch.zeros(Bl, Cn, L, dtype=torch.float32, device=dev)
        self.lab_l = torch.zeros(Bl, L, dtype=torch.int64, device=dev)
        self.raw_u = torch.zeros(max(Bu, 1), Cn, L, dtype=torch.float32, device=dev)
        # draw table layout (int32): sizes_l | starts_l | sizes_u | starts_u | ops_u [Bu, n, 4] | seed
        self.t_off = {"sl": 0, "tl": Bl, "su": 2 * Bl, "tu": 2 * Bl + Bu, "ops": 2 * Bl + 2 * Bu, "seed": 2 * Bl + 2 * Bu + Bu * n * 4}
        nt = self.t_off["seed"] + 1
        self.tab_d = torch.zeros(nt, dtype=torch.int32, device=dev)
        self.tab_h = [torch.zeros(nt, dtype=torch.int32).pin_memory() for _ in range(2)]
        self.tab_np = [t.numpy() for t in self.tab_h]
        self.tab_ev = [None, None]
        self.stage = [{"x": torch.empty_like(eng.x_s), "y": torch.empty_like(eng.y_l), "uw": torch.empty_like(eng.x_uw),
                       "ready": torch.cuda.Event(), "free": None} for _ in range(2)]
        self.stream = torch.cuda.Stream(device=dev)
        self.graphs = []
        self.queue = []
        self.next = 0
        self.replays = 0
        base, o = self.tab_d.data_ptr(), self.t_off
        al, au = self.aug_l, self.aug_u
        for k in range(2):
            sg = self.stage[k]
            g = torch.cuda.CUDAGraph()
            with torch.cuda.graph(g, stream=self.stream, capture_error_mode="thread_local"):
                st = self.stream.cuda_stream
                self.tab_d.copy_(self.tab_h[k], non_blocking=True)
                call("ssb_aug_spectrum", self.raw_l.data_ptr(), al.spec.data_ptr(), base + 4 * o["sl"], Bl, Cn, L, st)
                call("ssb_aug_resize_crop", al.spec.data_ptr(), self.lab_l.data_ptr(), al.weak.data_ptr(), sg["y"].data_ptr(),
                     base + 4 * o["sl"], base + 4 * o["tl"], Bl, Cn, L, 2 * L, st)
                call("ssb_aug_strong_standardize", al.weak.data_ptr(), sg["x"].data_ptr(), None, 0, None, None, 0, None,
                     Bl, Cn, L, cfg.fs, cfg.level / 10.0, st)
                if au is not None:
                    call("ssb_aug_spectrum", self.raw_u.data_ptr(), au.spec.data_ptr(), base + 4 * o["su"], Bu, Cn, L, st)
                    call("ssb_aug_resize_crop", au.spec.data_ptr(), None, au.weak.data_ptr(), None, base + 4 * o["su"],
                         base + 4 * o["tu"], Bu, Cn, L, 2 * L, st)
                    call("ssb_aug_strong_standardize", au.weak.data_ptr(), sg["uw"].data_ptr(), None, 0, None, None, 0, None,
                         Bu, Cn, L, cfg.fs, cfg.level / 10.0, st)
                    call("ssb_aug_strong_standardize", au.weak.data_ptr(), sg["x"][Bl:].data_ptr(), base + 4 * o["ops"],
                         self.cfg.num_layers, None, None, (self.aug_u.seed * 1000003) & 0xFFFFFFFF, base + 4 * o["seed"],
                         Bu, Cn, L, cfg.fs, cfg.level / 10.0, st)
            self.graphs.append(g)

    def prefetch_captured(self, raw_l: torch.Tensor, lab_l: torch.Tensor, raw_u: Optional[torch.Tensor] = None,
                          inputs_ready: Optional[torch.cuda.Event] = None) -> None:
        """Like prefetch(), with the device work of one batch replayed as a single captured graph: the host
        draws the scalars (vectorised), fills one pinned table and launches once."""
        eng, cfg, L = self.eng, self.cfg, self.eng.L
        if not hasattr(self, "graphs"):
            self._build_graphs()
        k = self.next
        self.next ^= 1
        sg, o, tab = self.stage[k], self.t_off, self.tab_np[k]
        if self.tab_ev[k] is not None:
            self.tab_ev[k].synchronize()      # the previous replay of graph k has read its pinned table
        sz, st, _ = draw_batch(self.rng, eng.Bl, L, cfg, False)
        tab[o["sl"]: o["sl"] + eng.Bl], tab[o["tl"]: o["tl"] + eng.Bl] = sz, st
        if self.aug_u is not None:
            sz, st, ops = draw_batch(self.rng, eng.Bu, L, cfg, True)
            tab[o["su"]: o["su"] + eng.Bu], tab[o["tu"]: o["tu"] + eng.Bu] = sz, st
            tab[o["ops"]: o["ops"] + ops.size] = ops.reshape(-1)
        self.replays += 1
        tab[o["seed"]] = self.replays & 0x7FFFFFFF
        if inputs_ready is not None:
            self.stream.wait_event(inputs_ready)
        if sg["free"] is not None:
            self.stream.wait_event(sg["free"])       # the engine has copied this staging set out
        with torch.cuda.stream(self.stream):
            # (copies into the static raw buffers are ordered behind the previous replay on this stream)
            self.raw_l.copy_(raw_l, non_blocking=True)
            self.lab_l.copy_(lab_l, non_blocking=True)
            if self.aug_u is not None:
                self.raw_u.copy_(raw_u, non_blocking=True)
            self.graphs[k].replay()
            sg["ready"].record()
            self.tab_ev[k] = torch.cuda.Event()
            self.tab_ev[k].record()
        self.queue.append(k)
