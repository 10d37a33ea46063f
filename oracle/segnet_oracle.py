"""CPU oracle: restatement of the SemiSegECG training hot path (TEST INFRASTRUCTURE ONLY).

Every function cites the reference location (relative to /root/reference) whose
behaviour it restates.  Nothing here is copied from the reference; the model is
expressed functionally over a ``state_dict``-shaped mapping (same key names as the
reference) so that the same tensors can be fed to the
reference, to this oracle and to the CUDA path.

Arithmetic: plain PyTorch on whatever device/dtype the inputs live on (CPU fp32 /
fp64 for parity truth; the pseudo-label helpers are also run on CUDA tensors by the
bit-exactness tests so that the comparison target is torch's own CUDA soft-max).
Third-party arithmetic: ``torch.nn.functional.conv1d`` (PyTorch; reference pins
torch==1.11.0+cu113 in requirements.txt:8, this image has torch 2.11.0).

Parity pinning: pinned against outputs of the reference itself, see
tests/golden/make_golden.py and tests/test_oracle_golden.py.
"""
from __future__ import annotations

import math
from dataclasses import dataclass, field
from typing import Dict, List, Optional, Sequence, Tuple

import torch
import torch.nn.functional as F

Tensor = torch.Tensor


# --------------------------------------------------------------------------------------
# architecture description
# --------------------------------------------------------------------------------------
@dataclass
class Arch:
    """Shape parameters of EncoderDecoder(resnet18-style BasicBlock net, FCNHead).

    Mirrors the constructor arguments in src/models/backbones/resnet.py:135-204 and
    src/models/decode_heads/fcn_head.py:10-24 that the hot path uses.
    """
    num_leads: int = 1
    stem_channels: int = 64
    base_channels: int = 64
    strides: Tuple[int, ...] = (1, 2, 2, 2)
    stage_blocks: Tuple[int, ...] = (2, 2, 2, 2)
    head_channels: int = 128
    num_classes: int = 4
    in_index: int = 3
    dropout_ratio: float = 0.1
    align_corners: bool = False
    bottleneck: bool = False      # Bottleneck blocks (resnet.py:75-132): 1x1 - 3x3(stride) - 1x1(x4), expansion 4

    @property
    def expansion(self) -> int:
        return 4 if self.bottleneck else 1

    @staticmethod
    def from_config(cfg: dict) -> "Arch":
        """YAML dict -> Arch (algorithms/base.py:32-43 reads the same two sub-dicts)."""
        bname, bkw = list(cfg["backbone"].items())[0]
        hname, hkw = list(cfg["decode_head"].items())[0]
        blocks = {"resnet18": (2, 2, 2, 2), "resnet34": (3, 4, 6, 3), "resnet50": (3, 4, 6, 3), "resnet101": (3, 4, 23, 3),
                  "resnet152": (3, 8, 36, 3)}[bname]
        ns = bkw.get("num_stages", 4)
        return Arch(
            num_leads=bkw["num_leads"],
            stem_channels=bkw.get("stem_channels", 64),
            base_channels=bkw.get("base_channels", 64),
            strides=tuple(bkw.get("strides", (1, 2, 2, 2)))[:ns],
            stage_blocks=tuple(blocks)[:ns],
            head_channels=hkw["channels"],
            num_classes=hkw["num_classes"],
            in_index=hkw.get("in_index", -1),
            dropout_ratio=hkw.get("dropout_ratio", 0.1),
            align_corners=hkw.get("align_corners", False),
            bottleneck=bname in ("resnet50", "resnet101", "resnet152"),
        )

    def planes(self, i: int) -> int:
        return self.base_channels * 2 ** i


def conv_out_len(L: int, k: int, s: int, p: int, d: int = 1) -> int:
    """L_out = floor((L + 2p - d(k-1) - 1)/s) + 1 (torch Conv1d/MaxPool1d length rule)."""
    return (L + 2 * p - d * (k - 1) - 1) // s + 1


def stage_lengths(arch: Arch, L: int) -> List[int]:
    """[stem, pooled, stage1..stageN] lengths; resnet.py:246-257 (k7 s2 p3, pool k3 s2 p1)."""
    out = [conv_out_len(L, 7, 2, 3)]
    out.append(conv_out_len(out[-1], 3, 2, 1))
    cur = out[-1]
    for s in arch.strides:
        cur = conv_out_len(cur, 3, s, 1)
        out.append(cur)
    return out


def param_names(arch: Arch) -> List[str]:
    """Parameter names in ``model.parameters()`` order (registration order of
    resnet.py:181-199 / BasicBlock 31-51 / fcn_head.py:70-83)."""
    names = ["backbone.stem.0.weight", "backbone.stem.1.weight", "backbone.stem.1.bias"]
    inpl = arch.stem_channels
    for i, nb in enumerate(arch.stage_blocks):
        pl = arch.planes(i)
        for j in range(nb):
            pre = f"backbone.layer{i + 1}.{j}"
            names += [f"{pre}.conv1.weight", f"{pre}.bn1.weight", f"{pre}.bn1.bias",
                      f"{pre}.conv2.weight", f"{pre}.bn2.weight", f"{pre}.bn2.bias"]
            if arch.bottleneck:
                names += [f"{pre}.conv3.weight", f"{pre}.bn3.weight", f"{pre}.bn3.bias"]
            if j == 0 and (arch.strides[i] != 1 or inpl != pl * arch.expansion):
                names += [f"{pre}.downsample.0.weight", f"{pre}.downsample.1.weight",
                          f"{pre}.downsample.1.bias"]
        inpl = pl * arch.expansion
    names += ["decode_head.convs.0.0.weight", "decode_head.convs.0.1.weight",
              "decode_head.convs.0.1.bias", "decode_head.cls_seg.weight", "decode_head.cls_seg.bias"]
    return names


def bn_prefixes(arch: Arch) -> List[str]:
    """BatchNorm module prefixes in ``model.buffers()`` order."""
    out = []
    for n in param_names(arch):
        if n.endswith(".bias") and not n.endswith("cls_seg.bias"):
            out.append(n[: -len(".bias")])
    return out


def buffer_names(arch: Arch) -> List[str]:
    out = []
    for p in bn_prefixes(arch):
        out += [p + ".running_mean", p + ".running_var", p + ".num_batches_tracked"]
    return out


# --------------------------------------------------------------------------------------
# layer restatements
# --------------------------------------------------------------------------------------
# FAST_KERNELS: the timing arm of bench.py (cpu_baseline / --impl reference) sets this so that BatchNorm, max-pool,
# upsampling and cross-entropy go through the same ATen kernels the reference's nn.Modules call (F.batch_norm,
# F.max_pool1d, F.interpolate, F.cross_entropy) instead of the explicit formulas below -- otherwise the CPU arm is
# ~1.5x slower than the reference itself.  tests/test_oracle_golden.py holds both modes to the
# golden vectors; the explicit formulas stay the default because they also run in fp64 with storage emulation.
FAST_KERNELS = False


def batchnorm(x: Tensor, sd: Dict[str, Tensor], pre: str, train: bool,
              new_buffers: Optional[Dict[str, Tensor]], eps: float = 1e-5, momentum: float = 0.1) -> Tensor:
    """nn.BatchNorm1d as used at resnet.py:41,50,254,290 and fcn_head.py:48.

    train: biased batch variance for normalisation, unbiased for the running estimate,
    running <- 0.9*running + 0.1*batch, num_batches_tracked += 1.  eval: running stats.
    """
    g, b = sd[pre + ".weight"], sd[pre + ".bias"]
    if FAST_KERNELS:
        rm, rv = sd[pre + ".running_mean"].to(x.dtype), sd[pre + ".running_var"].to(x.dtype)
        if not train:
            return F.batch_norm(x, rm, rv, g, b, False, momentum, eps)
        rm, rv = rm.detach().clone(), rv.detach().clone()
        y = F.batch_norm(x, rm, rv, g, b, True, momentum, eps)
        if new_buffers is not None:
            new_buffers[pre + ".running_mean"] = rm
            new_buffers[pre + ".running_var"] = rv
            new_buffers[pre + ".num_batches_tracked"] = sd[pre + ".num_batches_tracked"] + 1
        return y
    if train:
        n = x.shape[0] * x.shape[2]
        mean = x.mean(dim=(0, 2))
        var = ((x - mean[None, :, None]) ** 2).mean(dim=(0, 2))
        if new_buffers is not None:
            with torch.no_grad():
                rm, rv = sd[pre + ".running_mean"], sd[pre + ".running_var"]
                new_buffers[pre + ".running_mean"] = (1 - momentum) * rm + momentum * mean.detach().to(rm.dtype)
                new_buffers[pre + ".running_var"] = (1 - momentum) * rv + momentum * (var.detach() * (n / (n - 1))).to(rv.dtype)
                new_buffers[pre + ".num_batches_tracked"] = sd[pre + ".num_batches_tracked"] + 1
    else:
        mean = sd[pre + ".running_mean"].to(x.dtype)
        var = sd[pre + ".running_var"].to(x.dtype)
    invstd = 1.0 / torch.sqrt(var + eps)
    return (x - mean[None, :, None]) * (invstd * g)[None, :, None] + b[None, :, None]


def maxpool_k3s2p1(x: Tensor) -> Tensor:
    """nn.MaxPool1d(3, 2, 1) (resnet.py:257): -inf padding, window {2t-1, 2t, 2t+1}."""
    if FAST_KERNELS:
        return F.max_pool1d(x, 3, 2, 1)
    xp = F.pad(x, (1, 1), value=float("-inf"))
    return xp.unfold(2, 3, 2).max(dim=-1).values


def linear_upsample(x: Tensor, L_out: int, align_corners: bool = False) -> Tensor:
    """F.interpolate(mode='linear') as called at encoder_decoder.py:102-107.

    align_corners=False: src = max((t+0.5)*L_in/L_out - 0.5, 0); two-tap lerp.
    Index arithmetic is done in the tensor's dtype (ATen uses float for float tensors).
    """
    if FAST_KERNELS:
        return F.interpolate(x, size=L_out, mode="linear", align_corners=align_corners)
    L_in = x.shape[-1]
    t = torch.arange(L_out, dtype=x.dtype, device=x.device)
    if align_corners:
        scale = (L_in - 1) / (L_out - 1) if L_out > 1 else 0.0
        src = t * torch.tensor(scale, dtype=x.dtype, device=x.device)
    else:
        scale = torch.tensor(L_in, dtype=x.dtype, device=x.device) / L_out
        src = torch.clamp((t + 0.5) * scale - 0.5, min=0)
    i0 = torch.floor(src).long()
    i1 = torch.clamp(i0 + 1, max=L_in - 1)
    w1 = src - i0.to(x.dtype)
    w0 = 1.0 - w1
    return x[..., i0] * w0 + x[..., i1] * w1


def bf16_round(t: Tensor) -> Tensor:
    """Round to bf16 and back (straight-through for autograd): emulates bf16 STORAGE of a tensor."""
    return t + (t.detach().to(torch.bfloat16).to(t.dtype) - t.detach())


def forward(sd: Dict[str, Tensor], x: Tensor, arch: Arch, train: bool,
            dropout_mask: Optional[Tensor] = None,
            new_buffers: Optional[Dict[str, Tensor]] = None,
            taps: Optional[Dict[str, Tensor]] = None, quant=None,
            relu_masks: Optional[Dict[str, Tensor]] = None) -> Dict[str, Tensor]:
    """EncoderDecoder.forward (encoder_decoder.py:78-108) over resnet.py:353-363
    (stem -> maxpool -> BasicBlocks, resnet.py:55-72) and FCNHead.forward (fcn_head.py:89-97).

    x: [B, C, L].  dropout_mask: [B, head_channels, L_head] of {0,1} (keep mask) applied as
    ``h * mask / (1-p)`` in train mode; None means no dropout (p=0 or eval).
    Returns {'seg_logits': [B, ncls, L], 'low_logits': [B, ncls, L_head]}.
    If ``taps`` is a dict it is filled with named intermediate activations (module-output
    granularity) for per-layer parity checks.
    ``quant`` (e.g. ``bf16_round``) is applied to every conv weight, conv output and stored
    activation: it emulates the BF16 path's storage roundings on top of exact arithmetic, so that
    the BF16 kernels can be checked tightly (kernel correctness) separately from the precision
    loss that bf16 storage itself causes.
    ``relu_masks`` {tap name of a ReLU output: bool tensor} INJECTS the sign decisions of those ReLUs (like a dropout
    mask): relu(t) becomes t * mask.  ReLU'(0) is discontinuous, so two correct implementations whose pre-activations
    differ by rounding (1e-7) can take different decisions on an element that is ~0, which moves the gradients by
    O(1e-3); with the decisions injected, both sides differentiate the same piecewise-linear function.  (The stem's
    max-pool(relu(x)) is evaluated as relu(max-pool(x)) -- the same function -- so that its decisions live on the pooled
    tensor, the one the CUDA path materialises.)  Tests only.
    """
    q = quant if quant is not None else (lambda t: t)
    if quant is not None:
        sd = {k: (q(v) if (k.endswith("conv1.weight") or k.endswith("conv2.weight") or k.endswith("conv3.weight")
                           or k.endswith("downsample.0.weight")
                           or k == "decode_head.convs.0.0.weight") else v) for k, v in sd.items()}
    def tap(name, t):
        if taps is not None:
            taps[name] = t
        return t

    def act(name, t):
        if relu_masks is not None and name in relu_masks:
            if taps is not None:
                taps[name + ".pre"] = t.detach()
            return t * relu_masks[name].to(t.dtype)
        return torch.relu(t)

    L = x.shape[2]
    h = q(F.conv1d(x, sd["backbone.stem.0.weight"], None, stride=2, padding=3))
    tap("backbone.stem.0", h)
    if relu_masks is not None and "backbone.maxpool" in relu_masks:
        h = q(act("backbone.maxpool", maxpool_k3s2p1(batchnorm(h, sd, "backbone.stem.1", train, new_buffers))))
    else:
        h = torch.relu(batchnorm(h, sd, "backbone.stem.1", train, new_buffers))
        tap("backbone.stem", h)
        h = q(maxpool_k3s2p1(h))
    tap("backbone.maxpool", h)
    feats = []
    inpl = arch.stem_channels
    for i, nb in enumerate(arch.stage_blocks):
        pl = arch.planes(i)
        for j in range(nb):
            pre = f"backbone.layer{i + 1}.{j}"
            s = arch.strides[i] if j == 0 else 1
            ident = h
            if arch.bottleneck:       # resnet.py:112-132: 1x1 - BN - ReLU - 3x3(stride) - BN - ReLU - 1x1 - BN (+ identity) - ReLU
                o = q(F.conv1d(h, sd[pre + ".conv1.weight"], None))
                tap(pre + ".conv1", o)
                o = q(torch.relu(batchnorm(o, sd, pre + ".bn1", train, new_buffers)))
                tap(pre + ".relu1", o)
                o = q(F.conv1d(o, sd[pre + ".conv2.weight"], None, stride=s, padding=1))
                tap(pre + ".conv2", o)
                o = q(torch.relu(batchnorm(o, sd, pre + ".bn2", train, new_buffers)))
                tap(pre + ".relu2", o)
                o = q(F.conv1d(o, sd[pre + ".conv3.weight"], None))
                tap(pre + ".conv3", o)
                o = batchnorm(o, sd, pre + ".bn3", train, new_buffers)
                if (pre + ".downsample.0.weight") in sd:
                    ident = q(F.conv1d(h, sd[pre + ".downsample.0.weight"], None, stride=s, padding=0))
                    tap(pre + ".downsample.0", ident)
                    ident = batchnorm(ident, sd, pre + ".downsample.1", train, new_buffers)
                h = q(torch.relu(o + ident))
                tap(pre, h)
                continue
            o = q(F.conv1d(h, sd[pre + ".conv1.weight"], None, stride=s, padding=1))
            tap(pre + ".conv1", o)
            o = q(act(pre + ".relu1", batchnorm(o, sd, pre + ".bn1", train, new_buffers)))
            tap(pre + ".relu1", o)
            o = q(F.conv1d(o, sd[pre + ".conv2.weight"], None, stride=1, padding=1))
            tap(pre + ".conv2", o)
            o = batchnorm(o, sd, pre + ".bn2", train, new_buffers)
            if (pre + ".downsample.0.weight") in sd:
                ident = q(F.conv1d(h, sd[pre + ".downsample.0.weight"], None, stride=s, padding=0))
                tap(pre + ".downsample.0", ident)
                ident = batchnorm(ident, sd, pre + ".downsample.1", train, new_buffers)
            h = q(act(pre, o + ident))
            tap(pre, h)
        inpl = pl
        feats.append(h)
    f = feats[arch.in_index]
    h = q(F.conv1d(f, sd["decode_head.convs.0.0.weight"], None, stride=1, padding=1))
    tap("decode_head.convs.0.0", h)
    h = q(act("decode_head.convs.0", batchnorm(h, sd, "decode_head.convs.0.1", train, new_buffers)))
    tap("decode_head.convs.0", h)
    if train and dropout_mask is not None and arch.dropout_ratio > 0:
        h = h * dropout_mask.to(h.dtype) / (1.0 - arch.dropout_ratio)
    low = F.conv1d(h, sd["decode_head.cls_seg.weight"], sd["decode_head.cls_seg.bias"])
    tap("decode_head.cls_seg", low)
    seg = linear_upsample(low, L, arch.align_corners)
    return {"seg_logits": seg, "low_logits": low}


# --------------------------------------------------------------------------------------
# pseudo-labels and losses
# --------------------------------------------------------------------------------------
def pseudo_label(logits: Tensor, conf_thresh: float) -> Tuple[Tensor, Tensor, Tensor]:
    """fixmatch.py:89-91,115: conf = softmax(1).max(1)[0]; label = argmax(1); mask = conf >= thr.

    Uses torch's own soft-max/arg-max so that, run on a CUDA tensor, it is the exact
    comparison target for the bit-exactness tests.
    """
    conf = logits.softmax(dim=1).max(dim=1)[0]
    label = logits.argmax(dim=1)
    mask = conf >= conf_thresh
    return conf, label, mask


def log_softmax_c(z: Tensor) -> Tensor:
    m = z.max(dim=1, keepdim=True).values
    return z - m - torch.log(torch.exp(z - m).sum(dim=1, keepdim=True))


def ce_hard(z: Tensor, y: Tensor) -> Tensor:
    """F.cross_entropy(z, y) mean over all B*L positions (fixmatch.py:105; encoder_decoder.py:110-111)."""
    if FAST_KERNELS:
        return F.cross_entropy(z, y)
    return -(log_softmax_c(z).gather(1, y[:, None, :]).squeeze(1)).mean()


def ce_masked(z: Tensor, y: Tensor, mask: Tensor) -> Tensor:
    """(CE(z, y, 'none') * mask).mean(): denominator is ALL positions (fixmatch.py:114-116)."""
    if FAST_KERNELS:
        return (F.cross_entropy(z, y, reduction="none") * mask).mean()
    per = -(log_softmax_c(z).gather(1, y[:, None, :]).squeeze(1))
    return (per * mask.to(per.dtype)).mean()


def ce_soft(z: Tensor, p: Tensor) -> Tensor:
    """F.cross_entropy(z, probs): mean over B*L of -sum_c p_c log softmax(z)_c (mean_teacher.py:115)."""
    if FAST_KERNELS:
        return F.cross_entropy(z, p)
    return -(p * log_softmax_c(z)).sum(dim=1).mean()


def ce_soft_masked(z: Tensor, p: Tensor, mask: Tensor) -> Tensor:
    """reco.py:248-250: F.cross_entropy(pred_u_s, prob_u_w, reduction='none') * (conf_u_w >= thr), mean over ALL
    positions."""
    per = -(p * log_softmax_c(z)).sum(dim=1)
    return (per * mask.to(per.dtype)).mean()


def lr_at(epoch: float, cfg: dict) -> float:
    """utils/lr_sched.py:6-18 -- linear warm-up then half-cosine to min_lr."""
    if epoch < cfg["warmup_epochs"]:
        return cfg["lr"] * epoch / cfg["warmup_epochs"]
    return cfg["min_lr"] + (cfg["lr"] - cfg["min_lr"]) * 0.5 * (
        1.0 + math.cos(math.pi * (epoch - cfg["warmup_epochs"]) / (cfg["epochs"] - cfg["warmup_epochs"])))


def adamw_update(p: Tensor, g: Tensor, m: Tensor, v: Tensor, t: int, lr: float,
                 betas=(0.9, 0.999), eps: float = 1e-8, wd: float = 0.05) -> None:
    """torch.optim.AdamW single step, in place (utils/optimizer.py:22-34; amsgrad off)."""
    b1, b2 = betas
    p.mul_(1.0 - lr * wd)
    m.mul_(b1).add_(g, alpha=1.0 - b1)
    v.mul_(b2).addcmul_(g, g, value=1.0 - b2)
    bc1 = 1.0 - b1 ** t
    bc2 = 1.0 - b2 ** t
    denom = (v.sqrt() / math.sqrt(bc2)).add_(eps)
    p.addcdiv_(m, denom, value=-(lr / bc1))


# --------------------------------------------------------------------------------------
# the semi-supervised step
# --------------------------------------------------------------------------------------
class OracleTrainer:
    """Stateful restatement of the train_one_epoch bodies:
    fixmatch.py:73-138, mean_teacher.py:76-149, base.py:110-150 (use_amp=False path,
    GradScaler disabled as it is on CPU -- misc.py:239-253)."""

    def __init__(self, sd: Dict[str, Tensor], arch: Arch, train_cfg: dict, dtype=torch.float64):
        self.arch = arch
        self.cfg = train_cfg
        self.dtype = dtype
        self.pnames = param_names(arch)
        self.bnames = buffer_names(arch)
        self.sd: Dict[str, Tensor] = {}
        for k, v in sd.items():
            if k in self.pnames or (v.is_floating_point()):
                self.sd[k] = v.detach().clone().to(dtype)
            else:
                self.sd[k] = v.detach().clone()
        self.m = {k: torch.zeros_like(self.sd[k]) for k in self.pnames}
        self.v = {k: torch.zeros_like(self.sd[k]) for k in self.pnames}
        self.t = 0
        self.grads: Dict[str, Tensor] = {}
        self.taps: Dict[str, Tensor] = {}
        # mean-teacher state (mean_teacher.py:281-290): params alias the student until the
        # first EMA; buffers are the teacher's own fresh-init copies.
        self.teacher_sd: Optional[Dict[str, Tensor]] = None
        self.teacher_aliased = True
        self.quant = None   # set to bf16_round to emulate the BF16 path's storage roundings
        self.relu_masks = None   # tests: injected ReLU sign decisions of the student's TRAIN forward (see forward())

    # ---- helpers -------------------------------------------------------------------
    def _leaf_params(self) -> Dict[str, Tensor]:
        sdg = dict(self.sd)
        for k in self.pnames:
            sdg[k] = self.sd[k].detach().clone().requires_grad_(True)
        return sdg

    def _apply_grads(self, sdg: Dict[str, Tensor], lr: float) -> None:
        self.t += 1
        kw = self.cfg.get("optimizer_kwargs", {}) or {}
        betas = tuple(kw.get("betas", (0.9, 0.999)))
        eps = kw.get("eps", 1e-8)
        # clip_grad=max_norm (fixmatch.py:129-136 -> misc.py:248-250, torch.nn.utils.clip_grad_norm_): every gradient times
        # min(1, max_norm / (global L2 norm + 1e-6)); self.grads keeps the UNCLIPPED gradients (the returned norm is theirs)
        max_norm = self.cfg.get("max_norm", None)
        coef = 1.0
        if max_norm is not None:
            total = math.sqrt(sum(float((sdg[k].grad.double() ** 2).sum()) for k in self.pnames))
            coef = min(1.0, float(max_norm) / (total + 1e-6))
        for k in self.pnames:
            g = sdg[k].grad
            self.grads[k] = g.detach().clone()
            if max_norm is not None:
                g = g * coef
            adamw_update(self.sd[k], g, self.m[k], self.v[k], self.t, lr, betas, eps,
                         self.cfg["weight_decay"])

    def grad_norm(self) -> float:
        """misc.py:265-278 -- global L2 norm over all parameter gradients."""
        return math.sqrt(sum(float((g.double() ** 2).sum()) for g in self.grads.values()))

    def init_teacher(self, teacher_buffers: Optional[Dict[str, Tensor]] = None) -> None:
        self.teacher_sd = {}
        for k in self.bnames:
            src = teacher_buffers[k] if teacher_buffers is not None else self.sd[k]
            self.teacher_sd[k] = src.detach().clone().to(self.dtype) if src.is_floating_point() else src.detach().clone()
        self.teacher_aliased = True

    def _teacher_view(self) -> Dict[str, Tensor]:
        tv = dict(self.teacher_sd)
        if self.teacher_aliased:
            for k in self.pnames:
                tv[k] = self.sd[k]
        return tv

    def _ema(self, d: float) -> None:
        """mean_teacher.py:139-149: k <- k*d + q*(1-d) over parameters AND buffers (the
        int64 num_batches_tracked is promoted to float32 exactly as torch does)."""
        tv = self._teacher_view()
        for k in self.pnames:
            self.teacher_sd[k] = tv[k] * d + self.sd[k] * (1.0 - d)
        for k in self.bnames:
            self.teacher_sd[k] = self.teacher_sd[k] * d + self.sd[k] * (1.0 - d)
        self.teacher_aliased = False

    # ---- steps ---------------------------------------------------------------------
    def supervised_step(self, ecg: Tensor, target: Tensor, lr: float,
                        dropout_mask: Optional[Tensor] = None, want_taps: bool = False) -> Dict[str, float]:
        """base.py:110-150: model(x, y, return_loss=True) -> CE -> backward -> AdamW."""
        sdg = self._leaf_params()
        nb: Dict[str, Tensor] = {}
        self.taps = {} if want_taps else None
        out = forward(sdg, ecg.to(self.dtype), self.arch, True, dropout_mask, nb, self.taps, quant=self.quant,
                      relu_masks=self.relu_masks)
        if want_taps:
            for t_ in self.taps.values():
                if t_.requires_grad:
                    t_.retain_grad()
        self.low_logits = out["low_logits"]
        self.low_logits.retain_grad()
        loss = ce_hard(out["seg_logits"], target)
        loss.backward()
        self.sd.update(nb)
        self._apply_grads(sdg, lr)
        return {"loss": float(loss.detach())}

    def fixmatch_step(self, ecg_x: Tensor, mask_x: Tensor, ecg_u_w: Tensor, ecg_u_s: Tensor,
                      lr: float, dropout_mask: Optional[Tensor] = None, want_taps: bool = False) -> Dict[str, float]:
        """fixmatch.py:79-138."""
        thr = self.cfg["conf_thresh"]
        with torch.no_grad():
            pw = forward(self.sd, ecg_u_w.to(self.dtype), self.arch, False, quant=self.quant)["seg_logits"]
            conf = pw.softmax(dim=1).max(dim=1)[0]
            label = pw.argmax(dim=1)
            # the reference compares an fp32 conf with the python float threshold
            mask = conf >= (torch.tensor(thr, dtype=torch.float32).to(conf.dtype) if conf.dtype == torch.float32 else thr)
        self.pseudo = {"conf": conf, "label": label, "mask": mask, "logits_w": pw}
        sdg = self._leaf_params()
        nb: Dict[str, Tensor] = {}
        self.taps = {} if want_taps else None
        nl = ecg_x.shape[0]
        out = forward(sdg, torch.cat((ecg_x, ecg_u_s)).to(self.dtype), self.arch, True, dropout_mask, nb, self.taps, quant=self.quant,
                      relu_masks=self.relu_masks)
        if want_taps:
            for t_ in self.taps.values():
                if t_.requires_grad:
                    t_.retain_grad()
        self.low_logits = out["low_logits"]
        self.low_logits.retain_grad()
        seg = out["seg_logits"]
        loss_x = ce_hard(seg[:nl], mask_x)
        loss_u = ce_masked(seg[nl:], label, mask)
        loss = (loss_x + loss_u) / 2.0
        loss.backward()
        self.sd.update(nb)
        self._apply_grads(sdg, lr)
        return {"loss_total": float(loss.detach()), "loss_x": float(loss_x.detach()),
                "loss_u_s": float(loss_u.detach()), "mask_ratio": float(mask.to(torch.float32).mean())}

    def mean_teacher_step(self, ecg_x: Tensor, mask_x: Tensor, ecg_u_w: Tensor, ecg_u_s: Tensor,
                          lr: float, dropout_mask: Optional[Tensor] = None, want_taps: bool = False) -> Dict[str, float]:
        """mean_teacher.py:82-149."""
        if self.teacher_sd is None:
            self.init_teacher()
        d = self.cfg.get("ema_decay", 0.999)
        with torch.no_grad():
            pw = forward(self._teacher_view(), ecg_u_w.to(self.dtype), self.arch, False, quant=self.quant)["seg_logits"]
            prob = pw.softmax(dim=1)
        self.pseudo = {"prob": prob, "logits_w": pw}
        sdg = self._leaf_params()
        nb: Dict[str, Tensor] = {}
        self.taps = {} if want_taps else None
        nl = ecg_x.shape[0]
        out = forward(sdg, torch.cat((ecg_x, ecg_u_s)).to(self.dtype), self.arch, True, dropout_mask, nb, self.taps, quant=self.quant,
                      relu_masks=self.relu_masks)
        self.low_logits = out["low_logits"]
        self.low_logits.retain_grad()
        seg = out["seg_logits"]
        loss_x = ce_hard(seg[:nl], mask_x)
        loss_u = ce_soft(seg[nl:], prob)
        loss = (loss_x + loss_u) / 2.0
        loss.backward()
        self.sd.update(nb)
        self._apply_grads(sdg, lr)
        self._ema(d)
        return {"loss_total": float(loss.detach()), "loss_x": float(loss_x.detach()),
                "loss_u_s": float(loss_u.detach())}

    def hard_label_step(self, ecg_x: Tensor, mask_x: Tensor, ecg_u_w: Tensor, label_u: Tensor, lr: float,
                        want_taps: bool = False) -> Dict[str, float]:
        """The student half of cps.py:108-149 and stpp.py:154-197: train-mode forward of cat(ecg_x, ecg_u_w),
        CE against the labels and against given hard pseudo-labels (mean over every position), /2, AdamW."""
        sdg = self._leaf_params()
        nb: Dict[str, Tensor] = {}
        nl = ecg_x.shape[0]
        self.taps = {} if want_taps else None
        out = forward(sdg, torch.cat((ecg_x, ecg_u_w)).to(self.dtype), self.arch, True, None, nb, self.taps, quant=self.quant,
                      relu_masks=self.relu_masks)
        seg = out["seg_logits"]
        loss_x = ce_hard(seg[:nl], mask_x)
        loss_u = ce_hard(seg[nl:], label_u)
        loss = (loss_x + loss_u) / 2.0
        loss.backward()
        self.sd.update(nb)
        self._apply_grads(sdg, lr)
        return {"loss_total": float(loss.detach()), "loss_x": float(loss_x.detach()), "loss_u_s": float(loss_u.detach())}

    def hard_labels(self, ecg_u_w: Tensor, sd: Optional[Dict[str, Tensor]] = None) -> Tensor:
        """cps.py:98-103 / stpp.py:150-152: eval-mode forward (running statistics), argmax over classes."""
        with torch.no_grad():
            return forward(sd if sd is not None else self.sd, ecg_u_w.to(self.dtype), self.arch, False,
                           quant=self.quant)["seg_logits"].argmax(dim=1)

    def stpp_step(self, ecg_x: Tensor, mask_x: Tensor, ecg_u_w: Tensor, teacher_sd: Dict[str, Tensor], lr: float) -> Dict[str, float]:
        """stpp.py:140-197: hard pseudo-labels of a frozen teacher."""
        return self.hard_label_step(ecg_x, mask_x, ecg_u_w, self.hard_labels(ecg_u_w, teacher_sd), lr)

    def teacher_state(self) -> Dict[str, Tensor]:
        return self._teacher_view()


def cps_step(tr_1: "OracleTrainer", tr_2: "OracleTrainer", ecg_x: Tensor, mask_x: Tensor, ecg_u_w: Tensor,
             lr: float) -> Dict[str, float]:
    """cps.py:96-160: both models label the weak views BEFORE either is updated; each then trains against the other's
    labels; the logged losses are the two models' means."""
    lab_1, lab_2 = tr_1.hard_labels(ecg_u_w), tr_2.hard_labels(ecg_u_w)
    s1 = tr_1.hard_label_step(ecg_x, mask_x, ecg_u_w, lab_2, lr)
    s2 = tr_2.hard_label_step(ecg_x, mask_x, ecg_u_w, lab_1, lr)
    return {k: (s1[k] + s2[k]) / 2.0 for k in s1}


# --------------------------------------------------------------------------------------
# closed-form gradient of the fused loss w.r.t. the LOW-RES logits (used to check the
# fused CUDA loss kernel without autograd)
# --------------------------------------------------------------------------------------
def loss_grad_fullres(z: Tensor, nl: int, mask_x: Tensor, label_u: Optional[Tensor],
                      mask_u: Optional[Tensor], prob_u: Optional[Tensor]) -> Tensor:
    """d[(loss_x + loss_u)/2]/dz for z=[B_l+B_u, C, L] (Appendix A of SURVEY.md):
    (softmax - onehot)/(2*B_l*L) on labeled rows; mask*(softmax - onehot)/(2*B_u*L) or
    (softmax - p)/(2*B_u*L) on unlabeled rows."""
    p = z.softmax(dim=1)
    g = torch.zeros_like(z)
    L = z.shape[2]
    nu = z.shape[0] - nl
    oh = F.one_hot(mask_x, z.shape[1]).permute(0, 2, 1).to(z.dtype)
    g[:nl] = (p[:nl] - oh) / (2.0 * nl * L)
    if nu > 0:
        if prob_u is not None:
            g[nl:] = (p[nl:] - prob_u.to(z.dtype)) / (2.0 * nu * L)
        else:
            ohu = F.one_hot(label_u, z.shape[1]).permute(0, 2, 1).to(z.dtype)
            g[nl:] = (p[nl:] - ohu) * mask_u.to(z.dtype)[:, None, :] / (2.0 * nu * L)
    return g
