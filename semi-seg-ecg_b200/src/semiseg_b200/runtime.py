"""Binds an `EncoderDecoder` module tree to the flat arenas and kernel plans.

The nn.Module keeps the reference's surface (parameter names, `state_dict`, `.parameters()`),
but after `adopt()` every parameter / BN buffer is a VIEW into one flat fp32 arena
(`WeightSet.params` / `.bufs`), which is what the kernels read and the fused optimizer updates.
"""
from __future__ import annotations

import ctypes as C
from typing import Dict, Optional, Tuple

import torch

from . import _lib
from ._lib import StepParams, call
from .net import NetPlan, ParamLayout, SegNetSpec, TrainState, WeightSet
from .optim import register_runtime


def spec_from_modules(backbone, head) -> SegNetSpec:
    nstage = len(backbone.stage_blocks)
    idx = head.in_index if head.in_index >= 0 else nstage + head.in_index
    if idx != nstage - 1:
        raise NotImplementedError(f"FCNHead.in_index={head.in_index}: the hot path decodes from the last stage")
    if head.in_channels != backbone.feat_dim:
        raise ValueError(f"FCNHead.in_channels={head.in_channels} != backbone feature width {backbone.feat_dim}")
    return SegNetSpec(num_leads=backbone.num_leads, stem_channels=backbone.stem_channels,
                      base_channels=backbone.base_channels, strides=tuple(backbone.strides),
                      stage_blocks=tuple(backbone.stage_blocks), head_channels=head.channels,
                      num_classes=head.num_classes, dropout_ratio=head.dropout_ratio,
                      align_corners=bool(head.align_corners),
                      bottleneck=getattr(backbone.block, "expansion", 1) == 4)


class ModelRuntime:
    def __init__(self, model, nbt_float: bool = False):
        self.model = model
        self.nbt_float = nbt_float
        self.engines: Dict[Tuple, object] = {}
        self.ema_started = False
        register_runtime(self)
        self.spec = spec_from_modules(model.backbone, model.decode_head)
        self.layout = ParamLayout(self.spec)
        self.weights: Optional[WeightSet] = None
        self.state: Optional[TrainState] = None
        self._plans: Dict[Tuple, NetPlan] = {}
        self._sp_dev: Optional[torch.Tensor] = None
        self._rng_calls = 0
        self.seed = 0
        self._first_last = None

    # ---- arena adoption ------------------------------------------------------------
    def _named(self):
        named_p = dict(self.model.named_parameters())
        named_b = dict(self.model.named_buffers())
        return named_p, named_b

    def adopted(self) -> bool:
        if self.weights is None:
            return False
        named_p, _ = self._named()
        views = self.weights.param_views()
        for n, v in views.items():
            p = named_p[n]
            if p.data_ptr() != v.data_ptr() or p.device != v.device:
                return False
        return True

    def adopt(self, device=None) -> WeightSet:
        """(Re)build the arenas on the parameters' current device and alias the module's
        parameters / buffers to them.  Values are preserved."""
        named_p, named_b = self._named()
        want_p, want_b = self.layout.param_names(), self.layout.buffer_names()
        if list(named_p.keys()) != want_p:
            raise RuntimeError("module parameters do not match the kernel layout: "
                               f"{[n for n in named_p if n not in want_p][:3]} / {[n for n in want_p if n not in named_p][:3]}")
        dev = torch.device(device) if device is not None else next(iter(named_p.values())).device
        if dev.type != "cuda":
            raise RuntimeError("the SemiSegECG B200 hot path runs on CUDA only (no CPU fallback); move the model "
                               "to a CUDA device first")
        _lib.check(_lib.load().ssb_device_check(), "ssb_device_check")
        old_state = self.state
        w = WeightSet(self.layout, dev, nbt_float=self.nbt_float)
        st = TrainState(w)
        pv, bv, gv = w.param_views(), w.buffer_views(), w.param_views(st.grads)
        with torch.no_grad():
            for n in want_p:
                p = named_p[n]
                pv[n].copy_(p.data.to(dev))
                p.data = pv[n]
                p.grad = None
            for n in want_b:
                b = named_b[n]
                bv[n].copy_(b.data.to(device=dev, dtype=bv[n].dtype))
                b.data = bv[n]
            if old_state is not None and old_state.grads.numel() == st.grads.numel():
                st.exp_avg.copy_(old_state.exp_avg.to(dev))
                st.exp_avg_sq.copy_(old_state.exp_avg_sq.to(dev))
                st.step = old_state.step
        self.weights, self.state = w, st
        self.grad_views = gv
        self._plans.clear()
        self.engines.clear()
        self._sp_dev = torch.zeros(64, dtype=torch.uint8, device=dev)
        return w

    def ensure(self) -> WeightSet:
        if not self.adopted():
            self.adopt()
        return self.weights

    def quick_ok(self) -> bool:
        """O(1) version of adopted() for the per-step path: the first and the last parameter still alias the arena."""
        if self.weights is None:
            return False
        ps = self._first_last
        if ps is None:
            plist = list(self.model.parameters())
            ps = self._first_last = (plist[0], plist[-1])
        lay = self.layout.params
        base = self.weights.params.data_ptr()
        return ps[0].data_ptr() == base + 4 * lay[0][2] and ps[1].data_ptr() == base + 4 * lay[-1][2]

    # ---- plans ---------------------------------------------------------------------
    def plan(self, dtype: int, B: int, L: int, train: bool) -> NetPlan:
        self.ensure()
        key = (dtype, B, L, train)
        if key not in self._plans:
            if len(self._plans) >= 6:
                self._plans.pop(next(iter(self._plans)))
            self._plans[key] = NetPlan(self.weights, dtype, B, L, train, None,
                                       grads=self.state.grads if train else None,
                                       sp_ptr=self._sp_dev.data_ptr(), state=self.state if train else None)
            self._plans[key].generation = 0
        return self._plans[key]

    def bump_rng(self) -> None:
        sp = StepParams()
        sp.rng_seed = self.seed & 0xFFFFFFFF
        sp.rng_step = (0x40000000 + self._rng_calls) & 0xFFFFFFFF
        self._rng_calls += 1
        host = torch.frombuffer(bytearray(bytes(sp)), dtype=torch.uint8)
        self._sp_dev.copy_(host)


class _SegNetFn(torch.autograd.Function):
    """Whole-network autograd node: forward/backward run the hand-scheduled plan."""

    @staticmethod
    def forward(ctx, x, rt: ModelRuntime, plan: NetPlan, train_mode: bool, *params):
        st = torch.cuda.current_stream().cuda_stream
        plan.sh.refresh(st)
        if train_mode and rt.spec.dropout_ratio > 0:
            rt.bump_rng()
        xin = x.detach().to(torch.float32).contiguous()
        low = plan.forward(xin, st, train_mode=train_mode)
        B, Lh, ncls = low.shape
        out = torch.empty(B, ncls, plan.L, dtype=torch.float32, device=low.device)
        call("ssb_upsample_fwd", low.data_ptr(), out.data_ptr(), B, Lh, plan.L, ncls,
             1 if rt.spec.align_corners else 0, st)
        plan.generation += 1
        ctx.rt, ctx.plan, ctx.gen, ctx.train_mode = rt, plan, plan.generation, train_mode
        ctx.keep = xin
        return out

    @staticmethod
    def backward(ctx, dout):
        rt, plan = ctx.rt, ctx.plan
        if not ctx.train_mode or not plan.train:
            raise RuntimeError("backward through an eval-mode forward is not supported (BatchNorm running "
                               "statistics path has no backward kernels); call model.train() first")
        if plan.generation != ctx.gen:
            raise RuntimeError("the activations saved for this backward were overwritten by a later forward of "
                               "the same shape; call backward() before the next training forward")
        st = torch.cuda.current_stream().cuda_stream
        dout = dout.to(torch.float32).contiguous()
        B, Lh, ncls = plan.low.shape
        call("ssb_upsample_bwd", dout.data_ptr(), plan.dlow.data_ptr(), B, Lh, plan.L, ncls,
             1 if rt.spec.align_corners else 0, st)
        call("ssb_memset_zero", rt.state.grads.data_ptr(), rt.state.grads.numel() * 4, st)
        plan.backward(plan.dlow, st)
        grads = tuple(rt.grad_views[n].clone() for n in rt.layout.param_names())
        return (None, None, None, None) + grads
