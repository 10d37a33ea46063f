"""FusedAdamW: torch.optim.Optimizer facade over the fused multi-tensor AdamW kernel.

State (`exp_avg`, `exp_avg_sq`) is exposed through `optimizer.state[p]` as VIEWS into the flat
optimizer arenas, so `state_dict()` / `load_state_dict()` interchange with torch.optim.AdamW
checkpoints (reference checkpoint format: utils/misc.py:281-321)."""
from __future__ import annotations

import ctypes as C
import math
import weakref
from typing import Optional

import torch

from . import _lib
from ._lib import StepParams, call

_RUNTIMES = []  # weak references to ModelRuntime objects


def register_runtime(rt) -> None:
    _RUNTIMES.append(weakref.ref(rt))


def find_runtime(params):
    ids = {id(p) for p in params}
    for ref in list(_RUNTIMES):
        rt = ref()
        if rt is None:
            _RUNTIMES.remove(ref)
            continue
        mine = {id(p) for p in rt.model.parameters()}
        if ids == mine:
            return rt
    return None


class FusedAdamW(torch.optim.Optimizer):
    def __init__(self, params, lr=1e-3, betas=(0.9, 0.999), eps=1e-8, weight_decay=1e-2):
        super().__init__(params, dict(lr=lr, betas=tuple(betas), eps=eps, weight_decay=weight_decay))
        if len(self.param_groups) != 1:
            raise NotImplementedError("FusedAdamW: one parameter group (the reference uses a single group when "
                                      "layer_decay is null, fixmatch.py:298-308)")
        self._rt = None
        self._sp_dev: Optional[torch.Tensor] = None

    def runtime(self):
        if self._rt is None:
            self._rt = find_runtime(self.param_groups[0]["params"])
            if self._rt is None:
                raise RuntimeError("FusedAdamW: parameters do not belong to a models.EncoderDecoder")
        return self._rt

    def bind_state(self) -> None:
        """Expose arena views as torch-style optimizer state."""
        rt = self.runtime()
        rt.ensure()
        st = rt.state
        mv, vv = rt.weights.param_views(st.exp_avg), rt.weights.param_views(st.exp_avg_sq)
        for (n, p) in rt.model.named_parameters():
            self.state[p] = {"step": torch.tensor(float(st.step)), "exp_avg": mv[n], "exp_avg_sq": vv[n]}

    def state_dict(self):
        self.bind_state()
        return super().state_dict()

    def load_state_dict(self, sd):
        rt = self.runtime()
        rt.ensure()
        st = rt.state
        mv, vv = rt.weights.param_views(st.exp_avg), rt.weights.param_views(st.exp_avg_sq)
        names = [n for n, _ in rt.model.named_parameters()]
        with torch.no_grad():
            for i, n in enumerate(names):
                s = sd["state"].get(i)
                if s is None:
                    continue
                mv[n].copy_(s["exp_avg"].to(mv[n].device))
                vv[n].copy_(s["exp_avg_sq"].to(vv[n].device))
                st.step = int(s["step"])
        g = sd["param_groups"][0]
        for k in ("lr", "betas", "eps", "weight_decay"):
            if k in g:
                self.param_groups[0][k] = g[k]

    @torch.no_grad()
    def step(self, closure=None):
        """Module-level API: gathers p.grad into the gradient arena, then one fused kernel."""
        rt = self.runtime()
        rt.ensure()
        st = rt.state
        g = self.param_groups[0]
        for n, p in rt.model.named_parameters():
            if p.grad is None:
                rt.grad_views[n].zero_()
            elif p.grad.data_ptr() != rt.grad_views[n].data_ptr():
                rt.grad_views[n].copy_(p.grad)
        t = st.step + 1
        b1, b2 = g["betas"]
        sp = StepParams()
        sp.lr = g["lr"]
        sp.inv_bias1 = 1.0 / (1.0 - b1 ** t)
        sp.inv_sqrt_bias2 = 1.0 / math.sqrt(1.0 - b2 ** t)
        sp.grad_scale = 1.0
        sp.step = t
        host = torch.frombuffer(bytearray(bytes(sp)), dtype=torch.uint8)
        if self._sp_dev is None or self._sp_dev.device != st.grads.device:
            self._sp_dev = torch.zeros(64, dtype=torch.uint8, device=st.grads.device)
        self._sp_dev.copy_(host)
        w = rt.weights
        call("ssb_adamw_ema", w.params.data_ptr(), st.grads.data_ptr(), st.exp_avg.data_ptr(),
             st.exp_avg_sq.data_ptr(), None, w.params.numel(), float(b1), float(b2), float(g["eps"]),
             float(g["weight_decay"]), self._sp_dev.data_ptr(), torch.cuda.current_stream().cuda_stream)
        st.step = t
