"""EncoderDecoder segmentor: the reference's module surface
(src/models/encoder_decoder.py:10-136) over the libsemiseg_b200 kernels.

forward(inputs[B,C,L] f32, labels[B,L] i64 | None, return_loss, return_latent) ->
{'seg_logits': [B, num_classes, L] f32, 'loss'?}.  Inputs must live on a CUDA (sm_100) device:
the hot path has no CPU fallback.  Precision follows autocast like the reference's step does
(`torch.autocast` enabled -> bf16 storage / tcgen05 tensor-core convs; otherwise the fp32
exact-parity kernels).  The fused training step (algorithms.*.train_one_epoch) bypasses this
module-level path and drives the same kernels through `semiseg_b200.engine.StepEngine`.
"""
from typing import Optional

import torch
import torch.nn as nn
from torch import Tensor

from semiseg_b200 import _lib
from semiseg_b200.runtime import ModelRuntime, _SegNetFn


class EncoderDecoder(nn.Module):
    def __init__(self, backbone: nn.Module, decode_head: nn.Module, decode_head_loss: Optional[nn.Module] = None,
                 auxiliary_heads: Optional[nn.ModuleList] = None, auxiliary_head_losses: Optional[nn.ModuleList] = None,
                 use_latent_projection: bool = False, projection_in_dim: Optional[int] = None,
                 projection_out_dim: Optional[int] = None):
        super().__init__()
        if auxiliary_heads is not None or auxiliary_head_losses is not None:
            raise NotImplementedError("auxiliary heads are unused by every shipped config and dead code in the "
                                      "reference (encoder_decoder.py:113-134); not part of the hot path")
        if use_latent_projection:
            raise NotImplementedError("latent projection is ReCo-only (out of scope, SURVEY.md section 2)")
        self.backbone = backbone
        self.decode_head = decode_head
        self.loss_decode = decode_head_loss
        self._rt: Optional[ModelRuntime] = None
        self.precision: Optional[str] = None   # None: follow autocast; or 'fp32' / 'bf16'

    # ---- reference surface -----------------------------------------------------------
    @property
    def with_auxiliary_heads(self) -> bool:
        return False

    @property
    def with_decode_head(self) -> bool:
        return self.decode_head is not None

    @property
    def with_projection(self) -> bool:
        return False

    def no_weight_decay(self):
        out = set()
        if hasattr(self.backbone, "no_weight_decay"):
            out |= set(self.backbone.no_weight_decay())
        if hasattr(self.decode_head, "no_weight_decay"):
            out |= set(self.decode_head.no_weight_decay())
        return out

    # ---- runtime -------------------------------------------------------------------
    def runtime(self, nbt_float: bool = False) -> ModelRuntime:
        if self._rt is None or (nbt_float and not self._rt.nbt_float):
            object.__setattr__(self, "_rt", ModelRuntime(self, nbt_float=nbt_float))
        return self._rt

    def _dtype(self) -> int:
        if self.precision is not None:
            return {"fp32": _lib.F32, "bf16": _lib.BF16}[self.precision]
        return _lib.BF16 if torch.is_autocast_enabled() else _lib.F32

    def forward(self, inputs: Tensor, labels: Optional[Tensor] = None, return_loss: bool = False,
                return_latent: bool = False) -> dict:
        if return_latent:
            raise NotImplementedError("return_latent is ReCo-only (out of scope)")
        if not inputs.is_cuda:
            raise RuntimeError("EncoderDecoder.forward: inputs must be CUDA tensors; the B200 hot path has no "
                               "CPU fallback (libsemiseg_b200 kernels only)")
        rt = self.runtime()
        rt.ensure()
        B, _, L = inputs.shape
        need_grad = self.training and torch.is_grad_enabled()
        plan = rt.plan(self._dtype(), B, L, train=need_grad or self.training)
        params = [p for _, p in self.named_parameters()]
        if need_grad:
            seg_logits = _SegNetFn.apply(inputs, rt, plan, True, *params)
        else:
            with torch.no_grad():
                seg_logits = _SegNetFn.apply(inputs, rt, plan, self.training, *params)
        outputs = {"seg_logits": seg_logits}
        if return_loss:
            outputs["loss"] = self.loss_decode(seg_logits, labels)
        return outputs
