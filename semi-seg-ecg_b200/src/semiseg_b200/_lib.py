"""ctypes binding of libsemiseg_b200.so (C ABI declared in include/ssb.h).

The product path has no CPU fallback: if the shared library is missing, `load()` raises,
and every op raises `RuntimeError` with the library's own message on a non-zero status.
"""
from __future__ import annotations

import ctypes as C
import os
from typing import Optional

F32, BF16 = 0, 1
ALGO_SIMT, ALGO_TCGEN05 = 0, 1
LOSS_SUP, LOSS_FIXMATCH, LOSS_SOFT, LOSS_SOFT_MASKED = 0, 1, 2, 3

_HERE = os.path.dirname(os.path.abspath(__file__))
PKG_ROOT = os.path.abspath(os.path.join(_HERE, "..", ".."))          # semi-seg-ecg_b200/
LIB_PATH = os.path.join(PKG_ROOT, "lib", "libsemiseg_b200.so")
CSRC_DIR = os.path.join(PKG_ROOT, "csrc")


class Geom(C.Structure):
    _fields_ = [("B", C.c_int32), ("pitch", C.c_int32), ("len", C.c_int32), ("C", C.c_int32)]

    def rows(self) -> int:
        return self.B * self.pitch

    def __repr__(self):
        return f"Geom(B={self.B}, pitch={self.pitch}, len={self.len}, C={self.C})"


class BN(C.Structure):
    _fields_ = [("gamma", C.c_void_p), ("beta", C.c_void_p), ("running_mean", C.c_void_p),
                ("running_var", C.c_void_p), ("num_batches_tracked", C.c_void_p), ("sums", C.c_void_p),
                ("mean_invstd", C.c_void_p), ("bwd_sums", C.c_void_p), ("dgamma", C.c_void_p),
                ("dbeta", C.c_void_p), ("count_mul", C.c_int32), ("pad", C.c_int32),
                # SyncBN exchange inside the consuming kernels (include/ssb.h: ssb_bn.sync_*); all zero = off
                ("sync_peers", C.c_void_p), ("sync_sp", C.c_void_p), ("sync_world", C.c_int32), ("sync_rank", C.c_int32),
                ("sync_slot", C.c_uint32), ("sync_fwd_off", C.c_uint32), ("sync_bwd_off", C.c_uint32), ("pad2", C.c_uint32)]


class StepParams(C.Structure):
    _fields_ = [("lr", C.c_float), ("inv_bias1", C.c_float), ("inv_sqrt_bias2", C.c_float),
                ("ema_decay", C.c_float), ("ema_first", C.c_int32), ("step", C.c_int32),
                ("rng_seed", C.c_uint32), ("rng_step", C.c_uint32), ("grad_scale", C.c_float),
                ("conf_thresh", C.c_float), ("pad", C.c_float * 6)]


class AugOp(C.Structure):
    _fields_ = [("kind", C.c_int32), ("apply", C.c_int32), ("a", C.c_int32), ("b", C.c_int32)]


AUG_AMPLITUDE, AUG_POWERLINE, AUG_PARTIAL_WHITE, AUG_PARTIAL_SINE = 0, 1, 2, 3

assert C.sizeof(StepParams) == 64 and C.sizeof(Geom) == 16 and C.sizeof(AugOp) == 16 and C.sizeof(BN) == 128

_P, _I, _F, _SZ, _D = C.c_void_p, C.c_int, C.c_float, C.c_size_t, C.c_double
_BNP = C.POINTER(BN)

# name -> argtypes (restype is always int unless listed in _RESTYPES); this table is also what
# tests/test_cabi.py checks against include/ssb.h
SIGNATURES = {
    "ssb_version": [],
    "ssb_last_error": [],
    "ssb_device_check": [],
    "ssb_launch_count": [],
    "ssb_prepare": [],
    "ssb_memset_zero": [_P, _SZ, _P],
    "ssb_stem_conv_fwd": [_P, _P, _P, _I, _I, Geom, _I, _P],
    "ssb_stem_conv_fwd_stats": [_P, _P, _P, _I, _I, Geom, _P, _I, _P],
    "ssb_stem_conv_wgrad": [_P, _P, _P, _I, _I, Geom, _I, _P],
    "ssb_conv1d_fwd": [_P, _P, _P, Geom, Geom, _I, _I, _I, _I, _P],
    "ssb_conv1d_fwd_stats": [_P, _P, _P, Geom, Geom, _I, _I, _P, _I, _I, _P],
    "ssb_conv1d_bn_act_fwd": [_P, _P, _P, Geom, Geom, _I, _I, _BNP, _P, _I, _I, _I, _P],
    "ssb_conv1d_fwd_bn_train_fits": [Geom, Geom, _I, _I, _I, _I],
    "ssb_conv1d_fwd_bn_train": [_P, _P, _P, _P, Geom, Geom, _I, _I, _BNP, _P, _BNP, _I, _P, _I, _I, _P],
    "ssb_conv1d_fwd_dual": [_P, _P, _P, _P, Geom, Geom, _I, _I, _I, _P, _BNP, _P, _I, _I, _I, _P],
    "ssb_conv1d_dgrad": [_P, _P, _P, Geom, Geom, _I, _I, _I, _I, _I, _P],
    "ssb_conv1d_dgrad_bnred": [_P, _P, _P, Geom, Geom, _I, _I, _I, _P, _P, _BNP, _P, _BNP, _I, _I, _P],
    "ssb_conv1d_wgrad": [_P, _P, _P, Geom, Geom, _I, _I, _I, _I, _P],
    "ssb_weight_shadow": [_P, _P, _SZ, _I, _P],
    "ssb_bn_stats": [_P, Geom, _P, _I, _P],
    "ssb_bn_act_fwd": [_P, _BNP, _P, _BNP, _P, Geom, _I, _I, _I, _P],
    "ssb_stem_bn_relu_pool_fwd": [_P, _BNP, _P, _P, Geom, Geom, _I, _I, _P],
    "ssb_bn_bwd_reduce": [_P, _P, _P, _P, _BNP, _P, _BNP, Geom, _I, _P],
    "ssb_bn_bwd_apply": [_P, _P, _P, _P, _BNP, _P, _P, _BNP, _P, _P, Geom, _I, _P],
    "ssb_bn_bwd_fused_fits": [Geom, _I, _I, _I],
    "ssb_eval_metrics": [_P, _P, _P, _P, _P, _P, _I, _I, _I, _I, _I, _P],
    "ssb_bn_bwd_fused": [_P, _P, _P, _BNP, _P, _P, _BNP, _P, _P, Geom, _P, _P, _P, _SZ, _I, _P],
    "ssb_stem_bwd_reduce": [_P, _P, _P, _BNP, Geom, Geom, _I, _P],
    "ssb_stem_bwd_apply": [_P, _P, _P, _BNP, _P, Geom, Geom, _I, _P],
    "ssb_head_cls_fwd": [_P, _P, _P, _P, Geom, _I, _F, _P, _P, _I, _P],
    "ssb_head_cls_bwd": [_P, _P, _P, _P, _P, _P, Geom, _I, _F, _P, _P, _I, _P],
    "ssb_upsample_fwd": [_P, _P, _I, _I, _I, _I, _I, _P],
    "ssb_upsample_bwd": [_P, _P, _I, _I, _I, _I, _I, _P],
    "ssb_pseudo_label": [_P, _F, _P, _P, _P, _I, _I, _I, _P],
    "ssb_semi_loss": [_P, _P, _P, _P, _P, _I, _I, _I, _I, _I, _I, _F, _P, _I, _P, _P, _P, _P],
    "ssb_adamw_ema": [_P, _P, _P, _P, _P, _SZ, _D, _D, _D, _D, _P, _P],
    "ssb_adamw_ema_clip": [_P, _P, _P, _P, _P, _SZ, _D, _D, _D, _D, _P, _P, _D, _P],
    "ssb_ema": [_P, _P, _SZ, _P, _P],
    "ssb_ema_i64": [_P, _P, _SZ, _P, _P],
    "ssb_grad_norm": [_P, _SZ, _P, _P, _P],
    "ssb_syncbn_mailbox_bytes": [_I],
    "ssb_syncbn_fused_mailbox_bytes": [_I, _I],
    "ssb_syncbn_exchange": [_P, _I, _P, _I, _I, _I, _P],
    "ssb_aug_spectrum": [_P, _P, _P, _I, _I, _I, _P],
    "ssb_aug_resize_crop": [_P, _P, _P, _P, _P, _P, _I, _I, _I, _I, _P],
    "ssb_aug_strong_standardize": [_P, _P, _P, _I, _P, _P, C.c_uint32, _P, _I, _I, _I, _I, _F, _P],
}
_RESTYPES = {"ssb_last_error": C.c_char_p, "ssb_launch_count": C.c_int64, "ssb_syncbn_mailbox_bytes": C.c_size_t,
             "ssb_syncbn_fused_mailbox_bytes": C.c_size_t}

_lib: Optional[C.CDLL] = None


def load(path: Optional[str] = None) -> C.CDLL:
    """dlopen the kernel library (no CUDA initialisation happens at load time)."""
    global _lib
    if _lib is not None and path is None:
        return _lib
    p = path or os.environ.get("SSB_LIB", LIB_PATH)
    if not os.path.exists(p):
        raise RuntimeError(
            f"libsemiseg_b200.so not found at {p}: build it with `make -C {CSRC_DIR}` "
            "(or `python -c 'import __graft_entry__ as g; g.build()'`). There is no CPU fallback.")
    lib = C.CDLL(p)
    for name, argtypes in SIGNATURES.items():
        fn = getattr(lib, name)  # AttributeError => symbol missing => fail loudly
        fn.argtypes = argtypes
        fn.restype = _RESTYPES.get(name, C.c_int)
    if lib.ssb_version() != 1:
        raise RuntimeError(f"libsemiseg_b200.so ABI version {lib.ssb_version()} != 1")
    if path is None:
        _lib = lib
    return lib


_prepared = False


def prepare() -> None:
    """One-time device setup (ssb_prepare); also verifies the device is sm_100."""
    global _prepared
    if not _prepared:
        check(load().ssb_prepare(), "ssb_prepare")
        _prepared = True


def check(status: int, what: str = "") -> None:
    if status != 0:
        msg = load().ssb_last_error()
        raise RuntimeError(f"libsemiseg_b200 {what} failed ({status}): {msg.decode() if msg else '?'}")


_hook = None  # bench.py installs a per-launch timing hook here


def call(name: str, *args) -> None:
    """Invoke an int-returning entry point and raise on error."""
    if _hook is not None:
        _hook(name, args)
        return
    check(getattr(load(), name)(*args), name)


def raw_call(name: str, *args) -> None:
    check(getattr(load(), name)(*args), name)
