"""Network description, flat parameter arenas and the hand-scheduled forward/backward plan.

This is the host side of the hot path: it owns the memory layout in HBM (flat padded NLC
activations, flat fp32 parameter / gradient / optimizer arenas, storage-dtype weight copy) and
issues the C-ABI kernels of libsemiseg_b200 in dependency order on one CUDA stream.  No
autograd, no ATen compute ops: torch is used for allocation and streams only.

Reference behaviour restated here (file:line relative to the reference tree):
  stem/maxpool/stages     src/models/backbones/resnet.py:206-257, 259-324, 353-363
  BasicBlock              src/models/backbones/resnet.py:55-72
  FCNHead                 src/models/decode_heads/fcn_head.py:36-97
"""
from __future__ import annotations

import ctypes as C
import os
from dataclasses import dataclass, field
from typing import Dict, List, Optional, Sequence, Tuple

import torch

from . import _lib
from ._lib import BN, Geom, StepParams, call


# --------------------------------------------------------------------------------------
@dataclass(frozen=True)
class SegNetSpec:
    """Constructor arguments of resnet18-style backbone + FCNHead that the kernels cover."""
    num_leads: int = 1
    stem_channels: int = 64
    base_channels: int = 64
    strides: Tuple[int, ...] = (1, 2, 2, 2)
    stage_blocks: Tuple[int, ...] = (2, 2, 2, 2)
    head_channels: int = 128
    num_classes: int = 4
    dropout_ratio: float = 0.1
    align_corners: bool = False
    bottleneck: bool = False     # Bottleneck blocks (resnet50/101/152): 1x1 - 3x3(stride) - 1x1(x4)

    def planes(self, i: int) -> int:
        return self.base_channels * 2 ** i

    @property
    def expansion(self) -> int:
        return 4 if self.bottleneck else 1

    def out_planes(self, i: int) -> int:
        return self.planes(i) * self.expansion

    @property
    def feat_dim(self) -> int:
        return self.out_planes(len(self.stage_blocks) - 1)


def conv_out_len(L: int, k: int, s: int, p: int) -> int:
    return (L + 2 * p - (k - 1) - 1) // s + 1


@dataclass
class ConvDesc:
    name: str          # state_dict key of the weight
    cout: int
    cin: int
    k: int
    stride: int
    poff: int = 0      # offset (floats) in the parameter arena
    tap_major: bool = True   # stored [k][Cin][Cout] (GEMM convs); the stem keeps the reference layout


@dataclass
class BNDesc:
    prefix: str        # state_dict prefix
    C: int
    goff: int = 0      # gamma offset in parameter arena
    boff: int = 0      # beta offset
    roff: int = 0      # running_mean offset in buffer arena (running_var at roff + C)
    soff: int = 0      # offset (doubles) in sums arenas; (floats) in mean_invstd arena
    index: int = 0     # index into num_batches_tracked


@dataclass
class BlockDesc:
    prefix: str
    conv1: ConvDesc
    bn1: BNDesc
    conv2: ConvDesc
    bn2: BNDesc
    convd: Optional[ConvDesc]
    bnd: Optional[BNDesc]
    stage: int
    conv3: Optional[ConvDesc] = None     # Bottleneck only
    bn3: Optional[BNDesc] = None


ALIGN = 64  # floats; every tensor in an arena starts on a 256-byte boundary


class ParamLayout:
    """Names, shapes and arena offsets in the reference's `parameters()` / `buffers()` order
    (SURVEY.md 8b-viii).  Single source of truth for state_dict <-> arena mapping."""

    def __init__(self, spec: SegNetSpec):
        self.spec = spec
        self.params: List[Tuple[str, Tuple[int, ...], int]] = []   # (name, shape, offset)
        self.tap_major: Dict[str, bool] = {}
        self.convs: List[ConvDesc] = []
        self.bns: List[BNDesc] = []
        self.blocks: List[BlockDesc] = []
        self._poff = 0
        self._roff = 0
        self._soff = 0

        self.stem_conv = self._conv("backbone.stem.0.weight", spec.stem_channels, spec.num_leads, 7, 2, tap_major=False)
        self.stem_bn = self._bn("backbone.stem.1", spec.stem_channels)
        inpl = spec.stem_channels
        for i, nb in enumerate(spec.stage_blocks):
            pl = spec.planes(i)
            for j in range(nb):
                pre = f"backbone.layer{i + 1}.{j}"
                s = spec.strides[i] if j == 0 else 1
                if spec.bottleneck:
                    po = pl * 4
                    cin = inpl if j == 0 else po
                    c1 = self._conv(pre + ".conv1.weight", pl, cin, 1, 1)
                    b1 = self._bn(pre + ".bn1", pl)
                    c2 = self._conv(pre + ".conv2.weight", pl, pl, 3, s)
                    b2 = self._bn(pre + ".bn2", pl)
                    c3 = self._conv(pre + ".conv3.weight", po, pl, 1, 1)
                    b3 = self._bn(pre + ".bn3", po)
                    cd = bd = None
                    if j == 0 and (s != 1 or inpl != po):
                        cd = self._conv(pre + ".downsample.0.weight", po, inpl, 1, s)
                        bd = self._bn(pre + ".downsample.1", po)
                    self.blocks.append(BlockDesc(pre, c1, b1, c2, b2, cd, bd, i, c3, b3))
                    continue
                cin = inpl if j == 0 else pl
                c1 = self._conv(pre + ".conv1.weight", pl, cin, 3, s)
                b1 = self._bn(pre + ".bn1", pl)
                c2 = self._conv(pre + ".conv2.weight", pl, pl, 3, 1)
                b2 = self._bn(pre + ".bn2", pl)
                cd = bd = None
                if j == 0 and (s != 1 or inpl != pl):
                    cd = self._conv(pre + ".downsample.0.weight", pl, inpl, 1, s)
                    bd = self._bn(pre + ".downsample.1", pl)
                self.blocks.append(BlockDesc(pre, c1, b1, c2, b2, cd, bd, i))
            inpl = pl * spec.expansion
        self.head_conv = self._conv("decode_head.convs.0.0.weight", spec.head_channels, spec.feat_dim, 3, 1)
        self.head_bn = self._bn("decode_head.convs.0.1", spec.head_channels)
        self.cls_w_off = self._param("decode_head.cls_seg.weight", (spec.num_classes, spec.head_channels, 1))
        self.cls_b_off = self._param("decode_head.cls_seg.bias", (spec.num_classes,))
        self.n_params = self._poff
        self.n_bufs = self._roff
        self.n_sums = self._soff
        self.numel = sum(int(torch.Size(s).numel()) for _, s, _ in self.params)

    def _param(self, name, shape) -> int:
        off = self._poff
        self.params.append((name, tuple(shape), off))
        n = 1
        for d in shape:
            n *= d
        self._poff += (n + ALIGN - 1) // ALIGN * ALIGN
        return off

    def _conv(self, name, cout, cin, k, stride, tap_major=True) -> ConvDesc:
        d = ConvDesc(name, cout, cin, k, stride, self._param(name, (cout, cin, k)), tap_major)
        self.tap_major[name] = tap_major
        if tap_major:
            self.convs.append(d)
        return d

    def view(self, arena: torch.Tensor, name: str, shape, off: int) -> torch.Tensor:
        """The tensor `name` inside a flat arena, in the reference's shape.  GEMM-conv weights are stored
        tap-major ([k][Cin][Cout]) and come back as a strided [Cout, Cin, k] view: values and indexing are
        the reference's, only the memory order differs."""
        n = 1
        for d in shape:
            n *= d
        flat = arena[off: off + n]
        if self.tap_major.get(name, False):
            cout, cin, k = shape
            return flat.view(k, cin, cout).permute(2, 1, 0)
        return flat.view(shape)

    def _bn(self, prefix, Cn) -> BNDesc:
        d = BNDesc(prefix, Cn)
        d.goff = self._param(prefix + ".weight", (Cn,))
        d.boff = self._param(prefix + ".bias", (Cn,))
        d.roff = self._roff
        self._roff += (2 * Cn + ALIGN - 1) // ALIGN * ALIGN
        d.soff = self._soff
        self._soff += (2 * Cn + ALIGN - 1) // ALIGN * ALIGN
        d.index = len(self.bns)
        self.bns.append(d)
        return d

    def param_names(self) -> List[str]:
        return [n for n, _, _ in self.params]

    def buffer_names(self) -> List[str]:
        out = []
        for b in self.bns:
            out += [b.prefix + ".running_mean", b.prefix + ".running_var", b.prefix + ".num_batches_tracked"]
        return out


def _torch_dtype(dtype: int):
    return torch.float32 if dtype == _lib.F32 else torch.bfloat16


class Shadow:
    """Storage-dtype copy of a WeightSet's parameter arena (same element offsets).  fp32 kernels read the
    master arena itself; the bf16 copy is one flat conversion launch (ssb_weight_shadow)."""

    def __init__(self, w: "WeightSet", dtype: int):
        self.w = w
        self.dtype = dtype
        self.buf = torch.zeros(w.layout.n_params, dtype=torch.bfloat16, device=w.device) if dtype == _lib.BF16 else None

    def refresh(self, stream: int) -> None:
        if self.buf is not None:
            call("ssb_weight_shadow", self.w.params.data_ptr(), self.buf.data_ptr(), self.w.params.numel(), self.dtype, stream)

    def ptr(self, c: ConvDesc) -> int:
        if self.buf is None:
            return self.w.params.data_ptr() + 4 * c.poff
        return self.buf.data_ptr() + 2 * c.poff


class WeightSet:
    """One set of network weights in HBM: fp32 master arena (GEMM-conv weights tap-major), BN buffers,
    storage-dtype shadow copies."""

    def __init__(self, layout: ParamLayout, device, nbt_float: bool = False):
        self.layout = layout
        self.device = torch.device(device)
        self.params = torch.zeros(layout.n_params, dtype=torch.float32, device=device)
        self.bufs = torch.zeros(layout.n_bufs, dtype=torch.float32, device=device)
        # student: int64 counters; EMA teacher: float32 (mean_teacher.py:145-149 promotes them)
        self.nbt = torch.zeros(len(layout.bns), dtype=torch.float32 if nbt_float else torch.int64, device=device)
        for b in layout.bns:  # fresh BatchNorm: running_var = 1
            self.bufs[b.roff + b.C: b.roff + 2 * b.C] = 1.0
        self._shadows: Dict[int, Shadow] = {}

    def shadow(self, dtype: int) -> Shadow:
        if dtype not in self._shadows:
            self._shadows[dtype] = Shadow(self, dtype)
        return self._shadows[dtype]

    # ---- views ---------------------------------------------------------------------
    def param_view(self, name: str) -> torch.Tensor:
        for n, shape, off in self.layout.params:
            if n == name:
                return self.layout.view(self.params, n, shape, off)
        raise KeyError(name)

    def param_views(self, arena: Optional[torch.Tensor] = None) -> Dict[str, torch.Tensor]:
        a = self.params if arena is None else arena
        return {n: self.layout.view(a, n, s, off) for n, s, off in self.layout.params}

    def buffer_views(self) -> Dict[str, torch.Tensor]:
        out = {}
        for b in self.layout.bns:
            out[b.prefix + ".running_mean"] = self.bufs[b.roff: b.roff + b.C]
            out[b.prefix + ".running_var"] = self.bufs[b.roff + b.C: b.roff + 2 * b.C]
            out[b.prefix + ".num_batches_tracked"] = self.nbt[b.index]
        return out

    def state_dict(self) -> Dict[str, torch.Tensor]:
        """Reference-ordered state_dict (128 entries for resnet18+FCNHead)."""
        out: Dict[str, torch.Tensor] = {}
        pv, bv = self.param_views(), self.buffer_views()
        bn_by_prefix = {b.prefix: b for b in self.layout.bns}
        for n, _, _ in self.layout.params:
            out[n] = pv[n]
            if n.endswith(".bias") and n[:-5] in bn_by_prefix:
                p = n[:-5]
                for suffix in (".running_mean", ".running_var", ".num_batches_tracked"):
                    out[p + suffix] = bv[p + suffix]
        return out

    def load_state_dict(self, sd: Dict[str, torch.Tensor]) -> None:
        mine = self.state_dict()
        missing = [k for k in mine if k not in sd]
        unexpected = [k for k in sd if k not in mine]
        if missing or unexpected:
            raise RuntimeError(f"state_dict mismatch: missing {missing[:4]}..., unexpected {unexpected[:4]}...")
        with torch.no_grad():
            for k, v in mine.items():
                src = sd[k]
                if tuple(src.shape) != tuple(v.shape):
                    raise RuntimeError(f"shape mismatch for {k}: {tuple(src.shape)} vs {tuple(v.shape)}")
                v.copy_(src.to(device=v.device, dtype=v.dtype))

    # ---- pointers ------------------------------------------------------------------
    def w_ptr(self, c: ConvDesc) -> int:
        return self.params.data_ptr() + 4 * c.poff


class TrainState:
    """Gradient / optimizer arenas and BN statistic scratch for one trainable WeightSet."""

    def __init__(self, w: WeightSet):
        lay, dev = w.layout, w.device
        # everything a step zeroes before it starts lives in ONE arena (one memset node per step):
        # [gradients | BN forward sums | BN backward sums | loss sums] -- the tail is fp64
        n_tail = 2 * lay.n_sums + 64 + len(lay.bns) + 8 * lay.n_sums
        self.zero_arena = torch.zeros(lay.n_params + 2 * n_tail, dtype=torch.float32, device=dev)
        self.grads = self.zero_arena[: lay.n_params]
        tail = self.zero_arena[lay.n_params:].view(torch.float64)
        self.sums = tail[: lay.n_sums]
        self.bwd_sums = tail[lay.n_sums: 2 * lay.n_sums]
        self.loss_sums = tail[2 * lay.n_sums: 2 * lay.n_sums + 4]
        # one grid-barrier counter per BN layer (fused BN backward), zeroed with everything else
        nb = (len(lay.bns) + 1) // 2
        self.barriers = tail[2 * lay.n_sums + 8: 2 * lay.n_sums + 8 + nb].view(torch.int32)
        self.fwd_barriers = tail[2 * lay.n_sums + 8 + nb: 2 * lay.n_sums + 8 + 2 * nb].view(torch.int32)
        # 8 replicas of the BN backward sums (fused BN backward: block b accumulates into replica b % 8)
        r0_ = 2 * lay.n_sums + 64 + len(lay.bns)
        self.bwd_rep = tail[r0_: r0_ + 8 * lay.n_sums]
        self.exp_avg = torch.zeros(lay.n_params, dtype=torch.float32, device=dev)
        self.exp_avg_sq = torch.zeros(lay.n_params, dtype=torch.float32, device=dev)
        self.step = 0


class NetPlan:
    """Static buffers + kernel schedule for one (weights, batch, length, dtype, mode)."""

    def __init__(self, weights: WeightSet, dtype: int, B: int, L: int, train: bool, algo: Optional[int] = None,
                 grads: Optional[torch.Tensor] = None, sp_ptr: int = 0, bufs: Optional[torch.Tensor] = None,
                 wgrad_stream: Optional[torch.cuda.Stream] = None, state: Optional["TrainState"] = None,
                 dgrad_stream: Optional[torch.cuda.Stream] = None, eval_rows: int = 0,
                 eval_bufs: Optional[torch.Tensor] = None):
        _lib.prepare()
        self.w = weights
        self.sh = weights.shadow(dtype)
        self.lay = weights.layout
        self.spec = self.lay.spec
        self.B, self.L, self.train = B, L, train
        # FixMatch: `eval_rows` extra samples ride behind the B train samples in every activation tensor; they
        # take the eval-mode (running statistics `eval_bufs`) path of the SAME convs in the same launches
        self.eval_rows = eval_rows
        self.Ball = B + eval_rows
        if eval_rows and (not train or eval_bufs is None):
            raise ValueError("eval rows ride in a TRAINING plan and need the running-statistics arena they read")
        self.dtype = dtype
        self.device = weights.device
        self.tdt = _torch_dtype(self.dtype)
        self.grads = grads
        self.sp_ptr = sp_ptr
        # BN running-stat arena this plan reads/updates (an eval plan may be pointed at a snapshot)
        self.bufs = bufs if bufs is not None else weights.bufs
        # weight-gradient GEMMs are off the critical path of backward: issue them on a second stream
        self.wgrad_stream = wgrad_stream
        self.dgrad_stream = dgrad_stream   # branch for the shortcut convs' dgrad (off the main backward chain)
        self._pending_reads: Dict[int, torch.cuda.Event] = {}
        self.drop_mask_ptr = 0   # tests may inject an explicit keep-mask (u8 [B, Lh, Ch])
        self.debug = None        # tests: dict that receives clones of the block-output gradients
        self.sync_hook = None    # SyncBN: callable(tensor) all-reducing a statistics slice in place
        self.sync_fused = False  # SyncBN: the kernels that consume the statistics exchange them (ssb_bn.sync_*), no hook
        self.pre_block_event = None   # event the stream waits on after the stem (storage-dtype weight copy ready)
        # BN-backward reduce in the dgrad epilogue (ssb_conv1d_dgrad_bnred): 13 launches fewer per step, but measured
        # SLOWER (0.779 vs 0.765 ms at 16+16 x 2500; conv family +37 % at width 128): the extra loads, transposes and
        # per-chunk barriers sit on the epilogue, which is already the long pole of these tiles.  Off by default.
        self.fuse_reduce = bool(int(os.environ.get("SSB_FUSE_REDUCE", "0")))
        self.fuse_bn_bwd = bool(int(os.environ.get("SSB_FUSE_BN_BWD", "1")))   # reduce + apply in one launch (grid barrier)
        self._fused_ok: Dict[Tuple, bool] = {}
        # train-mode BN apply inside the conv launch (ssb_conv1d_fwd_bn_train: statistics -> grid barrier -> second pass
        # over the TMEM accumulator): 21 launches fewer, but measured SLOWER (0.781 vs 0.742 ms/step) -- the barrier plus
        # the second pass cost what the separate, PDL-overlapped BatchNorm launch cost, on the critical path.  Off.
        self.fuse_bn_fwd = bool(int(os.environ.get("SSB_FUSE_BN_FWD", "0")))
        self._bnf_ok: Dict[Tuple, bool] = {}
        self.block_done_hook = None   # callable(block_index) after a block's backward has been enqueued (bucketed all-reduce)
        if algo is None:
            algo = _lib.ALGO_TCGEN05 if self.dtype == _lib.BF16 else _lib.ALGO_SIMT
            if os.environ.get("SSB_FORCE_SIMT"):   # debugging aid: generic CUDA-core conv kernels everywhere
                algo = _lib.ALGO_SIMT
        self.algo = algo
        spec = self.spec
        if train and grads is None:
            raise ValueError("a training plan needs the gradient arena")
        for ch in (spec.stem_channels, spec.base_channels, spec.head_channels):
            if ch % 8:
                raise ValueError(f"channel counts must be multiples of 8, got {ch}")

        # ---- lengths and pitches (see include/ssb.h for the layout) ----
        L0 = conv_out_len(L, 7, 2, 3)
        Lp = conv_out_len(L0, 3, 2, 1)
        lens = []
        cur = Lp
        for s in spec.strides:
            cur = conv_out_len(cur, 3, s, 1)
            lens.append(cur)
        pitches = [0] * len(lens)
        pitches[-1] = lens[-1] + 2
        for i in range(len(lens) - 2, -1, -1):
            pitches[i] = pitches[i + 1] * spec.strides[i + 1]
            assert pitches[i] >= lens[i] + 2
        p_pool = pitches[0] * spec.strides[0]
        assert p_pool >= Lp + 2
        p_stem = 2 * p_pool
        assert p_stem >= L0 + 2
        self.g_stem = Geom(B, p_stem, L0, spec.stem_channels)
        self.g_pool = Geom(B, p_pool, Lp, spec.stem_channels)
        self.g_stage = [Geom(B, pitches[i], lens[i], spec.out_planes(i)) for i in range(len(lens))]
        # Bottleneck: conv2 / bn2 run at the stage's resolution with `planes` channels
        self.g_mid = [Geom(B, pitches[i], lens[i], spec.planes(i)) for i in range(len(lens))]
        self.g_head = Geom(B, pitches[-1], lens[-1], spec.head_channels)
        self.Lh = lens[-1]

        def act(g: Geom, C_: Optional[int] = None) -> torch.Tensor:
            return torch.zeros(self.Ball * g.pitch, C_ or g.C, dtype=self.tdt, device=self.device)

        def with_b(g: Geom, b: int) -> Geom:
            return Geom(b, g.pitch, g.len, g.C)
        self._with_b = with_b

        # ---- BN statistic arenas ----
        self.barriers = state.barriers if state is not None else None
        self.fwd_barriers = state.fwd_barriers if state is not None else None
        self.bwd_rep = state.bwd_rep if state is not None else None
        if state is not None:   # statistic arenas inside the step's zero arena (TrainState)
            self.sums, self.bwd_sums = state.sums, state.bwd_sums
        else:
            self.sums = torch.zeros(self.lay.n_sums, dtype=torch.float64, device=self.device)
            self.bwd_sums = torch.zeros(self.lay.n_sums, dtype=torch.float64, device=self.device)
        self.mean_invstd = torch.zeros(self.lay.n_sums, dtype=torch.float32, device=self.device)
        self._bn_structs: Dict[str, BN] = {}
        for b in self.lay.bns:
            self._bn_structs[b.prefix] = self._make_bn(b)

        # ---- activations ----
        self.x_in: Optional[torch.Tensor] = None  # [B, C, L] fp32, bound by the caller
        self.c0 = act(self.g_stem)
        self.p0 = act(self.g_pool)
        # max-pool routing slots saved by the train-mode stem tail for its backward pass
        self.pool_arg = torch.zeros(B * p_pool, spec.stem_channels, dtype=torch.uint8, device=self.device) if train else None
        self.blk_bufs: List[Dict[str, torch.Tensor]] = []
        gin = self.g_pool
        self.blk_geoms: List[Dict[str, Geom]] = []     # Bottleneck blocks: geometry of conv1's output (block input resolution)
        for bd in self.lay.blocks:
            gout = self.g_stage[bd.stage]
            if bd.conv3 is not None:
                g1 = Geom(B, gin.pitch, gin.len, bd.conv1.cout)
                gm = self.g_mid[bd.stage]
                bufs = {"c1": act(g1), "a1": act(g1), "c2": act(gm), "a2": act(gm), "c3": act(gout), "out": act(gout)}
                self.blk_geoms.append({"in": gin, "c1": g1, "mid": gm, "out": gout})
            else:
                bufs = {"c1": act(gout), "a1": act(gout), "c2": act(gout), "out": act(gout)}
                self.blk_geoms.append({"in": gin, "out": gout})
            if bd.convd is not None:
                bufs["cd"] = act(gout)
            self.blk_bufs.append(bufs)
            gin = gout
        self.ch = act(self.g_head)
        self.ah = act(self.g_head)
        self.low_all = torch.zeros(self.Ball, self.Lh, spec.num_classes, dtype=torch.float32, device=self.device)
        self.low = self.low_all[:B]
        self._bn_eval: Dict[str, BN] = {}
        if eval_rows:
            for b in self.lay.bns:   # same affine parameters, running statistics from the snapshot arena
                e = self._make_bn(b)
                e.running_mean = eval_bufs.data_ptr() + 4 * b.roff
                e.running_var = eval_bufs.data_ptr() + 4 * (b.roff + b.C)
                self._bn_eval[b.prefix] = e

        # ---- gradient scratch (per geometry) ----
        self._scratch: Dict[Tuple[int, int, int], Dict[str, torch.Tensor]] = {}
        if train:
            self.dlow = torch.zeros(B, self.Lh, spec.num_classes, dtype=torch.float32, device=self.device)
            self.dc0 = act(self.g_stem)
            extra = [g for bg_ in self.blk_geoms if "c1" in bg_ for g in (bg_["c1"], bg_["mid"])]
            for g in [self.g_pool, self.g_head] + self.g_stage + extra:
                key = (g.pitch, g.len, g.C)
                if key not in self._scratch:
                    self._scratch[key] = {n: act(g) for n in ("gA", "gE", "gB", "gC", "gD")}
            # the conv-output gradients (dc2, dc1, downsample) are ALSO read by the weight-gradient GEMMs on the
            # second stream: one buffer each per block, so that the main chain never waits for a wgrad to
            # release a buffer it wants to overwrite
            self.blk_grads: List[Dict[str, torch.Tensor]] = []
            for bd, bgm in zip(self.lay.blocks, self.blk_geoms):
                gout = self.g_stage[bd.stage]
                if bd.conv3 is not None:
                    bg = {"dc3": act(gout), "dc2": act(bgm["mid"]), "dc1": act(bgm["c1"])}
                else:
                    bg = {"dc2": act(gout), "dc1": act(gout)}
                if bd.convd is not None:
                    bg["dcd"] = act(gout)
                self.blk_grads.append(bg)
        self.launches_fwd = 0

    # ---- helpers -------------------------------------------------------------------
    def _make_bn(self, b: BNDesc) -> BN:
        w = self.w
        s = BN()
        s.gamma = w.params.data_ptr() + 4 * b.goff
        s.beta = w.params.data_ptr() + 4 * b.boff
        s.running_mean = self.bufs.data_ptr() + 4 * b.roff
        s.running_var = self.bufs.data_ptr() + 4 * (b.roff + b.C)
        s.num_batches_tracked = (w.nbt.data_ptr() + 8 * b.index) if w.nbt.dtype == torch.int64 else None
        s.sums = self.sums.data_ptr() + 8 * b.soff
        s.mean_invstd = self.mean_invstd.data_ptr() + 4 * b.soff
        s.bwd_sums = self.bwd_sums.data_ptr() + 8 * b.soff
        if self.grads is not None:
            s.dgamma = self.grads.data_ptr() + 4 * b.goff
            s.dbeta = self.grads.data_ptr() + 4 * b.boff
        return s

    def bn(self, b: BNDesc):
        return C.byref(self._bn_structs[b.prefix])

    def _g(self, c: ConvDesc) -> int:
        return self.grads.data_ptr() + 4 * c.poff

    def _conv_fwd(self, c: ConvDesc, x, y, gin: Geom, gout: Geom, st: int, stats: Optional[BNDesc] = None):
        """Conv1d; with `stats` the train-mode BatchNorm statistics of the output come out of the same
        launch (epilogue of the tcgen05 kernel; the generic path issues the statistics pass itself)."""
        if stats is None:
            call("ssb_conv1d_fwd", x.data_ptr(), self.sh.ptr(c), y.data_ptr(), gin, gout,
                 c.k, c.stride, self.dtype, self._algo_for(c), st)
            return
        call("ssb_conv1d_fwd_stats", x.data_ptr(), self.sh.ptr(c), y.data_ptr(), gin, gout,
             c.k, c.stride, self._bn_structs[stats.prefix].sums, self.dtype, self._algo_for(c), st)
        if self.sync_hook is not None:
            self.sync_hook(self.sums[stats.soff: stats.soff + 2 * stats.C])

    def _conv_bn_train(self, c: ConvDesc, b: BNDesc, x, y_raw, y_act, gin: Geom, gout: Geom, st: int, res=None,
                       b_res: Optional[BNDesc] = None) -> None:
        """train mode: y_raw = conv(x) (+ batch statistics), y_act = relu(bn(y_raw) [+ res | + bn_res(res)]).
        One launch when the conv's tiles are all co-resident (grid barrier + second pass over the TMEM accumulator),
        else conv(+statistics) followed by the BatchNorm pass."""
        key = (c.name, gout.B)
        if key not in self._bnf_ok:
            self._bnf_ok[key] = bool(self.fuse_bn_fwd and self.sync_hook is None and not self.sync_fused and self.barriers is not None and
                                     _lib.load().ssb_conv1d_fwd_bn_train_fits(gin, gout, c.k, c.stride, self.dtype, self._algo_for(c)))
        if self._bnf_ok[key]:
            call("ssb_conv1d_fwd_bn_train", x.data_ptr(), self.sh.ptr(c), y_raw.data_ptr(), y_act.data_ptr(), gin, gout, c.k,
                 c.stride, self.bn(b), res.data_ptr() if res is not None else None,
                 self.bn(b_res) if b_res is not None else None, 1, self.fwd_barriers.data_ptr() + 4 * b.index, self.dtype,
                 self._algo_for(c), st)
            return
        self._conv_fwd(c, x, y_raw, gin, gout, st, b)
        call("ssb_bn_act_fwd", y_raw.data_ptr(), self.bn(b), res.data_ptr() if res is not None else None,
             self.bn(b_res) if b_res is not None else None, y_act.data_ptr(), gout, 1, 1, self.dtype, st)

    def _conv_bn_act(self, c: ConvDesc, b: BNDesc, x, y, gin: Geom, gout: Geom, res, relu: int, st: int):
        """eval mode: y = [relu](bn_running(conv(x)) [+ res]) in one launch."""
        call("ssb_conv1d_bn_act_fwd", x.data_ptr(), self.sh.ptr(c), y.data_ptr(), gin, gout, c.k, c.stride, self.bn(b),
             res.data_ptr() if res is not None else None, relu, self.dtype, self._algo_for(c), st)

    def _algo_for(self, c: ConvDesc) -> int:
        if self.algo == _lib.ALGO_TCGEN05 and self.dtype == _lib.BF16 and c.cin % 64 == 0 and c.cout % 64 == 0:
            return _lib.ALGO_TCGEN05
        return _lib.ALGO_SIMT

    def _stats(self, buf, g: Geom, b: BNDesc, st: int):
        call("ssb_bn_stats", buf.data_ptr(), g, self._bn_structs[b.prefix].sums, self.dtype, st)
        if self.sync_hook is not None:
            self.sync_hook(self.sums[b.soff: b.soff + 2 * b.C])

    def _sync_bwd(self, first: BNDesc, last: Optional[BNDesc] = None):
        if self.sync_hook is not None:
            last = last or first
            self.sync_hook(self.bwd_sums[first.soff: last.soff + 2 * last.C])

    def _wgrad(self, c: ConvDesc, x, dy, gin: Geom, gout: Geom, st: int):
        """dW += x^T dy.  With a wgrad stream: fork after dy is produced, remember that dy is still
        being read so that the next writer of that scratch buffer waits (WAR)."""
        if "wgrad" in os.environ.get("SSB_DEBUG_SKIP", ""):   # timing experiments only: the step without its weight-gradient GEMMs
            return
        if self.wgrad_stream is None:
            call("ssb_conv1d_wgrad", x.data_ptr(), dy.data_ptr(), self._g(c), gin, gout, c.k, c.stride, self.dtype,
                 self._algo_for(c), st)
            return
        ready = torch.cuda.Event()
        ready.record()
        self.wgrad_stream.wait_event(ready)
        call("ssb_conv1d_wgrad", x.data_ptr(), dy.data_ptr(), self._g(c), gin, gout, c.k, c.stride, self.dtype,
             self._algo_for(c), self.wgrad_stream.cuda_stream)
        done = torch.cuda.Event()
        done.record(self.wgrad_stream)
        self._pending_reads[dy.data_ptr()] = done

    def _before_write(self, *bufs):
        for b in bufs:
            if b is None:
                continue
            ev = self._pending_reads.pop(b.data_ptr(), None)
            if ev is not None:
                torch.cuda.current_stream().wait_event(ev)

    def _join_wgrad(self):
        if self.wgrad_stream is not None:
            torch.cuda.current_stream().wait_stream(self.wgrad_stream)
            self._pending_reads.clear()

    def zero_stats(self, st: int):
        call("ssb_memset_zero", self.sums.data_ptr(), self.sums.numel() * 8, st)
        if self.train:
            call("ssb_memset_zero", self.bwd_sums.data_ptr(), self.bwd_sums.numel() * 8, st)
            if self.barriers is not None:   # (the backward and forward barrier counters are adjacent)
                call("ssb_memset_zero", self.barriers.data_ptr(), (self.barriers.numel() + self.fwd_barriers.numel()) * 4, st)
                call("ssb_memset_zero", self.bwd_rep.data_ptr(), self.bwd_rep.numel() * 8, st)

    # ---- forward -------------------------------------------------------------------
    def forward(self, x: torch.Tensor, st: int, train_mode: Optional[bool] = None, zero: bool = True,
                stream: Optional[torch.cuda.Stream] = None) -> torch.Tensor:
        """x: [B, C, L] fp32 contiguous CUDA tensor.  Returns low-res logits [B, Lh, ncls] fp32.
        train_mode: batch statistics + running-stat update + dropout (default: plan's mode).
        zero=False: the caller has already zeroed the statistic arenas (step engine: one memset)."""
        tm = self.train if train_mode is None else train_mode
        if tm and not self.train:
            raise ValueError("plan was built for eval mode")
        spec, lay, dt = self.spec, self.lay, self.dtype
        assert x.dtype == torch.float32 and x.is_contiguous() and tuple(x.shape) == (self.B, spec.num_leads, self.L), \
            f"input must be contiguous fp32 [{self.B},{spec.num_leads},{self.L}], got {tuple(x.shape)} {x.dtype}"
        self.x_in = x
        t = 1 if tm else 0
        if tm and zero:
            self.zero_stats(st)
        if tm:      # conv + the BatchNorm statistics of its output (one launch on the tensor-core path)
            call("ssb_stem_conv_fwd_stats", x.data_ptr(), self.w.w_ptr(lay.stem_conv), self.c0.data_ptr(), spec.num_leads,
                 self.L, self.g_stem, self._bn_structs[lay.stem_bn.prefix].sums, dt, st)
            if self.sync_hook is not None:
                self.sync_hook(self.sums[lay.stem_bn.soff: lay.stem_bn.soff + 2 * lay.stem_bn.C])
        else:
            call("ssb_stem_conv_fwd", x.data_ptr(), self.w.w_ptr(lay.stem_conv), self.c0.data_ptr(), spec.num_leads,
                 self.L, self.g_stem, dt, st)
        call("ssb_stem_bn_relu_pool_fwd", self.c0.data_ptr(), self.bn(lay.stem_bn), self.p0.data_ptr(),
             self.pool_arg.data_ptr() if tm else None, self.g_stem, self.g_pool, t, dt, st)
        h, gin = self.p0, self.g_pool
        if self.pre_block_event is not None:   # the stem reads the master weights; everything after it the storage-dtype copy
            (stream or torch.cuda.current_stream()).wait_event(self.pre_block_event)
        if not tm:
            # eval mode: BatchNorm is a fixed affine map -> folded, with the residual add and the ReLU, into
            # the conv epilogue (one launch per conv, no pre-activation tensors)
            for bd, bufs, bgm in zip(lay.blocks, self.blk_bufs, self.blk_geoms):
                gout = self.g_stage[bd.stage]
                if bd.conv3 is not None:     # Bottleneck (resnet.py:112-132)
                    self._conv_bn_act(bd.conv1, bd.bn1, h, bufs["a1"], gin, bgm["c1"], None, 1, st)
                    self._conv_bn_act(bd.conv2, bd.bn2, bufs["a1"], bufs["a2"], bgm["c1"], bgm["mid"], None, 1, st)
                    res = h
                    if bd.convd is not None:
                        self._conv_bn_act(bd.convd, bd.bnd, h, bufs["cd"], gin, gout, None, 0, st)
                        res = bufs["cd"]
                    self._conv_bn_act(bd.conv3, bd.bn3, bufs["a2"], bufs["out"], bgm["mid"], gout, res, 1, st)
                    h, gin = bufs["out"], gout
                    continue
                self._conv_bn_act(bd.conv1, bd.bn1, h, bufs["a1"], gin, gout, None, 1, st)
                res = h
                if bd.convd is not None:
                    self._conv_bn_act(bd.convd, bd.bnd, h, bufs["cd"], gin, gout, None, 0, st)
                    res = bufs["cd"]
                self._conv_bn_act(bd.conv2, bd.bn2, bufs["a1"], bufs["out"], gout, gout, res, 1, st)
                h, gin = bufs["out"], gout
            self.feat = h
            self._conv_bn_act(lay.head_conv, lay.head_bn, h, self.ah, gin, self.g_head, None, 1, st)
        else:
            aux = self.wgrad_stream if self.sync_hook is None else None   # idle during the forward
            for bd, bufs, bgm in zip(lay.blocks, self.blk_bufs, self.blk_geoms):
                gout = self.g_stage[bd.stage]
                if bd.conv3 is not None:     # Bottleneck: three conv + BN stages, the residual joins the third
                    self._conv_bn_train(bd.conv1, bd.bn1, h, bufs["c1"], bufs["a1"], gin, bgm["c1"], st)
                    self._conv_bn_train(bd.conv2, bd.bn2, bufs["a1"], bufs["c2"], bufs["a2"], bgm["c1"], bgm["mid"], st)
                    if bd.convd is not None:
                        self._conv_fwd(bd.convd, h, bufs["cd"], gin, gout, st, bd.bnd)
                        self._conv_bn_train(bd.conv3, bd.bn3, bufs["a2"], bufs["c3"], bufs["out"], bgm["mid"], gout, st,
                                            res=bufs["cd"], b_res=bd.bnd)
                    else:
                        self._conv_bn_train(bd.conv3, bd.bn3, bufs["a2"], bufs["c3"], bufs["out"], bgm["mid"], gout, st, res=h)
                    h, gin = bufs["out"], gout
                    continue
                side_done = None
                if bd.convd is not None and aux is not None:
                    # the 1x1 shortcut conv only needs the block input: own branch, joined before the residual add
                    fork = torch.cuda.Event()
                    fork.record()
                    aux.wait_event(fork)
                    self._conv_fwd(bd.convd, h, bufs["cd"], gin, gout, aux.cuda_stream, bd.bnd)
                    side_done = torch.cuda.Event()
                    side_done.record(aux)
                self._conv_bn_train(bd.conv1, bd.bn1, h, bufs["c1"], bufs["a1"], gin, gout, st)
                if bd.convd is not None:
                    if side_done is not None:
                        torch.cuda.current_stream().wait_event(side_done)
                    else:
                        self._conv_fwd(bd.convd, h, bufs["cd"], gin, gout, st, bd.bnd)
                    self._conv_bn_train(bd.conv2, bd.bn2, bufs["a1"], bufs["c2"], bufs["out"], gout, gout, st, res=bufs["cd"], b_res=bd.bnd)
                else:
                    self._conv_bn_train(bd.conv2, bd.bn2, bufs["a1"], bufs["c2"], bufs["out"], gout, gout, st, res=h)
                h, gin = bufs["out"], gout
            self.feat = h
            self._conv_bn_train(lay.head_conv, lay.head_bn, h, self.ch, self.ah, gin, self.g_head, st)
        p = spec.dropout_ratio if tm else 0.0
        call("ssb_head_cls_fwd", self.ah.data_ptr(), self.w.params.data_ptr() + 4 * lay.cls_w_off,
             self.w.params.data_ptr() + 4 * lay.cls_b_off, self.low.data_ptr(), self.g_head, spec.num_classes,
             p, self.drop_mask_ptr or None, self.sp_ptr or None, dt, st)
        return self.low

    def forward_merged(self, x_all: torch.Tensor, st: int):
        """Train-mode forward of the first B samples and eval-mode forward (running statistics of the snapshot
        arena) of the remaining `eval_rows` samples of x_all in ONE launch per conv (ssb_conv1d_fwd_dual): the
        pseudo-label pass of FixMatch (fixmatch.py:87-93) shares the student's weights, so it is just more rows.
        Returns (low_train [B, Lh, ncls], low_eval [eval_rows, Lh, ncls])."""
        spec, lay, dt, B, E = self.spec, self.lay, self.dtype, self.B, self.eval_rows
        assert E > 0 and x_all.dtype == torch.float32 and x_all.is_contiguous() and \
            tuple(x_all.shape) == (self.Ball, spec.num_leads, self.L)
        self.x_in = x_all[:B]
        es = 2 if dt == _lib.BF16 else 4
        wb = self._with_b

        def tail(buf: torch.Tensor, g: Geom) -> int:      # first eval row of a flat padded tensor
            return buf.data_ptr() + B * g.pitch * g.C * es

        def dual(c: ConvDesc, b: BNDesc, x, y_train, y_eval, gin, gout, res, relu):
            call("ssb_conv1d_fwd_dual", x.data_ptr(), self.sh.ptr(c), y_train.data_ptr(), y_eval.data_ptr(), wb(gin, self.Ball),
                 wb(gout, self.Ball), c.k, c.stride, B, self._bn_structs[b.prefix].sums, C.byref(self._bn_eval[b.prefix]),
                 res.data_ptr() if res is not None else None, relu, dt, self._algo_for(c), st)
            if self.sync_hook is not None:
                self.sync_hook(self.sums[b.soff: b.soff + 2 * b.C])

        call("ssb_stem_conv_fwd", x_all.data_ptr(), self.w.w_ptr(lay.stem_conv), self.c0.data_ptr(), spec.num_leads,
             self.L, wb(self.g_stem, self.Ball), dt, st)
        self._stats(self.c0, self.g_stem, lay.stem_bn, st)
        call("ssb_stem_bn_relu_pool_fwd", self.c0.data_ptr(), self.bn(lay.stem_bn), self.p0.data_ptr(),
             self.pool_arg.data_ptr(), self.g_stem, self.g_pool, 1, dt, st)
        call("ssb_stem_bn_relu_pool_fwd", tail(self.c0, self.g_stem), C.byref(self._bn_eval[lay.stem_bn.prefix]),
             tail(self.p0, self.g_pool), None, wb(self.g_stem, E), wb(self.g_pool, E), 0, dt, st)
        h, gin = self.p0, self.g_pool
        if self.pre_block_event is not None:
            torch.cuda.current_stream().wait_event(self.pre_block_event)
        for bd, bufs in zip(lay.blocks, self.blk_bufs):
            gout = self.g_stage[bd.stage]
            dual(bd.conv1, bd.bn1, h, bufs["c1"], bufs["a1"], gin, gout, None, 1)          # eval rows: a1 = relu(bn1(conv1))
            if bd.convd is not None:
                dual(bd.convd, bd.bnd, h, bufs["cd"], bufs["cd"], gin, gout, None, 0)      # eval rows: bnd(convd)
            call("ssb_bn_act_fwd", bufs["c1"].data_ptr(), self.bn(bd.bn1), None, None, bufs["a1"].data_ptr(), gout, 1, 1, dt, st)
            res = bufs["cd"] if bd.convd is not None else h
            dual(bd.conv2, bd.bn2, bufs["a1"], bufs["c2"], bufs["out"], gout, gout, res, 1)  # eval rows: relu(bn2(conv2) + res)
            if bd.convd is not None:
                call("ssb_bn_act_fwd", bufs["c2"].data_ptr(), self.bn(bd.bn2), bufs["cd"].data_ptr(), self.bn(bd.bnd),
                     bufs["out"].data_ptr(), gout, 1, 1, dt, st)
            else:
                call("ssb_bn_act_fwd", bufs["c2"].data_ptr(), self.bn(bd.bn2), h.data_ptr(), None,
                     bufs["out"].data_ptr(), gout, 1, 1, dt, st)
            h, gin = bufs["out"], gout
        self.feat = h
        dual(lay.head_conv, lay.head_bn, h, self.ch, self.ah, gin, self.g_head, None, 1)
        call("ssb_bn_act_fwd", self.ch.data_ptr(), self.bn(lay.head_bn), None, None, self.ah.data_ptr(), self.g_head, 1, 1, dt, st)
        wp = self.w.params.data_ptr()
        call("ssb_head_cls_fwd", self.ah.data_ptr(), wp + 4 * lay.cls_w_off, wp + 4 * lay.cls_b_off, self.low_all.data_ptr(),
             self.g_head, spec.num_classes, spec.dropout_ratio, self.drop_mask_ptr or None, self.sp_ptr or None, dt, st)
        call("ssb_head_cls_fwd", tail(self.ah, self.g_head), wp + 4 * lay.cls_w_off, wp + 4 * lay.cls_b_off,
             self.low_all[B:].data_ptr(), wb(self.g_head, E), spec.num_classes, 0.0, None, None, dt, st)
        return self.low, self.low_all[B:]

    def _dgrad(self, c: ConvDesc, dy, dx, gin: Geom, gout: Geom, acc: int, st: int, red=None):
        """dx (+)= conv_transpose(dy, w).  red = (y_act, x_pre, bn[, x_res, bn_res]): the BatchNorm-backward reduce of the
        gradient being produced rides in the epilogue (stride-1 convs), replacing a separate reduce launch."""
        if red is None or c.stride != 1:
            call("ssb_conv1d_dgrad", dy.data_ptr(), self.sh.ptr(c), dx.data_ptr(), gin, gout, c.k, c.stride, acc, self.dtype,
                 self._algo_for(c), st)
            return False
        y_act, x_pre, b = red[0], red[1], red[2]
        x_res = red[3] if len(red) > 3 else None
        b_res = red[4] if len(red) > 3 else None
        call("ssb_conv1d_dgrad_bnred", dy.data_ptr(), self.sh.ptr(c), dx.data_ptr(), gin, gout, c.k, c.stride, acc,
             y_act.data_ptr(), x_pre.data_ptr(), self.bn(b), x_res.data_ptr() if x_res is not None else None,
             self.bn(b_res) if b_res is not None else None, self.dtype, self._algo_for(c), st)
        return True

    def _bn_bwd(self, g1, y, x, b: BNDesc, dx, geom: Geom, st: int, x_res=None, b_res: Optional[BNDesc] = None, dx_res=None,
                g_ident=None, pre_reduced: bool = False):
        """BatchNorm(+ReLU) backward of one layer: both passes in one launch (grid barrier) when the tensor is small
        enough for the whole grid to be co-resident and no statistics exchange sits between them, else two launches."""
        ptr = lambda t: t.data_ptr() if t is not None else None   # noqa: E731
        if self.fuse_bn_bwd and not pre_reduced and self.sync_hook is None and self.barriers is not None:
            mode = 2 if b_res is not None else (1 if g_ident is not None else 0)
            key = (geom.B, geom.pitch, geom.len, geom.C, mode, y is not None)
            if key not in self._fused_ok:
                self._fused_ok[key] = bool(_lib.load().ssb_bn_bwd_fused_fits(geom, mode, 1 if y is not None else 0, self.dtype))
            if self._fused_ok[key]:
                call("ssb_bn_bwd_fused", g1.data_ptr(), ptr(y), x.data_ptr(), self.bn(b), dx.data_ptr(), ptr(x_res),
                     self.bn(b_res) if b_res is not None else None, ptr(dx_res), ptr(g_ident), geom,
                     self.barriers.data_ptr() + 4 * b.index, self.bwd_rep.data_ptr() + 8 * b.soff,
                     (self.bwd_rep.data_ptr() + 8 * b_res.soff) if b_res is not None else None, self.lay.n_sums, self.dtype, st)
                return
        if not pre_reduced:
            call("ssb_bn_bwd_reduce", g1.data_ptr(), None, ptr(y), x.data_ptr(), self.bn(b), ptr(x_res),
                 self.bn(b_res) if b_res is not None else None, geom, self.dtype, st)
        self._sync_bwd(b, b_res)
        call("ssb_bn_bwd_apply", g1.data_ptr(), None, ptr(y), x.data_ptr(), self.bn(b), dx.data_ptr(), ptr(x_res),
             self.bn(b_res) if b_res is not None else None, ptr(dx_res), ptr(g_ident), geom, self.dtype, st)

    def _bn2_red(self, bi: int):
        """reduce arguments of block bi's output BN (bn2 [+ downsample BN]) for the dgrad that produces its gradient"""
        bd, bufs = self.lay.blocks[bi], self.blk_bufs[bi]
        if bd.convd is not None:
            return (bufs["out"], bufs["c2"], bd.bn2, bufs["cd"], bd.bnd)
        return (bufs["out"], bufs["c2"], bd.bn2)

    # ---- backward ------------------------------------------------------------------
    def backward(self, dlow: torch.Tensor, st: int) -> None:
        """dlow: gradient w.r.t. the low-res logits [B, Lh, ncls] fp32.  Accumulates weight
        gradients into the gradient arena (the caller zeroes it once per step)."""
        assert self.train
        spec, lay, dt = self.spec, self.lay, self.dtype
        x = self.x_in
        sc_h = self._scratch[(self.g_head.pitch, self.g_head.len, self.g_head.C)]
        p = spec.dropout_ratio
        gw = self.grads.data_ptr()
        call("ssb_head_cls_bwd", dlow.data_ptr(), self.ah.data_ptr(), self.w.params.data_ptr() + 4 * lay.cls_w_off,
             sc_h["gA"].data_ptr(), gw + 4 * lay.cls_w_off, gw + 4 * lay.cls_b_off, self.g_head, spec.num_classes, p,
             self.drop_mask_ptr or None, self.sp_ptr or None, dt, st)
        # head BN + ReLU backward -> dch (gB)
        self._bn_bwd(sc_h["gA"], self.ah, self.ch, lay.head_bn, sc_h["gB"], self.g_head, st)
        gfeat = self.g_stage[-1]
        sc_f = self._scratch[(gfeat.pitch, gfeat.len, gfeat.C)]
        hc = lay.head_conv
        G = sc_f["gA"]
        # (the weight-gradient GEMM only needs dy: fork it BEFORE the dgrad so that the two run side by side)
        self._wgrad(hc, self.feat, sc_h["gB"], gfeat, self.g_head, st)
        fuse = self.fuse_reduce
        pre_reduced = self._dgrad(hc, sc_h["gB"], G, gfeat, self.g_head, 0, st, self._bn2_red(len(lay.blocks) - 1) if fuse else None)

        # blocks in reverse
        nblk = len(lay.blocks)
        for bi in range(nblk - 1, -1, -1):
            bd, bufs = lay.blocks[bi], self.blk_bufs[bi]
            gout = self.g_stage[bd.stage]
            if bi == 0:
                xin, gin = self.p0, self.g_pool
            else:
                xin, gin = self.blk_bufs[bi - 1]["out"], self.g_stage[lay.blocks[bi - 1].stage]
            sc = self._scratch[(gout.pitch, gout.len, gout.C)]
            sci = self._scratch[(gin.pitch, gin.len, gin.C)]
            if self.debug is not None:
                self.debug[bd.prefix] = self.to_ncl(G, gout)
            # destination for the gradient w.r.t. the block input
            if sci is sc:
                Gin = sc["gE"] if G is sc["gA"] else sc["gA"]
            else:
                Gin = sci["gA"]
            bg = self.blk_grads[bi]
            if bd.conv3 is not None:
                self._backward_bottleneck(bi, G, Gin, xin, gin, st)
                G = Gin
                pre_reduced = False
                if self.block_done_hook is not None:
                    self.block_done_hook(bi)
                continue
            dc2, dcd, da1 = bg["dc2"], bg.get("dcd"), sc["gD"]
            out, c2 = bufs["out"], bufs["c2"]
            self._before_write(dc2, dcd)
            if bd.convd is not None:
                self._bn_bwd(G, out, c2, bd.bn2, dc2, gout, st, x_res=bufs["cd"], b_res=bd.bnd, dx_res=dcd, pre_reduced=pre_reduced)
            else:
                self._bn_bwd(G, out, c2, bd.bn2, dc2, gout, st, g_ident=Gin, pre_reduced=pre_reduced)
            if self.debug is not None:
                self.debug[bd.prefix + ".conv2"] = self.to_ncl(dc2, gout)
                if bd.convd is not None:
                    self.debug[bd.prefix + ".downsample.0"] = self.to_ncl(dcd, gout)
            side_dgrad = None
            if bd.convd is not None and self.dgrad_stream is not None:
                # shortcut dgrad (dcd -> Gin, plain store) on its own branch, under the conv2-dgrad / bn1 chain
                fork = torch.cuda.Event()
                fork.record()
                self.dgrad_stream.wait_event(fork)
                cdn = bd.convd
                call("ssb_conv1d_dgrad", dcd.data_ptr(), self.sh.ptr(cdn), Gin.data_ptr(), gin, gout,
                     cdn.k, cdn.stride, 0, dt, self._algo_for(cdn), self.dgrad_stream.cuda_stream)
                side_dgrad = torch.cuda.Event()
                side_dgrad.record(self.dgrad_stream)
            c = bd.conv2
            self._wgrad(c, bufs["a1"], dc2, gout, gout, st)
            red1 = self._dgrad(c, dc2, da1, gout, gout, 0, st, (bufs["a1"], bufs["c1"], bd.bn1) if fuse else None)
            # bn1 + relu backward -> dc1
            dc1 = bg["dc1"]
            self._before_write(dc1)
            self._bn_bwd(da1, bufs["a1"], bufs["c1"], bd.bn1, dc1, gout, st, pre_reduced=red1)
            if self.debug is not None:
                self.debug[bd.prefix + ".conv1"] = self.to_ncl(dc1, gout)
            c = bd.conv1
            acc = 0 if bd.convd is not None else 1   # identity residual: Gin already holds g
            self._wgrad(c, xin, dc1, gin, gout, st)
            if bd.convd is not None:
                self._wgrad(bd.convd, xin, dcd, gin, gout, st)
            if side_dgrad is not None:
                torch.cuda.current_stream().wait_event(side_dgrad)   # Gin already holds the shortcut's gradient
                acc = 1
            # this launch finalises the gradient of the previous block's output when the block has no shortcut conv
            nxt = self._bn2_red(bi - 1) if (fuse and bi > 0 and bd.convd is None) else None
            pre_reduced = self._dgrad(c, dc1, Gin, gin, gout, acc, st, nxt)
            if bd.convd is not None and side_dgrad is None:
                c = bd.convd
                call("ssb_conv1d_dgrad", dcd.data_ptr(), self.sh.ptr(c), Gin.data_ptr(), gin, gout,
                     c.k, c.stride, 1, dt, self._algo_for(c), st)
            G = Gin
            if self.block_done_hook is not None:
                self.block_done_hook(bi)
        # stem tail + stem conv weight gradient (independent of the conv weight gradients still in flight)
        call("ssb_stem_bwd_reduce", G.data_ptr(), self.c0.data_ptr(), self.pool_arg.data_ptr(), self.bn(lay.stem_bn),
             self.g_stem, self.g_pool, dt, st)
        self._sync_bwd(lay.stem_bn)
        call("ssb_stem_bwd_apply", G.data_ptr(), self.c0.data_ptr(), self.pool_arg.data_ptr(), self.bn(lay.stem_bn),
             self.dc0.data_ptr(), self.g_stem, self.g_pool, dt, st)
        call("ssb_stem_conv_wgrad", x.data_ptr(), self.dc0.data_ptr(), self._g(lay.stem_conv), spec.num_leads, self.L,
             self.g_stem, dt, st)
        self._join_wgrad()

    def _backward_bottleneck(self, bi: int, G, Gin, xin, gin: Geom, st: int) -> None:
        """Backward of one Bottleneck block (reference resnet.py:112-132): G = gradient w.r.t. the block output; leaves the
        gradient w.r.t. the block input in Gin.  Same pieces as the BasicBlock path -- fused BN backward, dgrad on the
        main chain, weight-gradient GEMMs forked onto the second stream -- in a plain serial order."""
        bd, bufs, bgm, bg = self.lay.blocks[bi], self.blk_bufs[bi], self.blk_geoms[bi], self.blk_grads[bi]
        dt = self.dtype
        g1, gm, gout = bgm["c1"], bgm["mid"], bgm["out"]
        dc3, dc2, dc1, dcd = bg["dc3"], bg["dc2"], bg["dc1"], bg.get("dcd")
        da2 = self._scratch[(gm.pitch, gm.len, gm.C)]["gD"]
        da1 = self._scratch[(g1.pitch, g1.len, g1.C)]["gD"]
        self._before_write(dc3, dcd)
        if bd.convd is not None:
            self._bn_bwd(G, bufs["out"], bufs["c3"], bd.bn3, dc3, gout, st, x_res=bufs["cd"], b_res=bd.bnd, dx_res=dcd)
        else:
            self._bn_bwd(G, bufs["out"], bufs["c3"], bd.bn3, dc3, gout, st, g_ident=Gin)
        if self.debug is not None:
            self.debug[bd.prefix + ".conv3"] = self.to_ncl(dc3, gout)
        self._wgrad(bd.conv3, bufs["a2"], dc3, gm, gout, st)
        self._dgrad(bd.conv3, dc3, da2, gm, gout, 0, st)
        self._before_write(dc2)
        self._bn_bwd(da2, bufs["a2"], bufs["c2"], bd.bn2, dc2, gm, st)
        if self.debug is not None:
            self.debug[bd.prefix + ".conv2"] = self.to_ncl(dc2, gm)
        self._wgrad(bd.conv2, bufs["a1"], dc2, g1, gm, st)
        self._dgrad(bd.conv2, dc2, da1, g1, gm, 0, st)
        self._before_write(dc1)
        self._bn_bwd(da1, bufs["a1"], bufs["c1"], bd.bn1, dc1, g1, st)
        if self.debug is not None:
            self.debug[bd.prefix + ".conv1"] = self.to_ncl(dc1, g1)
        self._wgrad(bd.conv1, xin, dc1, gin, g1, st)
        if bd.convd is not None:
            self._wgrad(bd.convd, xin, dcd, gin, gout, st)
            call("ssb_conv1d_dgrad", dcd.data_ptr(), self.sh.ptr(bd.convd), Gin.data_ptr(), gin, gout,
                 bd.convd.k, bd.convd.stride, 0, dt, self._algo_for(bd.convd), st)
        self._dgrad(bd.conv1, dc1, Gin, gin, g1, 1, st)     # Gin already holds the identity / shortcut gradient

    # ---- test helpers: NCL fp32 copies of internal tensors ---------------------------
    def to_ncl(self, buf: torch.Tensor, g: Geom) -> torch.Tensor:
        v = buf[: g.B * g.pitch].view(g.B, g.pitch, -1)[:, 1: 1 + g.len, :]
        return v.permute(0, 2, 1).float().contiguous()
