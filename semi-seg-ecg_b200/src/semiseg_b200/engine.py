"""Step engine: the semi-supervised training step as one static CUDA-graph of hand-written kernels.

Replaces the bodies of the reference's hot loops (file:line relative to the reference tree):
  fixmatch.train_one_epoch      src/algorithms/fixmatch.py:73-155
  mean_teacher.train_one_epoch  src/algorithms/mean_teacher.py:76-164
  base.train_one_epoch          src/algorithms/base.py:110-159
  NativeScaler / AdamW / EMA    src/utils/misc.py:242-256, src/utils/optimizer.py:22-34,
                                src/algorithms/mean_teacher.py:139-149

One step = [zero grads] -> bf16 weight copy -> pseudo-label forward (self-eval or EMA teacher)
-> student forward (batch statistics, dropout) -> fused upsample+softmax+threshold+argmax+
loss+gradient -> backward -> [NCCL gradient all-reduce] -> fused AdamW(+EMA).  All buffers are
static; scalars that change per step (lr, bias corrections, RNG counter) live in a 64-byte
device struct refreshed by one H2D copy, so the whole step replays as a CUDA graph with no
host synchronisation.  Loss sums are read back asynchronously.
"""
from __future__ import annotations

import ctypes as C
import math
import os
from typing import Dict, List, Optional

import torch

from . import _lib
from ._lib import StepParams, call
from .net import NetPlan, ParamLayout, SegNetSpec, TrainState, WeightSet

# cps / stpp: hard pseudo-labels argmax(teacher(u_w)) from ANOTHER weight set (the peer
# model of Cross Pseudo Supervision, cps.py:96-134; the frozen teacher of ST++, stpp.py:150-178), plain CE over every
# position -- the FixMatch loss mode with threshold 0 (conf = max softmax >= 1/ncls > 0 keeps every position) -- and the
# student sees the weak view itself (torch.cat((ecg_x, ecg_u_w))).
ALGOS = {"supervised": _lib.LOSS_SUP, "fixmatch": _lib.LOSS_FIXMATCH, "mean_teacher": _lib.LOSS_SOFT,
         "cps": _lib.LOSS_FIXMATCH, "stpp": _lib.LOSS_FIXMATCH}
HARD_TEACHER = ("cps", "stpp")


def _stream() -> int:
    return torch.cuda.current_stream().cuda_stream


class StepEngine:
    def __init__(self, weights: WeightSet, state: TrainState, dtype: int, algorithm: str, Bl: int, Bu: int,
                 L: int, train_cfg: dict, teacher: Optional[WeightSet] = None, algo: Optional[int] = None,
                 use_graph: bool = True, process_group=None, sync_bn: bool = False, seed: int = 0,
                 materialize: bool = False, external_pseudo: bool = False):
        if algorithm not in ALGOS:
            raise ValueError(f"unknown algorithm {algorithm!r} (supported: {sorted(ALGOS)})")
        if weights.device.type != "cuda":
            raise RuntimeError("StepEngine needs CUDA tensors: the hot path has no CPU fallback")
        _lib.check(_lib.load().ssb_device_check(), "ssb_device_check")
        self.w, self.state, self.algorithm = weights, state, algorithm
        self.mode = ALGOS[algorithm]
        self.hard_teacher = algorithm in HARD_TEACHER
        # external_pseudo: the pseudo-label forward is NOT part of step(); the owner calls pseudo() first (CpsEngine: both
        # peers' pseudo-labels are taken before either model is updated, cps.py:96-103)
        self.external_pseudo = external_pseudo
        if external_pseudo and not self.hard_teacher:
            raise ValueError("external_pseudo is for the hard-teacher algorithms (cps, stpp)")
        self.Bl, self.Bu, self.L = Bl, (Bu if algorithm != "supervised" else 0), L
        self.cfg = train_cfg
        self.spec: SegNetSpec = weights.layout.spec
        self.device = weights.device
        self.use_graph = use_graph
        self.pg = process_group
        self.world = torch.distributed.get_world_size(process_group) if process_group is not None else 1
        # SSB_FORCE_COLLECTIVES=1: issue the collectives even at world size 1 (exercises the capture path on one GPU)
        self.collectives = process_group is not None and (self.world > 1 or bool(os.environ.get("SSB_FORCE_COLLECTIVES")))
        self.sync_bn = sync_bn and self.collectives
        if self.collectives:
            # create the NCCL communicator OUTSIDE of any stream capture (lazy init inside a capture deadlocks)
            warm = torch.zeros(8, device=weights.device)
            torch.distributed.all_reduce(warm, group=process_group)
            warm64 = torch.zeros(8, dtype=torch.float64, device=weights.device)
            torch.distributed.all_reduce(warm64, group=process_group)
            torch.cuda.synchronize()
        self.seed = seed
        dev = self.device
        S = self.Bl + self.Bu
        Cl = self.spec.num_leads
        # static input arena: labeled rows first, strong-aug rows after (replaces torch.cat, fixmatch.py:99)
        # FixMatch: the weak views sit right behind the student batch, so that the pseudo-label forward is the
        # tail rows of the student's own conv launches (NetPlan.forward_merged)
        # (measured: with the multi-branch graph the separate pseudo-label branch overlaps the student forward and
        # wins, 0.76 vs 0.80 ms; in single-stream mode the merged launches win, 0.89 vs 1.00 ms)
        # train.pseudo_dtype: "fp32" runs the pseudo-label / teacher forward through the FP32 kernels even when the
        # training step is bf16 -- what the reference does (that forward sits outside autocast: fixmatch.py:87-91,
        # mean_teacher.py:89-91, cps.py:95-101).  Default: the training dtype (bf16 on the tcgen05 path; its decisions
        # differ from the fp32 ones only within 2e-2 of the threshold / of a tie: tests/test_bench_shapes_gpu.py).
        pd = str(train_cfg.get("pseudo_dtype", "") or "").lower()
        if pd not in ("", "same", "fp32", "float32", "bf16"):
            raise ValueError(f"train.pseudo_dtype={pd!r}: expected 'fp32' or 'bf16'")
        self.dtype_t = _lib.F32 if pd in ("fp32", "float32") else dtype
        self.merged = algorithm == "fixmatch" and not weights.layout.spec.bottleneck and self.dtype_t == dtype and bool(int(os.environ.get(
            "SSB_MERGED_EVAL", "0" if int(os.environ.get("SSB_MULTI_STREAM", "1")) else "1")))
        self.x_all = torch.zeros(S + max(self.Bu, 1), Cl, L, dtype=torch.float32, device=dev)
        self.x_s = self.x_all[:S]
        self.y_l = torch.zeros(Bl, L, dtype=torch.int64, device=dev)
        self.x_uw = self.x_all[S:]
        # per-step scalars
        self.sp_host = [torch.zeros(64, dtype=torch.uint8).pin_memory() for _ in range(8)]
        self.sp_events = [torch.cuda.Event() for _ in range(8)]
        self.sp_used = [False] * 8
        self.sp_dev = torch.zeros(64, dtype=torch.uint8, device=dev)
        self.loss_sums = state.loss_sums   # inside the step's zero arena
        self.stats_host = [torch.zeros(4, dtype=torch.float64).pin_memory() for _ in range(8)]
        self.stats_events = [torch.cuda.Event() for _ in range(8)]
        self._pending: List[int] = []
        self._done: List[Dict[str, float]] = []
        self.gnorm_ws = torch.zeros(1, dtype=torch.float64, device=dev)
        self.gnorm = torch.zeros(1, dtype=torch.float32, device=dev)
        # teacher
        self.teacher = teacher
        if (algorithm == "mean_teacher" or self.hard_teacher) and teacher is None:
            raise ValueError(f"{algorithm} needs a teacher WeightSet")
        self.uses_teacher_weights = algorithm == "mean_teacher" or self.hard_teacher
        self.ema_first = True
        # plans
        self.dtype = dtype
        # concurrency inside the step: weight-gradient GEMMs and the pseudo-label forward are off the
        # critical path -> own streams, forked/joined with events (captured as graph branches)
        self.multi_stream = bool(int(os.environ.get("SSB_MULTI_STREAM", "1")))
        self.wgrad_stream = torch.cuda.Stream(device=dev) if self.multi_stream else None
        self.teacher_stream = torch.cuda.Stream(device=dev) if self.multi_stream else None
        self.repack_stream = torch.cuda.Stream(device=dev) if self.multi_stream else None
        # gradient all-reduce in buckets: the ranges of the last stages (+ head) are exchanged on their own stream as
        # soon as those gradients exist, under the backward of the earlier stages (SSB_BUCKETS = how many, 0 = off)
        self.comm_stream = torch.cuda.Stream(device=dev) if (self.multi_stream and int(os.environ.get("SSB_BUCKETS", "2"))) else None
        self.dgrad_stream = torch.cuda.Stream(device=dev) if self.multi_stream else None
        self.bufs_snap: Optional[torch.Tensor] = None
        if self.merged:
            # the eval rows must see the running stats from BEFORE this step's update (fixmatch.py:87-93)
            self.bufs_snap = torch.empty_like(weights.bufs)
        self.plan_s = NetPlan(weights, dtype, S, L, True, algo, grads=state.grads, sp_ptr=self.sp_dev.data_ptr(),
                              wgrad_stream=self.wgrad_stream, state=state, dgrad_stream=self.dgrad_stream,
                              eval_rows=self.Bu if self.merged else 0, eval_bufs=self.bufs_snap)
        self.plan_t: Optional[NetPlan] = None
        if self.mode != _lib.LOSS_SUP and not self.merged:
            tw = teacher if self.uses_teacher_weights else weights
            if algorithm == "fixmatch" and self.multi_stream:
                # self-eval pass must see the running stats from BEFORE this step's update (fixmatch.py:87-93)
                self.bufs_snap = torch.empty_like(weights.bufs)
            self.plan_t = NetPlan(tw, self.dtype_t, self.Bu, L, False, algo if self.dtype_t == dtype else None, bufs=self.bufs_snap)
        if self.sync_bn:
            # SyncBatchNorm (fixmatch.py:290-291): the statistics of every BN layer are summed over the ranks
            for s in self.plan_s._bn_structs.values():
                s.count_mul = self.world
            if not self._setup_fused_sync():          # exchange inside the consuming kernels (NVLink peer mailboxes)
                self.plan_s.sync_hook = self._make_sync_hook()   # else: one exchange launch (or NCCL call) per BN layer
        self._stage = None
        self.stage_h2d = bool(int(os.environ.get("SSB_STAGE_H2D", "1")))
        self.mat = None
        self.pseudo_graph: Optional[torch.cuda.CUDAGraph] = None
        if materialize and self.mode == _lib.LOSS_FIXMATCH:
            self.mat = {"conf": torch.zeros(self.Bu, L, dtype=torch.float32, device=dev),
                        "label": torch.zeros(self.Bu, L, dtype=torch.int64, device=dev),
                        "mask": torch.zeros(self.Bu, L, dtype=torch.uint8, device=dev)}
        self.graph: Optional[torch.cuda.CUDAGraph] = None
        self.it = 0
        self.launches_per_step = 0
        kw = train_cfg.get("optimizer_kwargs", {}) or {}
        betas = kw.get("betas", (0.9, 0.999))
        self.beta1, self.beta2 = float(betas[0]), float(betas[1])
        self.eps = float(kw.get("eps", 1e-8))
        self.wd = float(train_cfg.get("weight_decay", 0.0))
        if train_cfg.get("optimizer", "adamw") != "adamw":
            raise NotImplementedError("the fused optimizer kernel implements AdamW (all shipped configs)")

    def _setup_fused_sync(self) -> bool:
        """SyncBN without exchange launches: every BN struct of the training plan gets the peers' mailbox addresses and
        its own slices (ssb_bn.sync_*); the kernels that consume the statistics publish this rank's sums to the peers and
        sum the ranks' values from their own mailbox.  A slice is used once per step and tagged with the step count, so
        the BN layers' exchanges need no common order: the shortcut convs may stay on their side branch.  False if the
        peer mapping cannot be set up (or SSB_SYNCBN_FUSED=0): the caller falls back to the per-layer exchange."""
        self.syncbn_p2p = False
        self.syncbn_fused = False
        # Default: on at world size 2 -- the configuration this path was run and checked on (tools/dp_check.py on two
        # B200s, bench at N = 2).  At N = 4 / 8 the per-layer exchange launches below are the path with hardware runs
        # behind them (profiles/r2b_multi_gpu.md); the sweep that was to cover the in-kernel exchange there did not
        # complete (profiles/r2e_multi_gpu.md), so larger worlds need SSB_SYNCBN_FUSED=1 to opt in.
        want = os.environ.get("SSB_SYNCBN_FUSED")
        on = (self.world == 2) if want is None else bool(int(want))
        if not on or not int(os.environ.get("SSB_SYNCBN_P2P", "1")) or self.world < 2 or self.world > 16 or self.merged:
            return False
        lay = self.plan_s.lay
        slot = 2 * lay.n_sums                      # forward sums | backward sums of every layer
        try:
            import torch.distributed._symmetric_memory as symm_mem
            nbytes = int(_lib.load().ssb_syncbn_fused_mailbox_bytes(slot, self.world))
            box = symm_mem.empty(nbytes, dtype=torch.uint8, device=self.device)
            box.zero_()
            hdl = symm_mem.rendezvous(box, self.pg if self.pg is not None else torch.distributed.group.WORLD)
            ptrs = [int(p_) for p_ in hdl.buffer_ptrs]
            assert len(ptrs) == self.world and ptrs[hdl.rank] == box.data_ptr()
            torch.cuda.synchronize()
            torch.distributed.barrier(group=self.pg)
        except Exception as e:   # no peer access / allocator not available
            import warnings
            warnings.warn(f"SyncBN in-kernel exchange unavailable ({type(e).__name__}: {e}); using one exchange per layer")
            return False
        self._mailbox, self._mail_hdl = box, hdl
        self._peers_dev = torch.tensor(ptrs, dtype=torch.int64, device=self.device)
        for b in lay.bns:
            s = self.plan_s._bn_structs[b.prefix]
            s.sync_peers = self._peers_dev.data_ptr()
            s.sync_sp = self.sp_dev.data_ptr()
            s.sync_world, s.sync_rank = self.world, hdl.rank
            s.sync_slot = slot
            s.sync_fwd_off, s.sync_bwd_off = b.soff, lay.n_sums + b.soff
        self.plan_s.sync_fused = True
        self.syncbn_p2p = True
        self.syncbn_fused = True
        return True

    def _make_sync_hook(self):
        """Statistics exchange of one SyncBN slice.  Preferred: our own kernel over NVLink peer memory (a symmetric
        mailbox mapped with torch's symmetric-memory allocator) -- one ~5 us launch instead of an ~11 us NCCL
        all-reduce, 42 times per step.  If the peer mapping cannot be set up: NCCL (SSB_SYNCBN_P2P=0 forces it)."""
        nccl = lambda t: torch.distributed.all_reduce(t, group=self.pg)   # noqa: E731
        self.syncbn_p2p = False
        if not int(os.environ.get("SSB_SYNCBN_P2P", "1")) or self.world < 2 or self.world > 16:
            return nccl
        lay = self.plan_s.lay
        slot = max(2 * b.C for b in lay.bns) * 2 + 2 * 64    # two adjacent BN slices (block output BN + shortcut BN)
        try:
            import torch.distributed._symmetric_memory as symm_mem
            nbytes = int(_lib.load().ssb_syncbn_mailbox_bytes(slot))
            box = symm_mem.empty(nbytes, dtype=torch.uint8, device=self.device)
            box.zero_()
            hdl = symm_mem.rendezvous(box, self.pg if self.pg is not None else torch.distributed.group.WORLD)
            ptrs = [int(p_) for p_ in hdl.buffer_ptrs]
            assert len(ptrs) == self.world and ptrs[hdl.rank] == box.data_ptr()
            torch.cuda.synchronize()
            torch.distributed.barrier(group=self.pg)
        except Exception as e:   # no peer access / allocator not available: keep the NCCL path
            import warnings
            warnings.warn(f"SyncBN peer-memory exchange unavailable ({type(e).__name__}: {e}); using NCCL all-reduce")
            return nccl
        self._mailbox, self._mail_hdl = box, hdl
        self._peers_dev = torch.tensor(ptrs, dtype=torch.int64, device=self.device)
        rank = hdl.rank
        self.syncbn_p2p = True

        def hook(t: torch.Tensor):
            assert t.dtype == torch.float64 and t.numel() <= slot
            call("ssb_syncbn_exchange", t.data_ptr(), t.numel(), self._peers_dev.data_ptr(), self.world, rank, slot,
                 torch.cuda.current_stream().cuda_stream)
        return hook

    # ---- data ----------------------------------------------------------------------
    def _load_batch_staged(self, ecg_x, mask_x, ecg_u_w, ecg_u_s) -> None:
        """Host batches: the H2D copies go to one of two device staging slots on a COPY stream -- they run under the
        previous step's graph, whose input arena they must not touch -- and the compute stream moves the slot into the
        arena with two device copies once it gets there.  (Copying straight into the arena serialises ~0.8 MB of PCIe
        traffic with the step: 0.762 vs 0.696 ms per step at 16+16 x 2500, profiles/r2_bench_c1.json.)"""
        if self._stage is None:
            self._stage = [(torch.empty_like(self.x_all), torch.empty_like(self.y_l)) for _ in range(2)]
            self._stage_free = [None, None]
            self._stage_i = 0
            self.copy_stream = torch.cuda.Stream(device=self.device)
        slot = self._stage_i % 2
        self._stage_i += 1
        sx, sy = self._stage[slot]
        cs, cur = self.copy_stream, torch.cuda.current_stream()
        if self._stage_free[slot] is not None:
            cs.wait_event(self._stage_free[slot])           # the arena copy that last read this slot
        S = self.Bl + self.Bu
        with torch.cuda.stream(cs):
            sx[: self.Bl].copy_(ecg_x, non_blocking=True)
            sy.copy_(mask_x, non_blocking=True)
            if self.hard_teacher:
                sx[S:].copy_(ecg_u_w, non_blocking=True)
            elif self.mode != _lib.LOSS_SUP:
                sx[S:].copy_(ecg_u_w, non_blocking=True)
                sx[self.Bl: S].copy_(ecg_u_s, non_blocking=True)
            ready = torch.cuda.Event()
            ready.record(cs)
        cur.wait_event(ready)
        if self.mode == _lib.LOSS_SUP:
            self.x_s[: self.Bl].copy_(sx[: self.Bl], non_blocking=True)
        else:
            self.x_all.copy_(sx, non_blocking=True)
            if self.hard_teacher:
                self.x_s[self.Bl:].copy_(self.x_uw, non_blocking=True)
        self.y_l.copy_(sy, non_blocking=True)
        free = torch.cuda.Event()
        free.record(cur)
        self._stage_free[slot] = free

    def load_batch(self, ecg_x: torch.Tensor, mask_x: torch.Tensor, ecg_u_w: Optional[torch.Tensor] = None,
                   ecg_u_s: Optional[torch.Tensor] = None) -> None:
        """Stage one batch into the static input arena (H2D when given host tensors)."""
        if not ecg_x.is_cuda and self.stage_h2d:
            return self._load_batch_staged(ecg_x, mask_x, ecg_u_w, ecg_u_s)
        self.x_s[: self.Bl].copy_(ecg_x, non_blocking=True)
        self.y_l.copy_(mask_x, non_blocking=True)
        if self.hard_teacher:
            # the student trains on the weak view itself (cps.py:113, stpp.py:158): one H2D, one device copy
            self.x_uw.copy_(ecg_u_w, non_blocking=True)
            self.x_s[self.Bl:].copy_(self.x_uw, non_blocking=True)
        elif self.mode != _lib.LOSS_SUP:
            self.x_uw.copy_(ecg_u_w, non_blocking=True)
            self.x_s[self.Bl:].copy_(ecg_u_s, non_blocking=True)

    def h2d_bytes(self) -> int:
        n = self.x_s.numel() * 4 + self.y_l.numel() * 8 + 64
        if self.hard_teacher:
            return self.x_s[: self.Bl].numel() * 4 + self.x_uw.numel() * 4 + self.y_l.numel() * 8 + 64
        if self.mode != _lib.LOSS_SUP:
            n += self.x_uw.numel() * 4
        return n

    # ---- one step ------------------------------------------------------------------
    def _write_step_params(self, lr: float) -> None:
        st = self.state
        t = st.step + 1
        slot = self.it % 8
        if self.sp_used[slot]:
            self.sp_events[slot].synchronize()
        sp = StepParams()
        sp.lr = lr
        sp.inv_bias1 = 1.0 / (1.0 - self.beta1 ** t)
        sp.inv_sqrt_bias2 = 1.0 / math.sqrt(1.0 - self.beta2 ** t)
        sp.ema_decay = float(self.cfg.get("ema_decay", 0.999))
        sp.ema_first = 1 if self.ema_first else 0
        sp.step = t
        sp.rng_seed = self.seed & 0xFFFFFFFF
        sp.rng_step = st.step & 0xFFFFFFFF      # optimizer step count: survives a resume (the engine's own counter does not)
        sp.grad_scale = 1.0 / self.world
        sp.conf_thresh = 0.0 if self.hard_teacher else float(self.cfg.get("conf_thresh", 0.0))
        C.memmove(self.sp_host[slot].data_ptr(), C.addressof(sp), 64)
        self.sp_dev.copy_(self.sp_host[slot], non_blocking=True)
        self.sp_events[slot].record()
        self.sp_used[slot] = True

    def _enqueue(self) -> None:
        """Enqueue the whole step on the current stream (captured into a graph once)."""
        st = _stream()
        w, state = self.w, self.state
        n0 = _lib.load().ssb_launch_count()
        call("ssb_memset_zero", state.zero_arena.data_ptr(), state.zero_arena.numel() * 4, st)   # grads, BN sums, loss sums
        cur = torch.cuda.current_stream()
        repacked = None
        if self.repack_stream is not None:
            # the storage-dtype weight copy overlaps the stem (which reads the fp32 master weights): own graph branch
            fork0 = torch.cuda.Event()
            fork0.record()
            self.repack_stream.wait_event(fork0)
            self.plan_s.sh.refresh(self.repack_stream.cuda_stream)
            if self.plan_t is not None and self.uses_teacher_weights and not self.external_pseudo:
                self.plan_t.sh.refresh(self.repack_stream.cuda_stream)
            repacked = torch.cuda.Event()
            repacked.record(self.repack_stream)
        else:
            self.plan_s.sh.refresh(st)
            if self.plan_t is not None and self.uses_teacher_weights and not self.external_pseudo:
                self.plan_t.sh.refresh(st)
        self.plan_s.pre_block_event = repacked
        low_t = None
        if self.merged:
            self.bufs_snap.copy_(w.bufs, non_blocking=True)
            _, low_t = self.plan_s.forward_merged(self.x_all, st)
        teacher_branch = False
        if self.plan_t is not None and self.external_pseudo:
            low_t = self.plan_t.low            # filled by pseudo(), which the owner ran before this step
        elif self.plan_t is not None and "teacher" in os.environ.get("SSB_DEBUG_SKIP", ""):
            low_t = self.plan_t.low            # timing experiments only: the step without its pseudo-label forward
        elif self.plan_t is not None:
            self.plan_t.pre_block_event = repacked
            teacher_branch = self.teacher_stream is not None
            if self.teacher_stream is not None:
                if self.bufs_snap is not None:
                    self.bufs_snap.copy_(w.bufs, non_blocking=True)
                fork = torch.cuda.Event()
                fork.record()
                self.teacher_stream.wait_event(fork)
                low_t = self.plan_t.forward(self.x_uw, self.teacher_stream.cuda_stream, train_mode=False,
                                            stream=self.teacher_stream)
            else:
                low_t = self.plan_t.forward(self.x_uw, st, train_mode=False)
        if not self.merged:
            self.plan_s.forward(self.x_s, st, train_mode=True, zero=False, stream=cur)
        if teacher_branch:
            torch.cuda.current_stream().wait_stream(self.teacher_stream)
        low_s = self.plan_s.low
        m = self.mat
        call("ssb_semi_loss", low_s.data_ptr(), self.y_l.data_ptr(), low_t.data_ptr() if low_t is not None else None,
             self.plan_s.dlow.data_ptr(), self.loss_sums.data_ptr(), self.Bl, self.Bu, self.plan_s.Lh, self.L,
             self.spec.num_classes, self.mode, 0.0, self.sp_dev.data_ptr(), 1 if self.spec.align_corners else 0,
             m["conf"].data_ptr() if m else None, m["label"].data_ptr() if m else None,
             m["mask"].data_ptr() if m else None, st)
        # ---- backward, gradient exchange and optimizer, pipelined by parameter range ----
        # The arena is in forward order, the backward produces gradients from its END: the ranges of the last stages
        # (+ head; ~97 % of the parameters) are complete when the backward reaches the first block of those stages.
        # Each such range is handed to the update stream at once -- [all-reduce over the ranks,] fused AdamW(+EMA) on the
        # range -- under the backward of the earlier stages; only the early layers' small range is exchanged and updated
        # after the backward (a 17 us optimizer pass over the whole arena used to sit at the end of the critical path).
        pe_base = self.teacher.params.data_ptr() if self.algorithm == "mean_teacher" else None

        # max_norm (loss_scaler(..., clip_grad=max_norm), fixmatch.py:129-136 -> clip_grad_norm_, misc.py:248-250): the
        # global norm of the exchanged gradients first, then every range's update scales by min(1, max_norm/(norm+1e-6))
        max_norm = self.cfg.get("max_norm", None)

        def adamw(lo, hi, stream_ptr):
            args = (w.params.data_ptr() + 4 * lo, state.grads.data_ptr() + 4 * lo, state.exp_avg.data_ptr() + 4 * lo,
                    state.exp_avg_sq.data_ptr() + 4 * lo, (pe_base + 4 * lo) if pe_base else None, hi - lo, self.beta1,
                    self.beta2, self.eps, self.wd, self.sp_dev.data_ptr())
            if max_norm is not None:
                call("ssb_adamw_ema_clip", *args, self.gnorm.data_ptr(), float(max_norm), stream_ptr)
            else:
                call("ssb_adamw_ema", *args, stream_ptr)

        split = None
        want_norm = bool(self.cfg.get("grad_norm", False)) or max_norm is not None
        if self.comm_stream is not None:
            lay = self.plan_s.lay
            nst = len(self.spec.stage_blocks)
            firsts = {}
            for st_i in range(nst - 1, max(nst - 1 - int(os.environ.get("SSB_BUCKETS", "2")), 0) - 1, -1):
                firsts[next(i for i, b in enumerate(lay.blocks) if b.stage == st_i)] = st_i
            bounds = {bi: lay.blocks[bi].conv1.poff for bi in firsts}
            upper = {"v": state.grads.numel()}
            split = min(bounds.values())

            def bucket(bi):
                if bi not in bounds:
                    return
                lo, hi = bounds[bi], upper["v"]
                upper["v"] = lo
                cur_ = torch.cuda.current_stream()
                self.comm_stream.wait_stream(cur_)                       # BN / head gradients of the range (main stream)
                if self.wgrad_stream is not None:
                    self.comm_stream.wait_stream(self.wgrad_stream)      # conv weight gradients of the range
                # (the shortcut dgrads on their own branch are joined into the main stream inside their block)
                with torch.cuda.stream(self.comm_stream):
                    if self.collectives:
                        torch.distributed.all_reduce(state.grads[lo:hi], group=self.pg)
                    if not want_norm:      # (the global norm reads every gradient first; the update then waits for it)
                        adamw(lo, hi, self.comm_stream.cuda_stream)
            self.plan_s.block_done_hook = bucket
        self.plan_s.backward(self.plan_s.dlow, st)
        self.plan_s.block_done_hook = None
        if self.collectives:
            if split is not None:
                torch.distributed.all_reduce(state.grads[:split], group=self.pg)
            else:
                torch.distributed.all_reduce(state.grads, group=self.pg)
        if split is not None and want_norm:
            torch.cuda.current_stream().wait_stream(self.comm_stream)
        if want_norm:
            call("ssb_memset_zero", self.gnorm_ws.data_ptr(), 8, st)
            call("ssb_grad_norm", state.grads.data_ptr(), state.grads.numel(), self.gnorm_ws.data_ptr(),
                 self.gnorm.data_ptr(), st)
        if split is not None and not want_norm:
            adamw(0, split, st)
            torch.cuda.current_stream().wait_stream(self.comm_stream)
        else:
            adamw(0, w.params.numel(), st)
        if self.algorithm == "mean_teacher":
            call("ssb_ema", self.teacher.bufs.data_ptr(), w.bufs.data_ptr(), w.bufs.numel(), self.sp_dev.data_ptr(), st)
            call("ssb_ema_i64", self.teacher.nbt.data_ptr(), w.nbt.data_ptr(), w.nbt.numel(), self.sp_dev.data_ptr(), st)
        self.launches_per_step = int(_lib.load().ssb_launch_count() - n0)

    def pseudo(self) -> None:
        """external_pseudo engines: storage-dtype copy of the teacher's weights + its eval-mode forward on the staged weak
        views, into the logits buffer the next step() reads.  Its own small graph, replayed on the current stream."""
        assert self.external_pseudo and self.plan_t is not None
        if self.use_graph:
            if self.pseudo_graph is None:
                g = torch.cuda.CUDAGraph()
                with torch.cuda.graph(g, capture_error_mode="thread_local"):
                    self._enqueue_pseudo()
                self.pseudo_graph = g
            self.pseudo_graph.replay()
        else:
            self._enqueue_pseudo()

    def _enqueue_pseudo(self) -> None:
        st = _stream()
        self.plan_t.sh.refresh(st)
        self.plan_t.pre_block_event = None
        self.plan_t.forward(self.x_uw, st, train_mode=False)

    def step(self, lr: float) -> None:
        """Run one optimizer step on the staged batch (asynchronous)."""
        if self.it == 0 and getattr(self, "syncbn_p2p", False):
            # the peer-memory exchange waits for its peers with a bounded spin: let the ranks enter their first step together
            torch.cuda.synchronize()
            torch.distributed.barrier(group=self.pg)
        self._write_step_params(lr)
        if self.use_graph:
            if self.graph is None:
                # warm-up run outside capture is NOT done (it would apply an update);
                # all kernels are capture-safe (no sync, no allocation).
                g = torch.cuda.CUDAGraph()
                # the main chain is captured on a HIGH-priority stream (kernel nodes inherit it): when its kernels
                # compete for SMs with the wgrad / pseudo-label / shortcut branches, the critical path goes first
                prio = int(os.environ.get("SSB_MAIN_PRIORITY", "-1"))
                cap = torch.cuda.Stream(device=self.device, priority=prio) if prio else None
                with torch.cuda.graph(g, stream=cap, capture_error_mode="thread_local"):
                    self._enqueue()
                self.graph = g
            self.graph.replay()
        else:
            self._enqueue()
        self.state.step += 1
        self.ema_first = False
        slot = self.it % 8
        if slot in self._pending:  # ring wrapped: retire the old entry first
            self._retire(slot)
        self.stats_host[slot].copy_(self.loss_sums, non_blocking=True)
        self.stats_events[slot].record()
        self._pending.append(slot)
        self.it += 1

    def _to_stats(self, s: torch.Tensor) -> Dict[str, float]:
        nx = float(self.Bl * self.L)
        loss_x = float(s[0]) / nx
        if self.mode == _lib.LOSS_SUP:
            return {"loss": loss_x}
        nu = float(self.Bu * self.L)
        loss_u = float(s[1]) / nu
        out = {"loss_total": (loss_x + loss_u) / 2.0, "loss_x": loss_x, "loss_u_s": loss_u}
        if self.mode == _lib.LOSS_FIXMATCH and not self.hard_teacher:
            out["mask_ratio"] = float(s[2]) / nu
        return out

    def _retire(self, slot: int) -> None:
        self.stats_events[slot].synchronize()
        self._done.append(self._to_stats(self.stats_host[slot].clone()))
        self._pending.remove(slot)

    def read_stats(self, block: bool = True) -> List[Dict[str, float]]:
        """Stats of all steps issued since the last call, in order (synchronises on the
        outstanding asynchronous D2H copies only).  block=False: only the steps whose read-back has already
        arrived (no host wait, no bubble in the GPU's queue) -- the rest come with a later call."""
        for slot in list(self._pending):
            if not block and not self.stats_events[slot].query():
                break
            self._retire(slot)
        if getattr(self, "syncbn_p2p", False):
            # the peer-memory statistics exchange gives up on a silent peer after ~20 s and records the exchange
            # number in its mailbox header instead of hanging the GPU: statistics summed after that are wrong
            err = int(self._mailbox[64:68].view(torch.int32).item())
            if err:
                raise RuntimeError(f"SyncBN statistics exchange timed out waiting for a peer (exchange #{err}); "
                                   "the BatchNorm statistics of this rank are no longer trustworthy")
        out, self._done = self._done, []
        return out


class CpsEngine:
    """Cross Pseudo Supervision step (reference src/algorithms/cps.py:96-160): two models, each trained on
    cat(ecg_x, ecg_u_w) against the labels and the OTHER model's hard pseudo-labels; both pseudo-label passes read the
    weights and running statistics from before either update.  Two hard-teacher StepEngines over the two models'
    arenas: pseudo(1<-2), pseudo(2<-1), then the two training graphs -- they touch disjoint state, so the second one
    runs on its own stream beside the first."""

    def __init__(self, eng_1: StepEngine, eng_2: StepEngine):
        assert eng_1.algorithm == "cps" and eng_2.algorithm == "cps" and eng_1.external_pseudo and eng_2.external_pseudo
        assert eng_1.teacher is eng_2.w and eng_2.teacher is eng_1.w, "each engine's teacher must be the other model"
        self.engines = (eng_1, eng_2)
        # (with collectives on, both engines issue NCCL all-reduces on one communicator: keep them in one stream order)
        self.side = torch.cuda.Stream(device=eng_1.device) if (int(os.environ.get("SSB_CPS_CONCURRENT", "1")) and
                                                                not eng_1.collectives) else None

    def load_batch(self, ecg_x, mask_x, ecg_u_w) -> None:
        a, b = self.engines
        a.load_batch(ecg_x, mask_x, ecg_u_w)
        # the second engine's arena is filled from the first one's (device copies): one H2D per step, not two
        b.x_all.copy_(a.x_all, non_blocking=True)
        b.y_l.copy_(a.y_l, non_blocking=True)

    def h2d_bytes(self) -> int:
        return self.engines[0].h2d_bytes()

    def step(self, lr_1: float, lr_2: Optional[float] = None) -> None:
        a, b = self.engines
        lr_2 = lr_1 if lr_2 is None else lr_2
        if self.side is None:
            a.pseudo()
            b.pseudo()
            a.step(lr_1)
            b.step(lr_2)
            return
        cur = torch.cuda.current_stream()
        self.side.wait_stream(cur)                # the staged batch
        a.pseudo()
        with torch.cuda.stream(self.side):
            b.pseudo()
        # each training step ends by updating the weights the OTHER engine's pseudo-label pass reads
        cur.wait_stream(self.side)
        self.side.wait_stream(cur)
        a.step(lr_1)
        with torch.cuda.stream(self.side):
            b.step(lr_2)
        cur.wait_stream(self.side)

    def read_stats(self) -> List[Dict[str, float]]:
        """Per step, the mean of the two models' losses (cps.py:153-160)."""
        sa, sb = self.engines[0].read_stats(), self.engines[1].read_stats()
        return [{k: (x[k] + y[k]) / 2.0 for k in x} for x, y in zip(sa, sb)]
