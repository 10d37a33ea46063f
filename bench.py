#!/usr/bin/env python
"""Benchmark of the SemiSegECG training hot path on B200 (contract: see the task statement).

metric  : train samples/sec of the FixMatch step (labeled + unlabeled strips per second, whole job)
workload: BASELINE.json configs[1] -- FixMatch, resnet18 + FCNHead, LUDB 1/16 shape
          (configs/base/resnet18/fixmatch.yaml + configs/bench/ludb/1over16.yaml): 1 lead x 2500
          samples, B_l = B_u = 16 per GPU, conf_thresh 0.8, AdamW; synthetic LUDB-shaped data,
          random-init weights.
value   : throughput with inputs resident in HBM; e2e: through algorithms-level engine API with
          pinned-host batches copied H2D and the loss read back D2H inside the timed region.
--impl reference: the CPU implementation of the same step (oracle port of the reference's
          fixmatch.train_one_epoch body; the reference is Python and is not present on the GPU box).
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

REPO = os.path.dirname(os.path.abspath(__file__))
SRC = os.path.join(REPO, "semi-seg-ecg_b200", "src")
for p in (REPO, SRC):
    if p not in sys.path:
        sys.path.insert(0, p)

import numpy as np  # noqa: E402
import torch  # noqa: E402

WORKLOADS = {
    # name: (algorithm, num_leads, L, Bl, Bu, base_channels, stem_channels, head_in)
    "fixmatch_resnet18_ludb_1x2500_b16+16": ("fixmatch", 1, 2500, 16, 16, 64, 64),
    "mean_teacher_resnet18_qtdb_2x2500_b16+16": ("mean_teacher", 2, 2500, 16, 16, 64, 64),
    "fixmatch_resnet18w128_12x5000_b32+32": ("fixmatch", 12, 5000, 32, 32, 128, 128),
    # BASELINE configs[3]: cross-domain merged index, labeled:unlabeled 1:7 -> B_l = 4, B_u = 28 per GPU (1 x 2500)
    "fixmatch_resnet18_merged_1x2500_b4+28": ("fixmatch", 1, 2500, 4, 28, 64, 64),
    # BASELINE configs[4]: widened network, 12 x 5000, global batch 256 ... 4096 on 8 GPUs = 16+16 ... 256+256 per GPU
    "fixmatch_resnet18w128_12x5000_b16+16": ("fixmatch", 12, 5000, 16, 16, 128, 128),
    "fixmatch_resnet18w128_12x5000_b64+64": ("fixmatch", 12, 5000, 64, 64, 128, 128),
    "fixmatch_resnet18w128_12x5000_b128+128": ("fixmatch", 12, 5000, 128, 128, 128, 128),
    "fixmatch_resnet18w128_12x5000_b256+256": ("fixmatch", 12, 5000, 256, 256, 128, 128),
}
DEFAULT_WORKLOAD = "fixmatch_resnet18_ludb_1x2500_b16+16"


def preroll_steps(workload, preroll_s):
    """Untimed steps before the warm-up at N > 1: a function of the CONFIG alone, so every rank runs the same number
    (each step holds collectives).  ~1 ms per step at the default workload, scaled by the workload's size."""
    _, _, L, Bl, Bu, base, _ = WORKLOADS[workload]
    rel_work = (Bl + Bu) * L * (base / 64.0) ** 2 / (32 * 2500.0)
    return int(min(max(preroll_s * 1000.0 / max(rel_work, 1.0), 20), 1000))


def make_host_batch(seed, rank, Bl, Bu, C, L):
    from semiseg_b200 import synthetic
    return synthetic.make_batch(seed * 1000 + rank, Bl, Bu, C, L)


def load_cfg(workload):
    from utils.config import load_config
    algo, C, L, Bl, Bu, base, stem = WORKLOADS[workload]
    cfgdir = os.path.join(REPO, "semi-seg-ecg_b200", "configs")
    cfg = load_config(os.path.join(cfgdir, "base", "resnet18", f"{algo}.yaml"),
                      os.path.join(cfgdir, "bench", "ludb", "1over16.yaml"))
    bb = cfg["backbone"]["resnet18"]
    bb["num_leads"] = C
    if base != 64 or stem != 64:
        bb["base_channels"], bb["stem_channels"] = base, stem
        cfg["decode_head"]["FCNHead"]["in_channels"] = base * 8
    cfg["dataset"]["signal_length"] = L
    cfg["dataloader"]["batch_size"] = Bl
    return cfg, algo, C, L, Bl, Bu


def config_block(workload, world, sync_bn):
    """The `config` object of the JSON line -- identical for the B200 arm and the reference arm of the same launch."""
    algo, C, L, Bl, Bu, base, stem = WORKLOADS[workload]
    return {"workload": workload, "algorithm": algo, "per_gpu_batch": f"{Bl}+{Bu}", "leads": C, "length": L,
            "base_channels": base, "parallelism": f"dp{world}", "sync_bn": bool(sync_bn and world > 1),
            "l2_policy": "no explicit flush: the per-step working set (activations + parameter / gradient / Adam arenas; "
                         "164 MiB at 16+16 x 1 x 2500, reported as working_set_mib) exceeds the 126 MB L2 and the inputs "
                         "rotate over 4 distinct batches; 0.4 s of untimed steps (N > 1: the same fixed count on every rank) precede the W warm-up steps (clocks out of idle)"}


class ClockSampler(threading.Thread):
    def __init__(self, gpu_index):
        super().__init__(daemon=True)
        self.gpu = gpu_index
        self.samples, self.reasons, self.stop_flag = [], set(), False
        self.max_mhz = None

    def run(self):
        q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
             "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        while not self.stop_flag:
            try:
                out = subprocess.run(["nvidia-smi", f"--query-gpu={q}", "--format=csv,noheader,nounits", "-i", str(self.gpu)],
                                     capture_output=True, text=True, timeout=5).stdout.strip().split(",")
                self.samples.append(float(out[0]))
                self.max_mhz = float(out[1])
                for n, v in zip(names, out[2:]):
                    if v.strip().lower().startswith("active"):
                        self.reasons.add(n)
            except Exception:
                pass
            time.sleep(0.1)

    def summary(self):
        return {"sm_mhz": float(np.median(self.samples)) if self.samples else None, "sm_max_mhz": self.max_mhz,
                "reasons": sorted(self.reasons), "samples": len(self.samples)}


# ---- algorithmic cost model (per launch) ------------------------------------------------------
def launch_cost(name, args, es):
    """(flops, bytes) of one C-ABI launch from its arguments; es = activation element size."""
    from semiseg_b200._lib import Geom
    geoms = [a for a in args if isinstance(a, Geom)]
    if name in ("ssb_conv1d_fwd", "ssb_conv1d_fwd_stats", "ssb_conv1d_bn_act_fwd", "ssb_conv1d_dgrad", "ssb_conv1d_wgrad",
                "ssb_conv1d_dgrad_bnred", "ssb_conv1d_fwd_dual"):
        gi, go = geoms
        k = args[6] if name == "ssb_conv1d_fwd_dual" else args[5]
        flops = 2.0 * go.B * go.len * go.C * gi.C * k
        w = gi.C * go.C * k
        if name == "ssb_conv1d_wgrad":
            byt = (gi.B * gi.len * gi.C + go.B * go.len * go.C) * es + w * 4
        else:
            byt = (gi.B * gi.len * gi.C + go.B * go.len * go.C) * es + w * es
        return flops, byt, f"{name[4:]}[{gi.C}->{go.C},k{k},L{gi.len}->{go.len}]"
    if geoms:
        g = geoms[-1]
        n = g.B * g.len * g.C
        mult = {"ssb_bn_stats": 1, "ssb_bn_act_fwd": 2.5, "ssb_bn_bwd_reduce": 3, "ssb_bn_bwd_apply": 4, "ssb_bn_bwd_fused": 4,
                "ssb_stem_bn_relu_pool_fwd": 3, "ssb_stem_bwd_reduce": 3, "ssb_stem_bwd_apply": 3.5}.get(name, 2)
        return 0.0, n * es * mult, f"{name[4:]}[C{g.C},L{g.len}]"
    if name == "ssb_adamw_ema":
        n = args[5]
        return 0.0, n * (36.0 if args[4] else 28.0), "adamw_ema"
    return 0.0, 0.0, name[4:]


FAMILIES = [
    ("conv_tn (tcgen05 implicit-GEMM fprop/dgrad + fused BN epilogues)", ("conv1d_fwd", "conv1d_dgrad", "conv1d_bn_act_fwd")),
    ("conv_wgrad (tcgen05, MN-major operands, split over rows)", ("conv1d_wgrad",)),
    ("bn/relu/residual elementwise passes", ("bn_", "stem_bn", "stem_bwd")),
    ("stem direct conv (fwd + wgrad)", ("stem_conv",)),
    ("head / loss / optimizer", ("head_", "semi_loss", "adamw", "ema", "weight_shadow", "memset", "grad_norm", "upsample", "pseudo")),
]


def family_of(label):
    for fam, prefixes in FAMILIES:
        if any(label.startswith(p) for p in prefixes):
            return fam
    return "other"


def steady_state_profile(eng_e, batch, lr, es, repeats=20):
    """Per distinct C-ABI launch of one step: device time of the launch replayed back-to-back inside a
    captured CUDA graph (CUDA events on the capture stream around `repeats` x 5 launches) -> steady-state
    duration with launch gaps, warm caches.  Returns [(label, n_per_step, us, flops, bytes)]."""
    from semiseg_b200 import _lib
    rec = []

    def hook(name, a):
        rec.append((name, a))
        _lib.raw_call(name, *a)
    _lib._hook = hook
    eng_e.load_batch(*batch)
    eng_e.step(lr)
    _lib._hook = None
    torch.cuda.synchronize()
    eng_e.read_stats()
    uniq = {}
    for name, a in rec:
        fl, by, label = launch_cost(name, a, es)
        d = uniq.setdefault(label, {"name": name, "args": a, "n": 0, "flops": fl, "bytes": by})
        d["n"] += 1
    out = []
    s = torch.cuda.Stream()
    with torch.cuda.stream(s):
        for label, d in uniq.items():
            args = list(d["args"])
            args[-1] = s.cuda_stream
            g = torch.cuda.CUDAGraph()
            with torch.cuda.graph(g, stream=s, capture_error_mode="thread_local"):
                for _ in range(repeats):
                    _lib.raw_call(d["name"], *args)
            for _ in range(2):
                g.replay()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(s)
            for _ in range(5):
                g.replay()
            e1.record(s)
            e1.synchronize()
            out.append((label, d["n"], e0.elapsed_time(e1) * 1e3 / (5 * repeats), d["flops"], d["bytes"]))
    torch.cuda.synchronize()
    return out


def roofline_from_profile(prof, peaks, traffic=None):
    """Aggregate the steady-state launch times by kernel family; the dominant family (largest share of the
    summed device time) gets the roofline entry: achieved = summed algorithmic FLOPs (or bytes) / summed time.
    These are kernels timed in isolation -> the BURST bf16 peak is the denominator (frac_of_sustained beside it)."""
    fam = {}
    tot = 0.0
    for label, n, us, fl, by in prof:
        f = fam.setdefault(family_of(label), {"us": 0.0, "flops": 0.0, "bytes": 0.0, "launches": 0})
        f["us"] += us * n
        f["flops"] += fl * n
        f["bytes"] += by * n
        f["launches"] += n
        tot += us * n
    ridge = peaks["bf16_tflops"] * 1e12 / (peaks["hbm_gbs"] * 1e9)
    ranked = sorted(fam.items(), key=lambda kv: -kv[1]["us"])
    name, f = ranked[0]
    if f["flops"] and f["bytes"] and f["flops"] / f["bytes"] > ridge:
        ach = f["flops"] / (f["us"] * 1e-6) / 1e12
        roof = {"bound": "tensor", "achieved": round(ach, 2), "peak": peaks["bf16_tflops"], "unit": "TFLOP/s",
                "frac": round(ach / peaks["bf16_tflops"], 4), "traffic": None,
                "frac_of_sustained": round(ach / peaks["bf16_tflops_sustained"], 4)}
    else:
        ach = f["bytes"] / (f["us"] * 1e-6) / 1e9
        roof = {"bound": "hbm", "achieved": round(ach, 1), "peak": peaks["hbm_gbs"], "unit": "GB/s",
                "frac": round(ach / peaks["hbm_gbs"], 4), "traffic": None}
    if traffic and name in traffic.get("families", {}):
        tf = traffic["families"][name]
        roof["traffic"] = round(tf["dram_bytes_per_step"] / max(tf["launches_per_step"], 1))
        roof["traffic_detail"] = {"dram_mbyte_per_step": round(tf["dram_bytes_per_step"] / 1e6, 2),
                                  "algorithmic_mbyte_per_step": round(f["bytes"] / 1e6, 2),
                                  "dram_over_algorithmic": round(tf["dram_bytes_per_step"] / max(f["bytes"], 1), 3),
                                  "launches_in_capture": tf["launches_per_step"], "source": traffic.get("source")}
    roof.update({"kernel": name, "launches_per_step": f["launches"], "us_per_launch": round(f["us"] / f["launches"], 2),
                 "share_of_serial_sum": round(f["us"] / tot, 4), "peak_source": peaks["src"],
                 "peak_kind": "burst (kernels timed in isolation)",
                 "algorithmic_per_step": {"gflop": round(f["flops"] / 1e9, 2), "mbyte": round(f["bytes"] / 1e6, 2)},
                 "timing": "CUDA events on the launch stream around each distinct launch replayed 20x5 times inside a "
                           "captured graph (steady state, warm L2), summed over the family's launches of one step"})
    fams = [{"family": k, "share": round(v["us"] / tot, 4), "launches_per_step": v["launches"], "us_per_step": round(v["us"], 1),
             "tflops": round(v["flops"] / (v["us"] * 1e-6) / 1e12, 2) if v["flops"] else None,
             "gbs": round(v["bytes"] / (v["us"] * 1e-6) / 1e9, 1) if v["bytes"] else None} for k, v in ranked]
    return roof, fams, tot


def load_peaks():
    """bf16_tflops = the BURST figure (denominator for kernels timed in isolation); bf16_tflops_sustained for
    anything timed inside the long step."""
    peaks = {"hbm_gbs": 6650.0, "bf16_tflops": 1590.0, "bf16_tflops_sustained": 1400.0, "src": "fallback (B200_PROFILING.md)"}
    pf = os.path.join(REPO, "MEASURED_PEAKS.json")
    if os.path.exists(pf):
        mp = json.load(open(pf))
        peaks = {"hbm_gbs": mp["hbm_gbs"], "bf16_tflops": mp["bf16_tflops"],
                 "bf16_tflops_sustained": mp.get("bf16_tflops_sustained", mp["bf16_tflops"]), "src": "measured (MEASURED_PEAKS.json)"}
    return peaks


def load_traffic(workload):
    """DRAM bytes per kernel family per step from the committed ncu pass (profiles/r2_traffic.json, written by
    tools/profile_summary.py traffic from `ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,...` of
    `bench.py --profile-mode`); None if there is no capture of this workload."""
    pf = os.path.join(REPO, "profiles", "r2_traffic.json")
    if not os.path.exists(pf):
        return None
    t = json.load(open(pf))
    return t.get(workload)


def _finish(world, engines=()):
    """Multi-rank teardown.  Round 1 left with os._exit(0) because destroying the NCCL communicator while CUDA graphs that
    captured collectives on it were alive could hang.  Now: synchronise, barrier, RESET the captured graphs (they hold the
    communicator's work), then destroy the process group -- with a 15 s watchdog that falls back to os._exit(0), so a
    teardown problem can never turn into a hung benchmark."""
    if world > 1:
        import torch.distributed as dist
        torch.cuda.synchronize()
        dist.barrier()
        torch.cuda.synchronize()
        sys.stdout.flush()
        sys.stderr.flush()
        guard = threading.Timer(15.0, lambda: os._exit(0))
        guard.daemon = True
        guard.start()
        try:
            for e in engines:
                for name in ("graph", "pseudo_graph"):
                    g = getattr(e, name, None)
                    if g is not None:
                        g.reset()
                        setattr(e, name, None)
            torch.cuda.synchronize()
            dist.destroy_process_group()
        finally:
            guard.cancel()


def run_b200(args):
    import torch.distributed as dist
    from algorithms.base import init_model_from_cfg
    from algorithms.mean_teacher import init_teacher
    from semiseg_b200 import _lib
    from semiseg_b200.trainer import get_engine
    from utils.lr_sched import lr_at

    rank = int(os.environ.get("RANK", 0))
    world = int(os.environ.get("WORLD_SIZE", 1))
    local = int(os.environ.get("LOCAL_RANK", 0))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    assert world == args.gpus or world == 1 and args.gpus == 1, f"--gpus {args.gpus} but WORLD_SIZE={world}"

    cfg, algo, C, L, Bl, Bu = load_cfg(args.workload)
    tcfg = cfg["train"]
    dtype = _lib.BF16 if args.dtype == "bf16" else _lib.F32
    torch.manual_seed(cfg["seed"])           # identical initial weights on every rank
    model = init_model_from_cfg(cfg).to(dev)
    # --sync-bn: default = the YAML's ddp.sync_bn (true in every shipped config, fixmatch.yaml:131 of the reference)
    want_sync_bn = bool(cfg["ddp"].get("sync_bn", True)) if args.sync_bn is None else bool(args.sync_bn)
    model.sync_bn = want_sync_bn and world > 1
    model.seed = cfg["seed"] + rank
    teacher = init_teacher(cfg, model, dev) if algo == "mean_teacher" else None
    eng = get_engine(algo, model, teacher, Bl, Bu, L, dtype, tcfg, use_graph=not args.no_graph)

    # synthetic batches: a pool of distinct host (pinned) and device copies, shard = seed + rank
    pool = 4
    host, devb = [], []
    for i in range(pool):
        lab, unl = make_host_batch(cfg["seed"] + 17 * i, rank, Bl, Bu, C, L)
        h = [torch.from_numpy(lab["ecg"]).pin_memory(), torch.from_numpy(lab["target"]).pin_memory(),
             torch.from_numpy(unl["ecg"]).pin_memory(), torch.from_numpy(unl["ecg_aug"]).pin_memory()]
        host.append(h)
        devb.append([t.to(dev) for t in h])

    def sync_all():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    epoch_f = [20.0]  # a post-warm-up point of the LR schedule (lr ~ 9.8e-4)

    def step_from(batch):
        eng.load_batch(*batch)
        eng.step(lr_at(epoch_f[0], tcfg))
        epoch_f[0] += 1e-3

    def timed(batches, steps, read_each_step):
        ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        per = [torch.cuda.Event(enable_timing=True) for _ in range(steps)] if os.environ.get("SSB_BENCH_PERSTEP") else None
        sync_all()
        ev0.record()
        for i in range(steps):
            step_from(batches[i % pool])
            if per is not None:
                per[i].record()
        ev1.record()
        sync_all()
        ms = ev0.elapsed_time(ev1)
        if per is not None and rank == 0:      # development: where a short timed region spends its time
            ts = [ev0.elapsed_time(e_) for e_ in per]
            print("per-step end times (ms):", " ".join(f"{t:.3f}" for t in ts[:24]), "| deltas:",
                  " ".join(f"{b - a:.3f}" for a, b in zip([0.0] + ts[:23], ts[:24])), file=sys.stderr)
        if world > 1:
            t = torch.tensor([ms], device=dev)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            ms = float(t)
        stats = eng.read_stats()
        return ms, stats

    # pre-roll: bring the GPU out of its idle clocks before the W warm-up steps (a 20-step timed region is 14 ms long; measured
    # 0.726 ms per step over 20 steps right after start-up against 0.694 over 200) -- untimed, stated in config.l2_policy
    # Every rank must run the SAME number of steps (each step holds collectives: the gradient all-reduce and the SyncBN
    # exchange, whose mailbox words are tagged with the step count).  N = 1: steps until preroll_s of wall time have passed;
    # N > 1: a FIXED count (about the same time) -- a per-rank wall-clock loop let the ranks stop at different counts, after
    # which their collectives were off by one step: the N = 4 run of call c18 hung in exactly that way.
    if not args.profile_mode:
        t_pre = time.time()
        n_pre = 0
        n_fixed = preroll_steps(args.workload, args.preroll_s) if world > 1 else None
        while (n_pre < n_fixed) if n_fixed is not None else (time.time() - t_pre < args.preroll_s):
            step_from(devb[n_pre % pool])
            n_pre += 1
            if n_pre % 50 == 0:
                torch.cuda.synchronize()
        torch.cuda.synchronize()
        eng.read_stats()
    for i in range(max(args.warmup, 3)):
        step_from(devb[i % pool])
    eng.read_stats()
    if args.profile_mode:      # plain launch sequence for ncu: K more steps, nothing else
        torch.cuda.synchronize()
        for i in range(args.steps):
            step_from(devb[i % pool])
        torch.cuda.synchronize()
        if rank == 0:
            print(json.dumps({"profile_mode": True, "steps": args.steps, "launches_per_step": eng.launches_per_step}), flush=True)
        _finish(world, [eng])
        return
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
        time.sleep(0.3)
    ms_dev, stats_dev = timed(devb, args.steps, False)

    # ---- e2e: the call a user of the reference makes -- algorithms.<algo>.train_one_epoch(model, labeled_loader,
    # unlabeled_loader, optimizer, device, epoch, loss_scaler, log_writer, use_amp, config['train']) -- over loaders that
    # yield PINNED HOST batch dicts ({'ecg','target'} / {'ecg','ecg_aug'}, semi_dataset.py:235-244): every step's H2D
    # copies, the per-step LR / param-group writes, the engine lookup and the read-back of the loss sums are inside.
    import contextlib
    import algorithms
    from utils.misc import NativeScalerWithGradNormCount
    from utils.optimizer import get_optimizer_from_config
    optimizer = get_optimizer_from_config(tcfg, model.parameters())
    scaler = NativeScalerWithGradNormCount()
    toe = getattr(algorithms, algo).train_one_epoch

    def epoch_call(n):
        labs = [{"ecg": host[i % pool][0], "target": host[i % pool][1]} for i in range(n)]
        unls = [{"ecg": host[i % pool][2], "ecg_aug": host[i % pool][3]} for i in range(n)]
        with contextlib.redirect_stdout(sys.stderr):      # the epoch loop logs like the reference; stdout carries ONE line
            if algo == "mean_teacher":
                return toe(model, teacher, labs, unls, optimizer, dev, 20, scaler, None, args.dtype == "bf16", tcfg)
            return toe(model, labs, unls, optimizer, dev, 20, scaler, None, args.dtype == "bf16", tcfg)
    epoch_call(3)
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    sync_all()
    ev0.record()
    stats_epoch = epoch_call(args.steps)
    ev1.record()
    sync_all()
    ms_e2e = ev0.elapsed_time(ev1)
    if world > 1:
        t = torch.tensor([ms_e2e], device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms_e2e = float(t)
    stats_e2e = [stats_epoch]
    assert get_engine(algo, model, teacher, Bl, Bu, L, dtype, tcfg, use_graph=not args.no_graph) is eng or args.no_graph, \
        "train_one_epoch must run on the engine that was timed"
    sampler.stop_flag = True
    per_step = Bl + Bu
    value = per_step * world * args.steps / (ms_dev / 1e3)
    e2e = per_step * world * args.steps / (ms_e2e / 1e3)
    assert all(np.isfinite(s["loss_total"]) for s in stats_dev + stats_e2e), "non-finite loss in the timed region"

    # ---- N>1: every rank must hold the same weights after the timed loops (replicated optimizer, averaged gradients) ----
    replicas_equal = None
    if world > 1:
        w_ = model.runtime().weights.params
        chk = torch.stack([w_.double().sum(), (w_.double() * w_.double()).sum()])
        lo, hi = chk.clone(), chk.clone()
        dist.all_reduce(lo, op=dist.ReduceOp.MIN)
        dist.all_reduce(hi, op=dist.ReduceOp.MAX)
        rel = float(((hi - lo).abs() / (hi.abs() + 1e-30)).max())
        # SyncBN off: ranks see different BN statistics only through their own shard's forward; the averaged gradient and
        # therefore the parameters are identical on every rank either way
        replicas_equal = {"max_rel_spread_of_param_checksums": rel, "ok": bool(rel < 1e-9)}
        assert replicas_equal["ok"], f"ranks hold different weights after the timed loop: {replicas_equal}"

    # ---- SyncBN on (the reference YAML's default, ddp.sync_bn: true): same timed loop, reported next to the headline ----
    syncbn_other = None
    if world > 1 and not args.no_syncbn_other:
        main_sb = bool(model.sync_bn)
        model.sync_bn = not main_sb
        eng_sb = get_engine(algo, model, teacher, Bl, Bu, L, dtype, tcfg, use_graph=not args.no_graph)
        model.sync_bn = main_sb
        eng_main, eng = eng, eng_sb
        for i in range(max(args.warmup, 3)):
            step_from(devb[i % pool])
        eng.read_stats()
        ms_sb, stats_sb = timed(devb, args.steps, False)
        eng = eng_main
        syncbn_other = {"sync_bn": not main_sb, "value": round(per_step * world * args.steps / (ms_sb / 1e3), 1), "unit": "samples/s",
                        "ms_per_step": round(ms_sb / args.steps, 4),
                        "exchange": ("peer-memory kernel" if getattr(eng_sb, "syncbn_p2p", False) else "nccl") if not main_sb else None,
                        "finite": bool(all(np.isfinite(s_["loss_total"]) for s_ in stats_sb))}

    # ---- per-kernel timing pass: every distinct launch of the step, steady state, on its launch stream ----
    roof, top, fams, large = None, [], [], None
    if rank == 0 or world > 1:   # (every rank takes part when the step contains collectives)
        es = 2 if dtype == _lib.BF16 else 4
        peaks = load_peaks()
        main_sb = bool(model.sync_bn)
        model.sync_bn = False    # (the peer-memory statistics exchange must not be replayed out of lockstep; same kernels otherwise)
        eng_e = get_engine(algo, model, teacher, Bl, Bu, L, dtype, tcfg, use_graph=False)
        model.sync_bn = main_sb
        eng_e.load_batch(*devb[0]); eng_e.step(lr_at(epoch_f[0], tcfg))
        torch.cuda.synchronize()
        prof = steady_state_profile(eng_e, devb[0], lr_at(epoch_f[0], tcfg), es)
        roof, fams, tot = roofline_from_profile(prof, peaks, load_traffic(args.workload))
        # step level: all algorithmic conv FLOPs of one step over the measured step time (kernels inside the long step ->
        # the SUSTAINED peak is the denominator)
        step_flops = sum(fl * n for _, n, _, fl, _ in prof)
        ach_step = step_flops / (ms_dev / args.steps * 1e-3) / 1e12
        roof["step_level"] = {"achieved": round(ach_step, 2), "peak": peaks["bf16_tflops_sustained"], "unit": "TFLOP/s",
                              "frac": round(ach_step / peaks["bf16_tflops_sustained"], 4), "peak_kind": "sustained",
                              "gflop_per_step": round(step_flops / 1e9, 2),
                              "serial_sum_us": round(tot, 1), "step_us": round(ms_dev / args.steps * 1e3, 1)}
        for label, n, us, fl, by in sorted(prof, key=lambda r: -r[1] * r[2])[:8]:
            top.append({"kernel": label, "share": round(n * us / tot, 4), "us_per_launch": round(us, 2), "launches_per_step": n,
                        "tflops": round(fl / (us * 1e-6) / 1e12, 2) if fl else None, "gbs": round(by / (us * 1e-6) / 1e9, 1) if by else None})
        if world == 1 and not args.no_large:
            large = large_batch_roofline(peaks)
    # ---- augmentation row (8a-15): raw strips resident on the device -> GPU weak/strong/standardise -> step ----
    aug = None
    if world == 1 and not args.no_aug:
        from semiseg_b200.augment import AugConfig, FixMatchBatcher
        acfg = AugConfig.from_config(cfg)
        acfg.target_length = L
        bat = FixMatchBatcher(eng, acfg, seed=cfg["seed"])
        raw = [(torch.randn(Bl, C, L, device=dev) * 0.4 + 0.1, devb[i][1], torch.randn(Bu, C, L, device=dev) * 0.4) for i in range(pool)]
        np.random.seed(cfg["seed"])
        for i in range(3):
            bat.load(*raw[i % pool]); eng.step(lr_at(epoch_f[0], tcfg))
        torch.cuda.synchronize()
        ev0, ev1, ev2 = (torch.cuda.Event(enable_timing=True) for _ in range(3))
        nrep = min(args.steps, 100)
        t0 = time.time()
        ev0.record()
        for i in range(nrep):
            bat.load(*raw[i % pool])
        ev1.record()
        for i in range(nrep):
            bat.load(*raw[i % pool]); eng.step(lr_at(epoch_f[0], tcfg))
        ev2.record()
        # overlapped: batch i+1 is augmented on a side stream (one captured graph launch) while step i runs
        ev4 = torch.cuda.Event(enable_timing=True)
        bat2 = FixMatchBatcher(eng, acfg, seed=cfg["seed"] + 1)
        bat2.prefetch_captured(*raw[0])
        bat2.commit(); eng.step(lr_at(epoch_f[0], tcfg)); bat2.prefetch_captured(*raw[1])
        torch.cuda.synchronize()
        ev3b = torch.cuda.Event(enable_timing=True)
        ev3b.record()
        for i in range(nrep):
            bat2.commit(); eng.step(lr_at(epoch_f[0], tcfg))
            bat2.prefetch_captured(*raw[(i + 1) % pool])
        bat2.commit()
        ev4.record()
        torch.cuda.synchronize()
        eng.read_stats()
        aug = {"augment_only_ms_per_batch": round(ev0.elapsed_time(ev1) / nrep, 4),
               "augment_only_samples_per_s": round(per_step * nrep / (ev0.elapsed_time(ev1) / 1e3), 1),
               "step_with_augment_ms": round(ev1.elapsed_time(ev2) / nrep, 4),
               "step_with_augment_samples_per_s": round(per_step * nrep / (ev1.elapsed_time(ev2) / 1e3), 1),
               "step_with_overlapped_augment_ms": round(ev3b.elapsed_time(ev4) / nrep, 4),
               "step_with_overlapped_augment_samples_per_s": round(per_step * nrep / (ev3b.elapsed_time(ev4) / 1e3), 1),
               "cpu_oracle_samples_per_s": cpu_augment_rate(C, L),
               "what": "FixMatchBatcher.load (weak Fourier resize-crop + labels, RandAugment 3 of 4 ops, standardise x3) from raw "
                       "strips resident in HBM into the engine's input arena; host makes only the scalar draws"}
    if rank != 0:
        _finish(world, list(model.runtime().engines.values()))
        return
    # working set, to justify the L2 policy
    plan = eng.plan_s
    ws = sum(t.numel() * t.element_size() for bufs in plan.blk_bufs for t in bufs.values())
    ws += sum(t.numel() * t.element_size() for t in (plan.c0, plan.p0, plan.ch, plan.ah, plan.dc0))
    ws += 4 * model.runtime().weights.params.numel() * 4
    cpu, parity = None, None
    if world == 1 and not args.no_cpu_baseline:
        cpu, first = cpu_baseline(args.workload, budget_s=15.0)
        parity = parity_at_bench_shape(args.workload, dtype, first)
    lib_gpu = None
    if world == 1 and not args.no_library:
        sys.path.insert(0, os.path.join(REPO, "tools"))
        import library_baseline
        base_, stem_ = WORKLOADS[args.workload][5:7]
        lib_gpu = library_baseline.run(C, L, Bl, Bu, base_, stem_, steps=min(max(args.steps, 10), 50))
        best = max(v["samples_per_s"] for k, v in lib_gpu.items() if isinstance(v, dict))
        lib_gpu["this_repo_over_best_library_mode"] = {"value": round(value / best, 2), "e2e": round(e2e / best, 2)}
    line = {
        "metric": "train samples/sec FixMatch 1D U-Net (resnet18+FCNHead segmentor) step",
        "value": round(value, 1), "unit": "samples/s", "n_gpus": world, "steps": args.steps, "warmup": max(args.warmup, 3),
        "ms_per_step": round(ms_dev / args.steps, 4), "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": args.dtype, "data": "synthetic LUDB-shaped strips (z-scored Gaussian, 4-class piecewise labels), random-init weights",
        "config": config_block(args.workload, world, want_sync_bn),
        "run": {"sync_bn_exchange": ("peer-memory kernel" if getattr(eng, "syncbn_p2p", False) else "nccl") if model.sync_bn else None,
                "cuda_graph": not args.no_graph, "working_set_mib": round(ws / 2**20, 1)},
        "e2e": {"value": round(e2e, 1), "unit": "samples/s", "h2d_bytes_per_step": eng.h2d_bytes(), "d2h_bytes_per_step": 32,
                "ms_per_step": round(ms_e2e / args.steps, 4),
                "through": f"algorithms.{algo}.train_one_epoch over list loaders of pinned host batch dicts"},
        "gpu_launches": eng.launches_per_step * args.steps,
        "launches_per_step": eng.launches_per_step,
        "clocks": sampler.summary(),
        "roofline": roof, "kernel_families": fams, "top_kernels": top,
        "loss_last": stats_dev[-1] if stats_dev else None,
    }
    if cpu is not None:
        line["cpu_baseline"] = cpu
    if parity is not None:
        line["parity_at_bench_shape"] = parity
    if lib_gpu is not None:
        line["library_gpu_baseline"] = lib_gpu
    if large is not None:
        line["large_batch_roofline"] = large
    if aug is not None:
        line["gpu_augmentation"] = aug
    if syncbn_other is not None:
        line["sync_bn_on" if syncbn_other["sync_bn"] else "sync_bn_off"] = syncbn_other
    if replicas_equal is not None:
        line["replicas_equal"] = replicas_equal
    if world == 1 and not args.no_aug and args.workload == DEFAULT_WORKLOAD:
        line["other_algorithms"] = other_algorithm_rates(cfg, dev, dtype, C, L, Bl, Bu, host[0])
    print(json.dumps(line), flush=True)
    _finish(world, list(model.runtime().engines.values()))


def other_algorithm_rates(cfg, dev, dtype, C, L, Bl, Bu, batch, steps=100):
    """Supplementary: the same network and per-GPU batch through the other step modes of the engine (SURVEY.md 8f
    rank 2) -- Mean-Teacher, ST++ (frozen hard teacher) and CPS (two models, two optimizer steps per step) --
    device-resident inputs, CUDA events, samples/s counted as in the headline (B_l + B_u strips per step)."""
    from algorithms.base import init_model_from_cfg
    from algorithms.mean_teacher import init_teacher
    from semiseg_b200.engine import CpsEngine
    from semiseg_b200.trainer import get_engine
    from utils.lr_sched import lr_at
    tcfg = dict(cfg["train"], ema_decay=0.99)
    x, y, uw, us = [t.to(dev) for t in batch]
    out = {}

    def rate(load, step):
        for _ in range(5):
            load()
            step()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(steps):
            load()
            step()
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / steps
        return {"ms_per_step": round(ms, 4), "samples_per_s": round((Bl + Bu) / (ms / 1e3), 1)}

    lr = lr_at(20.0, tcfg)
    torch.manual_seed(1)
    m = init_model_from_cfg(cfg).to(dev)
    t = init_teacher(cfg, m, dev)
    e = get_engine("mean_teacher", m, t, Bl, Bu, L, dtype, tcfg)
    out["mean_teacher"] = dict(rate(lambda: e.load_batch(x, y, uw, us), lambda: e.step(lr)), launches_per_step=e.launches_per_step)
    assert all(np.isfinite(s_["loss_total"]) for s_ in e.read_stats())
    m = init_model_from_cfg(cfg).to(dev)
    t = init_model_from_cfg(cfg).to(dev).eval()
    e = get_engine("stpp", m, t, Bl, Bu, L, dtype, tcfg)
    out["stpp"] = dict(rate(lambda: e.load_batch(x, y, uw), lambda: e.step(lr)), launches_per_step=e.launches_per_step)
    assert all(np.isfinite(s_["loss_total"]) for s_ in e.read_stats())
    m1, m2 = init_model_from_cfg(cfg).to(dev), init_model_from_cfg(cfg).to(dev)
    e1 = get_engine("cps", m1, m2, Bl, Bu, L, dtype, tcfg, external_pseudo=True)
    e2 = get_engine("cps", m2, m1, Bl, Bu, L, dtype, tcfg, external_pseudo=True)
    cps = CpsEngine(e1, e2)
    out["cps"] = dict(rate(lambda: cps.load_batch(x, y, uw), lambda: cps.step(lr)),
                      launches_per_step=e1.launches_per_step + e2.launches_per_step,
                      note="two models trained per step; the two training graphs run side by side")
    assert all(np.isfinite(s_["loss_total"]) for s_ in cps.read_stats())
    # other members of the backbone family behind the same registry: FixMatch step, same batch
    import copy
    for name, head_in in (("resnet34", 512), ("resnet50", 2048)):
        c2 = copy.deepcopy(cfg)
        c2["backbone"] = {name: c2["backbone"]["resnet18"]}
        c2["decode_head"]["FCNHead"]["in_channels"] = head_in
        mm = init_model_from_cfg(c2).to(dev)
        ee = get_engine("fixmatch", mm, None, Bl, Bu, L, dtype, dict(cfg["train"]))
        out[f"fixmatch_{name}"] = dict(rate(lambda: ee.load_batch(x, y, uw, us), lambda: ee.step(lr)),
                                       launches_per_step=ee.launches_per_step,
                                       params=int(sum(p_.numel() for p_ in mm.parameters())))
        assert all(np.isfinite(s_["loss_total"]) for s_ in ee.read_stats())
        del ee, mm
    # evaluation path: eval forward + fused metric tail as one graph, batch = B_l + B_u strips
    from semiseg_b200.evaluate import EvalEngine
    xe, ye = torch.cat((x, uw)), torch.cat((y, y))
    ev = EvalEngine(m1.runtime().weights, dtype, Bl + Bu, L)
    ev.refresh_weights()
    out["evaluate"] = dict(rate(lambda: None, lambda: ev.run(xe, ye)), launches_per_batch=ev.launches,
                           note="eval-mode forward with BN folded into the conv epilogues + ssb_eval_metrics; device-resident batch")
    return out


def large_batch_roofline(peaks, workload="fixmatch_resnet18w128_12x5000_b64+64"):
    """Supplementary evidence (not the bench value): the same kernels on BASELINE.json configs[4]'s shapes
    (12 x 5000, base width 128, 64+64 strips on this GPU = global batch 1024 on 8 GPUs), where the convs are
    tensor-bound instead of latency-bound.  Same steady-state per-launch timing as the main roofline entry."""
    from algorithms.base import init_model_from_cfg
    from semiseg_b200 import _lib
    from semiseg_b200.trainer import get_engine
    cfg, algo, C, L, Bl, Bu = load_cfg(workload)
    dev = torch.device("cuda", torch.cuda.current_device())
    torch.manual_seed(cfg["seed"])
    model = init_model_from_cfg(cfg).to(dev)
    eng = get_engine(algo, model, None, Bl, Bu, L, _lib.BF16, cfg["train"], use_graph=False)
    lab, unl = make_host_batch(cfg["seed"], 0, Bl, Bu, C, L)
    batch = [torch.from_numpy(lab["ecg"]).to(dev), torch.from_numpy(lab["target"]).to(dev),
             torch.from_numpy(unl["ecg"]).to(dev), torch.from_numpy(unl["ecg_aug"]).to(dev)]
    eng.load_batch(*batch); eng.step(1e-3)
    torch.cuda.synchronize()
    prof = steady_state_profile(eng, batch, 1e-3, 2, repeats=5)
    roof, fams, tot = roofline_from_profile(prof, peaks)
    ridge = peaks["bf16_tflops"] * 1e12 / (peaks["hbm_gbs"] * 1e9)
    tb = [(l, n, us, fl, by) for l, n, us, fl, by in prof if fl and by and fl / by > ridge and family_of(l).startswith("conv_tn")]
    tfl = sum(fl * n for _, n, _, fl, _ in tb)
    tus = sum(us * n for _, n, us, _, _ in tb)
    best = max(tb, key=lambda r: r[3] / r[2]) if tb else None
    out = {"workload": workload, "serial_us_per_step": round(tot, 1), "roofline": roof, "kernel_families": fams}
    if tb:
        out["tensor_bound_convs"] = {"launches_per_step": sum(n for _, n, _, _, _ in tb), "tflops": round(tfl / (tus * 1e-6) / 1e12, 1),
                                     "frac_of_peak": round(tfl / (tus * 1e-6) / 1e12 / peaks["bf16_tflops"], 4),
                                     "frac_of_sustained_peak": round(tfl / (tus * 1e-6) / 1e12 / peaks["bf16_tflops_sustained"], 4),
                                     "best": {"kernel": best[0], "tflops": round(best[3] / (best[2] * 1e-6) / 1e12, 1)}}
    del eng, model
    torch.cuda.empty_cache()
    return out


def cpu_augment_rate(C, L, budget_s=2.0):
    """The reference's per-item augmentation (oracle port: numpy/scipy float64, one core) on this host."""
    from oracle import augment_oracle as A
    rng = np.random.default_rng(0)
    x = rng.standard_normal((C, L))
    y = rng.integers(0, 4, (1, L))
    np.random.seed(0)
    t0, n = time.time(), 0
    while time.time() - t0 < budget_s:
        A.labeled_item(x, y, L)
        A.unlabeled_item(x, L)
        n += 2
    return round(n / (time.time() - t0), 1)


PARITY_THRESH = 0.3    # at random init no position reaches the YAML's 0.8: the parity step uses a threshold near the median


def cpu_step_fn(workload, dropout_off_first=False):
    """One FixMatch/Mean-Teacher step of the oracle port on the host CPU (fp32, all threads).
    dropout_off_first: the FIRST call runs without dropout at conf_thresh = PARITY_THRESH and returns the record that
    parity_at_bench_shape compares the CUDA path with (losses, mask ratio, every parameter gradient)."""
    from algorithms.base import init_model_from_cfg
    from oracle import segnet_oracle as O
    O.FAST_KERNELS = True    # BatchNorm / max-pool / upsample / CE through the ATen kernels the reference's modules call
    cfg, algo, C, L, Bl, Bu = load_cfg(workload)
    torch.manual_seed(cfg["seed"])
    model = init_model_from_cfg(cfg)
    arch = O.Arch.from_config(cfg)
    tr = O.OracleTrainer({k: v.detach() for k, v in model.state_dict().items()}, arch, cfg["train"], dtype=torch.float32)
    lab, unl = make_host_batch(cfg["seed"], 0, Bl, Bu, C, L)
    lab = {k: torch.from_numpy(v) for k, v in lab.items()}
    unl = {k: torch.from_numpy(v) for k, v in unl.items()}
    g = torch.Generator().manual_seed(0)
    Lh = O.stage_lengths(arch, L)[-1]

    calls = [0]

    def step():
        calls[0] += 1
        if dropout_off_first and calls[0] == 1:
            thr0 = tr.cfg.get("conf_thresh")
            tr.cfg = dict(tr.cfg, conf_thresh=PARITY_THRESH)
            fn = tr.mean_teacher_step if algo == "mean_teacher" else tr.fixmatch_step
            st_ = fn(lab["ecg"], lab["target"], unl["ecg"], unl["ecg_aug"], 1e-3, dropout_mask=None)
            tr.cfg = dict(tr.cfg, conf_thresh=thr0)
            return {"stats": st_, "grads": {k: v.clone() for k, v in tr.grads.items()}, "pnames": list(tr.pnames)}
        dm = (torch.rand(Bl + Bu, arch.head_channels, Lh, generator=g) >= arch.dropout_ratio)
        if algo == "mean_teacher":
            return tr.mean_teacher_step(lab["ecg"], lab["target"], unl["ecg"], unl["ecg_aug"], 1e-3, dropout_mask=dm)
        return tr.fixmatch_step(lab["ecg"], lab["target"], unl["ecg"], unl["ecg_aug"], 1e-3, dropout_mask=dm)
    return step, Bl + Bu


def parity_at_bench_shape(workload, dtype, first):
    """Outside every timed region: ONE step of the CUDA path (captured graph, the benchmarked batch and network) from
    the same seeded weights on the same batch as the CPU oracle's first step (no dropout, threshold PARITY_THRESH),
    compared with it -- the oracle is the checker here, never the thing measured."""
    import copy
    from algorithms.base import init_model_from_cfg
    from algorithms.mean_teacher import init_teacher
    from semiseg_b200 import _lib
    from semiseg_b200.trainer import get_engine
    cfg, algo, C, L, Bl, Bu = load_cfg(workload)
    cfg = copy.deepcopy(cfg)
    cfg["decode_head"]["FCNHead"]["dropout_ratio"] = 0.0
    dev = torch.device("cuda", torch.cuda.current_device())
    torch.manual_seed(cfg["seed"])
    model = init_model_from_cfg(cfg).to(dev)
    teacher = init_teacher(cfg, model, dev) if algo == "mean_teacher" else None
    tcfg = dict(cfg["train"], conf_thresh=PARITY_THRESH)
    eng = get_engine(algo, model, teacher, Bl, Bu, L, dtype, tcfg, use_graph=True)
    lab, unl = make_host_batch(cfg["seed"], 0, Bl, Bu, C, L)
    eng.load_batch(torch.from_numpy(lab["ecg"]), torch.from_numpy(lab["target"]), torch.from_numpy(unl["ecg"]),
                   torch.from_numpy(unl["ecg_aug"]))
    eng.step(1e-3)
    s, = eng.read_stats()
    so = first["stats"]
    tol = 2e-2 if dtype == _lib.BF16 else 1e-4
    out = {"tolerance": tol, "threshold": PARITY_THRESH, "cuda": {k: round(float(v), 6) for k, v in s.items()},
           "oracle_fp32_cpu": {k: round(float(v), 6) for k, v in so.items()}}
    ok = all(abs(s[k] - so[k]) <= tol * max(1.0, abs(so[k])) for k in so)
    grads = model.runtime().weights.param_views(model.runtime().state.grads)
    g = torch.cat([grads[n].flatten().cpu().double() for n in first["pnames"]])
    r = torch.cat([first["grads"][n].flatten().double() for n in first["pnames"]])
    out["global_gradient_rel_err"] = float((g - r).norm() / (r.norm() + 1e-30))
    head = ["decode_head.cls_seg.weight", "decode_head.cls_seg.bias"]
    out["head_gradient_rel_err"] = max(float((grads[n].cpu().double() - first["grads"][n].double()).norm() /
                                             (first["grads"][n].double().norm() + 1e-30)) for n in head)
    ok = ok and out["head_gradient_rel_err"] <= tol
    out["ok"] = bool(ok)
    del eng, model
    return out


def cpu_baseline(workload, budget_s=15.0):
    """(cpu_baseline dict, the oracle's first-step record for parity_at_bench_shape)"""
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    step, per = cpu_step_fn(workload, dropout_off_first=True)
    first = step()
    t0 = time.time()
    n = 0
    while time.time() - t0 < budget_s and n < 100:
        step()
        n += 1
    dt = time.time() - t0
    ratio = None
    rf = os.path.join(REPO, "profiles", "r2_cpu_arm_vs_reference.json")
    if os.path.exists(rf):
        ratio = json.load(open(rf)).get("port_over_reference")
    return {"value": round(per * n / dt, 2), "unit": "samples/s", "cores": torch.get_num_threads(), "kind": "port",
            "sample": f"{n} steps of the same workload (oracle port of fixmatch.train_one_epoch body, torch CPU fp32, "
                      f"{dt:.1f} s after 1 warm-up step)",
            "port_ms_over_reference_ms_in_build_container": ratio}, first


def run_reference(args):
    rank = int(os.environ.get("RANK", 0))
    if rank != 0:
        return
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    step, per = cpu_step_fn(args.workload)
    for _ in range(max(args.warmup, 1)):
        step()
    t0 = time.time()
    for _ in range(args.steps):
        step()
    dt = time.time() - t0
    v = round(per * args.steps / dt, 2)
    cfg, algo, C, L, Bl, Bu = load_cfg(args.workload)
    print(json.dumps({
        "impl": "reference", "metric": "train samples/sec FixMatch 1D U-Net (resnet18+FCNHead segmentor) step",
        "value": v, "unit": "samples/s", "n_gpus": args.gpus, "steps": args.steps, "warmup": max(args.warmup, 1),
        "ms_per_step": round(dt / args.steps * 1e3, 2), "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f32", "data": "synthetic LUDB-shaped strips, random-init weights",
        "config": config_block(args.workload, args.gpus, bool(cfg["ddp"].get("sync_bn", True)) if args.sync_bn is None else bool(args.sync_bn)),
        "cpu_baseline": {"value": v, "unit": "samples/s", "cores": torch.get_num_threads(), "kind": "port",
                         "sample": f"{args.steps} steps of the workload on the host CPU (oracle port of the reference step; the "
                                   "reference is Python and is not shipped to the GPU box)"},
        "e2e": {"value": v, "unit": "samples/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=None)
    ap.add_argument("--warmup", type=int, default=None)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--workload", default=DEFAULT_WORKLOAD, choices=sorted(WORKLOADS))
    ap.add_argument("--dtype", default="bf16", choices=["bf16", "fp32"])
    ap.add_argument("--sync-bn", type=int, default=None, help="1: SyncBatchNorm statistic exchange; 0: per-rank BN "
                    "(ddp.sync_bn: false); default: the YAML's ddp.sync_bn (true, like the reference)")
    ap.add_argument("--no-graph", action="store_true")
    ap.add_argument("--preroll-s", type=float, default=0.4, help="seconds of untimed steps before the warm-up steps (GPU out of idle clocks)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-syncbn-other", action="store_true", help="N>1: skip timing the step with the OTHER SyncBN setting "
                    "(reported as sync_bn_off / sync_bn_on next to the headline)")
    ap.add_argument("--no-library", action="store_true", help="skip the torch.nn / cuDNN eager baseline of the same step")
    ap.add_argument("--no-aug", action="store_true", help="skip the supplementary GPU-augmentation measurement")
    ap.add_argument("--no-large", action="store_true", help="skip the supplementary large-batch roofline block")
    ap.add_argument("--profile-mode", action="store_true", help="warm-up + K plain steps only (for ncu)")
    a = ap.parse_args()
    if a.impl == "reference":
        a.steps = a.steps or 10
        a.warmup = a.warmup if a.warmup is not None else 1
        run_reference(a)
    else:
        a.steps = a.steps or 200
        a.warmup = a.warmup if a.warmup is not None else 20
        run_b200(a)


if __name__ == "__main__":
    main()
