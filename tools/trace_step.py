#!/usr/bin/env python
"""Timeline of ONE replayed step graph from inside the kernels (development tool).

  make -C semi-seg-ecg_b200/csrc TRACE=1
  SSB_LIB=semi-seg-ecg_b200/lib/libsemiseg_b200_trace.so python tools/trace_step.py [--workload W] [--out profiles/x.md]

The trace build makes block (0,0,0) of every launch record %globaltimer at kernel entry and when it gets past its
dependency wait (griddepcontrol.wait), tagged with the source line of that wait.  This script warms the engine up,
resets the records, replays one step, and prints the launches ordered by start time: offset from the step's first
kernel, how long the block sat in its dependency wait, and the distance to the next start of the same chain.
Kernel END times are not recorded; for a chain of dependent kernels start(i+1) - start(i) is kernel i's latency
contribution (duration + launch gap).
"""
import argparse
import ctypes as C
import os
import re
import sys

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, REPO)
sys.path.insert(0, os.path.join(REPO, "semi-seg-ecg_b200", "src"))

import torch  # noqa: E402

TUS = ["augment", "bn", "conv_simt", "conv_sm100", "head_loss", "optim"]


class Rec(C.Structure):
    _fields_ = [("t_entry", C.c_uint64), ("t_start", C.c_uint64), ("line", C.c_uint32), ("grid", C.c_uint32),
                ("block", C.c_uint32), ("pad", C.c_uint32)]


def kernel_at(tu, line, cache={}):
    """name of the __global__ function enclosing `line` of csrc/<tu>.cu"""
    if tu not in cache:
        cache[tu] = open(os.path.join(REPO, "semi-seg-ecg_b200", "csrc", tu + ".cu")).read().splitlines()
    src = cache[tu]
    for i in range(min(line, len(src)) - 1, -1, -1):
        if "__global__" in src[i]:
            m = re.search(r"\b(\w+)\s*\(", " ".join(src[i:i + 3]).split("__global__", 1)[1].replace("__launch_bounds__", ""))
            j = i
            txt = " ".join(src[j:j + 4])
            names = re.findall(r"\b([a-z_0-9]+_kernel\w*|zero_parity_rows\w*)\b", txt)
            return names[0] if names else (m.group(1) if m else "?")
    return "?"


def dump(lib, reset):
    out = []
    buf = (Rec * 4096)()
    for tu in TUS:
        fn = getattr(lib, "ssb_trace_dump_" + tu)
        fn.argtypes = [C.c_void_p, C.c_int, C.c_int]
        fn.restype = C.c_int
        n = fn(C.cast(buf, C.c_void_p), 4096, 1 if reset else 0)
        for i in range(n):
            r = buf[i]
            out.append((r.t_start, r.t_entry, tu, r.line, r.grid, r.block, r.pad))
    return sorted(out)


def main():
    import bench
    from algorithms.base import init_model_from_cfg
    from algorithms.mean_teacher import init_teacher
    from semiseg_b200 import _lib
    from semiseg_b200.trainer import get_engine
    ap = argparse.ArgumentParser()
    ap.add_argument("--workload", default=bench.DEFAULT_WORKLOAD)
    ap.add_argument("--out", default="")
    a = ap.parse_args()
    lib = _lib.load()
    assert hasattr(lib, "ssb_trace_dump_bn"), "load the trace build: SSB_LIB=.../libsemiseg_b200_trace.so"
    cfg, algo, Cl, L, Bl, Bu = bench.load_cfg(a.workload)
    dev = torch.device("cuda", 0)
    torch.cuda.set_device(0)
    torch.manual_seed(cfg["seed"])
    model = init_model_from_cfg(cfg).to(dev)
    teacher = init_teacher(cfg, model, dev) if algo == "mean_teacher" else None
    eng = get_engine(algo, model, teacher, Bl, Bu, L, _lib.BF16, cfg["train"])
    lab, unl = bench.make_host_batch(cfg["seed"], 0, Bl, Bu, Cl, L)
    batch = [torch.from_numpy(lab["ecg"]).to(dev), torch.from_numpy(lab["target"]).to(dev),
             torch.from_numpy(unl["ecg"]).to(dev), torch.from_numpy(unl["ecg_aug"]).to(dev)]
    for _ in range(10):
        eng.load_batch(*batch)
        eng.step(1e-3)
    torch.cuda.synchronize()
    dump(lib, True)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    eng.load_batch(*batch)
    torch.cuda.synchronize()
    e0.record()
    eng.step(1e-3)
    e1.record()
    torch.cuda.synchronize()
    allrecs = dump(lib, True)
    recs = [r[:6] for r in allrecs if r[6] == 0]
    marks = [r for r in allrecs if r[6] == 1]
    t0 = recs[0][0]
    lines = [f"# in-kernel timeline of one replayed step: {a.workload}", "",
             f"{len(recs)} launches recorded; step {e0.elapsed_time(e1) * 1e3:.1f} us by CUDA events (one replay, trace build); "
             f"first kernel start -> last kernel start {(recs[-1][0] - t0) / 1e3:.1f} us.", "",
             "`start` = block (0,0,0) past its dependency wait, us after the step's first kernel; `waited` = time that block sat "
             "in griddepcontrol.wait (scheduled early under the predecessor = PDL at work); `to next` = distance to the next "
             "start in the list (NOT this kernel's duration when branches interleave).", "",
             "| # | start us | waited us | to next us | kernel | grid | block |", "|---:|---:|---:|---:|---|---:|---:|"]
    for i, (ts, te, tu, line, grid, block) in enumerate(recs):
        nxt = (recs[i + 1][0] - ts) / 1e3 if i + 1 < len(recs) else 0.0
        name = kernel_at(tu, line)
        # phase marks of this launch: same kernel function, after this start and before the next start of the same function
        later = [r[0] for r in recs[i + 1:] if r[2] == tu and kernel_at(r[2], r[3]) == name]
        lim = later[0] if later else ts + 10**9
        mk = sorted((m[0], m[3]) for m in marks if m[2] == tu and ts <= m[0] < lim and (kernel_at(tu, m[3]) == name or name.startswith("conv_tn")))
        ms = " ".join(f"+{(mt - ts) / 1e3:.1f}@{ml}" for mt, ml in mk)
        lines.append(f"| {i} | {(ts - t0) / 1e3:.1f} | {(ts - te) / 1e3:.1f} | {nxt:.1f} | `{name}` ({tu}.cu:{line}) {ms} | {grid} | {block} |")
    txt = "\n".join(lines) + "\n"
    if a.out:
        open(a.out, "w").write(txt)
    print(txt)


if __name__ == "__main__":
    main()
