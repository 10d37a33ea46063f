#!/usr/bin/env python
"""Steady-state cost of every distinct launch of one training step.

Records the C-ABI launches of one eager FixMatch step, then captures, for each distinct launch,
a CUDA graph of R back-to-back (stream-ordered, hence dependent) repeats and times its replay:
µs per launch with a warm L2 and graph launch gaps -- the floor a latency-bound step is made of.
Prints one JSON line per kernel label plus the sum over the step (serial lower bound).

  python tools/latency_probe.py [--workload NAME] [--repeats 40]        (SSB_PDL=0/1 to compare)
"""
import argparse
import json
import os
import sys

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, REPO)
import bench  # noqa: E402  (sets sys.path for the package)
import torch  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--workload", default=bench.DEFAULT_WORKLOAD)
    ap.add_argument("--repeats", type=int, default=40)
    a = ap.parse_args()
    from algorithms.base import init_model_from_cfg
    from semiseg_b200 import _lib
    from semiseg_b200.trainer import get_engine
    cfg, algo, C, L, Bl, Bu = bench.load_cfg(a.workload)
    dev = torch.device("cuda", 0)
    torch.cuda.set_device(0)
    torch.manual_seed(0)
    model = init_model_from_cfg(cfg).to(dev)
    os.environ["SSB_MULTI_STREAM"] = "0"
    eng = get_engine(algo, model, None, Bl, Bu, L, _lib.BF16, cfg["train"], use_graph=False)
    lab, unl = bench.make_host_batch(0, 0, Bl, Bu, C, L)
    batch = [torch.from_numpy(lab["ecg"]).to(dev), torch.from_numpy(lab["target"]).to(dev),
             torch.from_numpy(unl["ecg"]).to(dev), torch.from_numpy(unl["ecg_aug"]).to(dev)]
    eng.load_batch(*batch)
    eng.step(1e-3)
    torch.cuda.synchronize()
    rec = []

    def hook(name, args):
        rec.append((name, args))
        _lib.raw_call(name, *args)
    _lib._hook = hook
    eng.load_batch(*batch)
    eng.step(1e-3)
    _lib._hook = None
    torch.cuda.synchronize()
    uniq = {}
    for name, args in rec:
        fl, by, label = bench.launch_cost(name, args, 2)
        d = uniq.setdefault(label, {"name": name, "args": args, "n": 0, "flops": fl, "bytes": by})
        d["n"] += 1
    lib = _lib.load()
    s = torch.cuda.Stream()
    total = 0.0
    out = []
    with torch.cuda.stream(s):
        for label, d in uniq.items():
            name, args = d["name"], list(d["args"])
            args[-1] = s.cuda_stream
            n0 = lib.ssb_launch_count()
            g = torch.cuda.CUDAGraph()
            with torch.cuda.graph(g, stream=s, capture_error_mode="thread_local"):
                for _ in range(a.repeats):
                    _lib.raw_call(name, *args)
            per_call = (lib.ssb_launch_count() - n0) / a.repeats
            for _ in range(3):
                g.replay()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(s)
            for _ in range(5):
                g.replay()
            e1.record(s)
            e1.synchronize()
            us = e0.elapsed_time(e1) * 1e3 / (5 * a.repeats)
            total += us * d["n"]
            out.append({"kernel": label, "us": round(us, 2), "n_per_step": d["n"], "kernels_per_call": per_call,
                        "tflops": round(d["flops"] / us / 1e6, 1) if d["flops"] else None,
                        "gbs": round(d["bytes"] / us / 1e3, 1) if d["bytes"] else None})
    for o in sorted(out, key=lambda o: -o["us"] * o["n_per_step"]):
        print(json.dumps(o))
    print(json.dumps({"serial_sum_us": round(total, 1), "launch_calls": len(rec), "pdl": os.environ.get("SSB_PDL", "1")}))


if __name__ == "__main__":
    main()
