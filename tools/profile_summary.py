#!/usr/bin/env python
"""Turn ncu output brought back in gpurun_out/ into the committed summaries under profiles/.

  python tools/profile_summary.py launches gpurun_out/launches.csv profiles/r1_launches.md [--note "..."]
  python tools/profile_summary.py full gpurun_out/prof.ncu-rep profiles/r1_conv_tn_full.md [--note "..."]

`launches`: the `--metrics gpu__time_duration.sum` launch list -> per-kernel count, mean duration and
SHARE of the summed device time (cold-cache, serialised: compare shares, not absolutes).
`full`: one `--set full` capture -> the handful of counters the roofline argument uses, per launch.
"""
import argparse
import collections
import csv
import io
import subprocess

FULL_KEYS = [
    "gpu__time_duration.sum", "launch__grid_size", "launch__block_size", "launch__registers_per_thread",
    "launch__shared_mem_per_block_dynamic", "dram__bytes_read.sum", "dram__bytes_write.sum",
    "dram__throughput.avg.pct_of_peak_sustained_elapsed", "lts__t_bytes.sum",
    "lts__throughput.avg.pct_of_peak_sustained_elapsed", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
    "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
    "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed",
    "sm__warps_active.avg.pct_of_peak_sustained_active", "sm__inst_executed.sum",
    "smsp__cycles_active.avg", "sm__cycles_elapsed.max",
]


def launches(src, dst, note):
    rows = list(csv.reader(open(src)))
    hi = [i for i, r in enumerate(rows) if r and r[0] == "ID"][0]
    hdr, data = rows[hi], rows[hi + 1:]
    ki, vi, gi, bi = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Grid Size"), hdr.index("Block Size")
    mi = hdr.index("Metric Name")
    d = collections.defaultdict(list)
    grids = collections.defaultdict(set)
    for r in data:
        if len(r) <= vi or r[mi] != "gpu__time_duration.sum":      # (a pass may carry more metrics per launch)
            continue
        name = r[ki].split("(")[0].replace("void ", "").replace("<unnamed>::", "")
        d[name].append(float(r[vi].replace(",", "")))
        grids[name].add(f"{r[gi]}x{r[bi]}")
    tot = sum(sum(v) for v in d.values())
    n = sum(len(v) for v in d.values())
    out = io.StringIO()
    out.write(f"# ncu launch list: {src}\n\n{note}\n\n")
    out.write(f"{n} launches, summed device time {tot / 1e3:.1f} us (cold-cache, serialised under ncu; shares are what count)\n\n")
    out.write("| kernel | launches | mean us | share | grids (grid x block) |\n|---|---:|---:|---:|---|\n")
    for k, v in sorted(d.items(), key=lambda kv: -sum(kv[1])):
        g = sorted(grids[k])
        gs = ", ".join(g[:3]) + (" ..." if len(g) > 3 else "")
        out.write(f"| `{k}` | {len(v)} | {sum(v) / len(v) / 1e3:.2f} | {sum(v) / tot * 100:.1f}% | {gs} |\n")
    open(dst, "w").write(out.getvalue())
    print(out.getvalue())


def full(src, dst, note):
    raw = subprocess.run(["ncu", "-i", src, "--page", "raw", "--csv"], capture_output=True, text=True, check=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    hdr, units, data = rows[0], rows[1], rows[2:]
    out = io.StringIO()
    out.write(f"# ncu --set full: {src}\n\n{note}\n\n")
    for r in data:
        out.write(f"## `{r[hdr.index('Kernel Name')][:110]}`  grid {r[hdr.index('Grid Size')]} block {r[hdr.index('Block Size')]}\n\n")
        out.write("| metric | value | unit |\n|---|---:|---|\n")
        for k in FULL_KEYS:
            if k in hdr:
                i = hdr.index(k)
                out.write(f"| {k} | {r[i]} | {units[i]} |\n")
        out.write("\n")
    open(dst, "w").write(out.getvalue())
    print(out.getvalue())


# kernel function name (substring) -> index of bench.py's FAMILIES
KERNEL_FAMILY = [("stem_conv", 3), ("conv_tn", 0), ("zero_parity_rows", 0), ("tap_gemm", 0), ("conv_wgrad", 1), ("wgrad_kernel", 1),
                 ("bn_", 2), ("stem_bn", 2), ("stem_bwd", 2)]


def traffic(src, dst, note, workload, steps):
    """`ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --print-units base --csv` of
    `bench.py --profile-mode --steps <steps>` -> DRAM bytes per kernel family per step, merged into the JSON file
    `dst` under the workload's name (bench.py reads it for roofline.traffic)."""
    import json
    import os
    import sys
    sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
    from bench import FAMILIES
    rows = list(csv.reader(open(src)))
    hi = [i for i, r in enumerate(rows) if r and r[0] == "ID"][0]
    hdr, data = rows[hi], rows[hi + 1:]
    ii, ki, mi, vi = hdr.index("ID"), hdr.index("Kernel Name"), hdr.index("Metric Name"), hdr.index("Metric Value")
    per = collections.OrderedDict()
    for r in data:
        if len(r) <= vi:
            continue
        d = per.setdefault(r[ii], {"name": r[ki].split("(")[0].replace("void ", "").replace("<unnamed>::", "")})
        d[r[mi]] = float(r[vi].replace(",", ""))
    fams = collections.defaultdict(lambda: {"dram_bytes_per_step": 0.0, "launches_per_step": 0.0, "time_us_per_step": 0.0})
    kern = collections.defaultdict(lambda: {"dram_bytes": 0.0, "launches": 0, "time_us": 0.0})
    for d in per.values():
        fam = FAMILIES[-1][0]
        for sub, idx in KERNEL_FAMILY:
            if sub in d["name"]:
                fam = FAMILIES[idx][0]
                break
        b = d.get("dram__bytes_read.sum", 0.0) + d.get("dram__bytes_write.sum", 0.0)
        f = fams[fam]
        f["dram_bytes_per_step"] += b / steps
        f["launches_per_step"] += 1.0 / steps
        f["time_us_per_step"] += d.get("gpu__time_duration.sum", 0.0) / 1e3 / steps
        k = kern[d["name"]]
        k["dram_bytes"] += b
        k["launches"] += 1
        k["time_us"] += d.get("gpu__time_duration.sum", 0.0) / 1e3
    out = json.load(open(dst)) if os.path.exists(dst) else {}
    out[workload] = {"source": f"{note} ({src}; {len(per)} launches = {steps} steps; cold-cache, serialised under ncu)",
                     "families": {k: {kk: round(vv, 2) for kk, vv in v.items()} for k, v in fams.items()},
                     "kernels": {k: {"launches": v["launches"], "dram_bytes_per_launch": round(v["dram_bytes"] / v["launches"]),
                                     "us_per_launch": round(v["time_us"] / v["launches"], 2)} for k, v in
                                 sorted(kern.items(), key=lambda kv: -kv[1]["time_us"])}}
    json.dump(out, open(dst, "w"), indent=1)
    print(json.dumps(out[workload]["families"], indent=1))


if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("mode", choices=["launches", "full", "traffic"])
    ap.add_argument("src")
    ap.add_argument("dst")
    ap.add_argument("--note", default="")
    ap.add_argument("--workload", default="fixmatch_resnet18_ludb_1x2500_b16+16")
    ap.add_argument("--steps", type=int, default=2)
    a = ap.parse_args()
    if a.mode == "traffic":
        traffic(a.src, a.dst, a.note, a.workload, a.steps)
    else:
        (launches if a.mode == "launches" else full)(a.src, a.dst, a.note)
