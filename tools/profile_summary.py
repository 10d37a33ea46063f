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
    d = collections.defaultdict(list)
    grids = collections.defaultdict(set)
    for r in data:
        if len(r) <= vi:
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


if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("mode", choices=["launches", "full"])
    ap.add_argument("src")
    ap.add_argument("dst")
    ap.add_argument("--note", default="")
    a = ap.parse_args()
    (launches if a.mode == "launches" else full)(a.src, a.dst, a.note)
