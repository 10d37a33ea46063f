#!/usr/bin/env python
"""How fast is bench.py's CPU arm (the oracle port) compared with the UNMODIFIED reference on the same host cores?

Build container only (needs /root/reference).  Two child processes (the repo and the reference use the same top-level
package names): one times `fixmatch.train_one_epoch` of the reference through oracle/ref_harness.py on list-backed
loaders, the other times `bench.cpu_step_fn` (what `--impl reference` and `cpu_baseline` run) on the same workload
(BASELINE configs[1]: 16+16 strips of 1 x 2500, resnet18 + FCNHead, fp32, all host threads).  Prints one JSON line;
the committed copy is profiles/r2_cpu_arm_vs_reference.json.
"""
import json
import os
import subprocess
import sys

REPO = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

REF_CHILD = r"""
import sys, time, json, os
sys.path.insert(0, %(repo)r)
import importlib.util, numpy as np, torch
from oracle.ref_harness import ListLoader, import_reference
spec = importlib.util.spec_from_file_location("ssb_synthetic", os.path.join(%(repo)r, "semi-seg-ecg_b200", "src", "semiseg_b200", "synthetic.py"))
synthetic = importlib.util.module_from_spec(spec); spec.loader.exec_module(synthetic)
R = import_reference()
torch.set_num_threads(os.cpu_count())
cfg = {"backbone": {"resnet18": {"num_leads": 1, "num_stages": 4, "out_indices": [0, 1, 2, 3], "dilations": [1, 1, 1, 1],
       "strides": [1, 2, 2, 2], "deep_stem": False, "avg_down": False, "contract_dilation": False, "stem_channels": 64, "base_channels": 64}},
       "decode_head": {"FCNHead": {"in_channels": 512, "in_index": 3, "channels": 128, "num_convs": 1, "concat_input": False,
       "dropout_ratio": 0.1, "num_classes": 4, "align_corners": False}}}
tc = {"epochs": 100, "accum_iter": 1, "warmup_epochs": 10, "min_lr": 1e-4, "blr": None, "lr": 1e-3, "weight_decay": 0.05, "max_norm": None,
      "layer_decay": None, "optimizer": "adamw", "optimizer_kwargs": {"betas": [0.9, 0.999]}, "conf_thresh": 0.8}
torch.manual_seed(0)
model = R.base.init_model_from_cfg(cfg)
n = %(steps)d
a, b = synthetic.make_batch(0, 16, 16, 1, 2500)
lab = ListLoader([{k: torch.from_numpy(v) for k, v in a.items()}] * n)
unl = ListLoader([{k: torch.from_numpy(v) for k, v in b.items()}] * n)
opt = R.optimizer.get_optimizer_from_config(tc, model.parameters())
scaler = R.misc.NativeScalerWithGradNormCount()
dev = torch.device("cpu")
ts = []
for rep in range(%(reps)d + 1):
    t0 = time.time()
    R.fixmatch.train_one_epoch(model, lab, unl, opt, dev, 20, scaler, None, False, tc)
    ts.append((time.time() - t0) / n * 1e3)
print("RESULT " + json.dumps({"ms_per_step": min(ts[1:]), "all": ts[1:]}))
"""

PORT_CHILD = r"""
import sys, time, json, os
sys.path.insert(0, %(repo)r)
import torch
import bench
torch.set_num_threads(os.cpu_count())
step, per = bench.cpu_step_fn(bench.DEFAULT_WORKLOAD)
n = %(steps)d
ts = []
for rep in range(%(reps)d + 1):
    t0 = time.time()
    for _ in range(n):
        step()
    ts.append((time.time() - t0) / n * 1e3)
print("RESULT " + json.dumps({"ms_per_step": min(ts[1:]), "all": ts[1:]}))
"""


def run(child, **kw):
    out = subprocess.run([sys.executable, "-c", child % dict(repo=REPO, **kw)], capture_output=True, text=True, check=True).stdout
    return json.loads([ln for ln in out.splitlines() if ln.startswith("RESULT ")][-1][7:])


def main():
    steps, reps = 5, 4
    refs, ports = [], []
    for _ in range(2):          # alternate the arms: the build container shares its host (+-30 % run to run)
        refs.append(run(REF_CHILD, steps=steps, reps=reps))
        ports.append(run(PORT_CHILD, steps=steps, reps=reps))
    ref = {"ms_per_step": min(r["ms_per_step"] for r in refs), "all": sum((r["all"] for r in refs), [])}
    port = {"ms_per_step": min(r["ms_per_step"] for r in ports), "all": sum((r["all"] for r in ports), [])}
    print(json.dumps({"workload": "fixmatch_resnet18_ludb_1x2500_b16+16, fp32, CPU", "cores": os.cpu_count(),
                      "reference_ms_per_step": round(ref["ms_per_step"], 1), "port_ms_per_step": round(port["ms_per_step"], 1),
                      "port_over_reference": round(port["ms_per_step"] / ref["ms_per_step"], 3),
                      "reference_runs": [round(t, 1) for t in ref["all"]], "port_runs": [round(t, 1) for t in port["all"]],
                      "how": f"fastest of 2 x {reps} runs of {steps} steps (one warm-up run each), arms alternated, each in its own process"}))


if __name__ == "__main__":
    main()
