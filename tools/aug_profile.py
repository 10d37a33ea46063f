#!/usr/bin/env python
"""A few FixMatchBatcher.load calls on the benchmark shape, nothing else (for `ncu -k regex:aug_`)."""
import os
import sys

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, REPO)
sys.path.insert(0, os.path.join(REPO, "semi-seg-ecg_b200", "src"))
import numpy as np  # noqa: E402
import torch  # noqa: E402

import bench  # noqa: E402
from algorithms.base import init_model_from_cfg  # noqa: E402
from semiseg_b200 import _lib  # noqa: E402
from semiseg_b200.augment import AugConfig, FixMatchBatcher  # noqa: E402
from semiseg_b200.trainer import get_engine  # noqa: E402

cfg, algo, C, L, Bl, Bu = bench.load_cfg(bench.DEFAULT_WORKLOAD)
dev = torch.device("cuda", 0)
model = init_model_from_cfg(cfg).to(dev)
eng = get_engine(algo, model, None, Bl, Bu, L, _lib.BF16, cfg["train"])
acfg = AugConfig.from_config(cfg)
acfg.target_length = L
bat = FixMatchBatcher(eng, acfg, seed=0)
np.random.seed(0)
raw = (torch.randn(Bl, C, L, device=dev) * 0.4 + 0.1, torch.randint(0, 4, (Bl, L), device=dev), torch.randn(Bu, C, L, device=dev) * 0.4)
n = int(sys.argv[1]) if len(sys.argv) > 1 else 4
for _ in range(n):
    bat.load(*raw)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(20):
    bat.load(*raw)
e1.record()
torch.cuda.synchronize()
print(f"augment: {e0.elapsed_time(e1) / 20 * 1e3:.1f} us per batch")
