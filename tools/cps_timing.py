"""Times the CPS step (two models) with the two training graphs serial vs side by side (SSB_CPS_CONCURRENT)."""
import os
import sys

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(REPO, "semi-seg-ecg_b200", "src"))
sys.path.insert(0, REPO)
import torch  # noqa: E402

import bench  # noqa: E402
from algorithms.base import init_model_from_cfg  # noqa: E402
from semiseg_b200 import _lib  # noqa: E402
from semiseg_b200.engine import CpsEngine  # noqa: E402
from semiseg_b200.trainer import get_engine  # noqa: E402

cfg, algo, C, L, Bl, Bu = bench.load_cfg(bench.DEFAULT_WORKLOAD)
dev = torch.device("cuda")
lab, unl = bench.make_host_batch(1, 0, Bl, Bu, C, L)
x, y, uw = torch.from_numpy(lab["ecg"]).to(dev), torch.from_numpy(lab["target"]).to(dev), torch.from_numpy(unl["ecg"]).to(dev)
m1, m2 = init_model_from_cfg(cfg).to(dev), init_model_from_cfg(cfg).to(dev)
e1 = get_engine("cps", m1, m2, Bl, Bu, L, _lib.BF16, cfg["train"], external_pseudo=True)
e2 = get_engine("cps", m2, m1, Bl, Bu, L, _lib.BF16, cfg["train"], external_pseudo=True)
cps = CpsEngine(e1, e2)


def t(fn, n=100):
    for _ in range(5):
        fn()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(n):
        fn()
    b.record()
    torch.cuda.synchronize()
    return a.elapsed_time(b) / n


def full():
    cps.load_batch(x, y, uw)
    cps.step(1e-3)


print("full step ms", t(full), "side stream" if cps.side is not None else "serial")
print("pseudo A+B ms", t(lambda: (e1.pseudo(), e2.pseudo())))
print("train A only ms", t(lambda: e1.step(1e-3)))
