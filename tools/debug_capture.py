"""Find which C-ABI call invalidates a stream capture (debug aid)."""
import os, sys
REPO = os.path.abspath(os.path.join(os.path.dirname(__file__), ".."))
sys.path.insert(0, REPO); sys.path.insert(0, os.path.join(REPO, "semi-seg-ecg_b200", "src")); sys.path.insert(0, os.path.join(REPO, "tests"))
import numpy as np, torch
from helpers import TRAIN_CFG, batches, model_cfg
from algorithms.base import init_model_from_cfg
from semiseg_b200 import _lib
from semiseg_b200.trainer import get_engine

torch.manual_seed(0)
model = init_model_from_cfg(model_cfg(2, 8, 8, 16, 0.1)).to("cuda")
cfg = dict(TRAIN_CFG, conf_thresh=0.3)
(lab, unl), = batches(11, 1, 3, 3, 2, 300)
eng = get_engine("fixmatch", model, None, 3, 3, 300, _lib.BF16, cfg, use_graph=True)
eng.load_batch(lab["ecg"], lab["target"], unl["ecg"], unl["ecg_aug"])
n = [0]
def hook(name, a):
    _lib.raw_call(name, *a)
    n[0] += 1
    try:
        ok = torch.cuda.is_current_stream_capturing()
    except Exception as e:
        print(f"capture invalidated after call #{n[0]} {name}: {e}")
        raise
_lib._hook = hook
try:
    eng.step(1e-3)
    print("captured OK, calls:", n[0])
    print(eng.read_stats())
except Exception as e:
    print("FAILED:", type(e).__name__, str(e)[:300])
