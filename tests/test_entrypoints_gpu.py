"""The reference's launch surface end to end on synthetic data: `algorithms.<name>.train(config)` (src/train.py:81-90)
for every algorithm on the engine, then `test(config)` and the inference script on the written checkpoint."""
import json
import os

import numpy as np
import pytest
import torch

from helpers import REPO

pytestmark = pytest.mark.gpu

CFG = os.path.join(REPO, "semi-seg-ecg_b200", "configs")


def _config(name, tmp_path):
    from utils.config import load_config
    cfg = load_config(os.path.join(CFG, "base", "resnet18", f"{name}.yaml"), os.path.join(CFG, "bench", "synthetic.yaml"),
                      {"output_dir": str(tmp_path), "exp_name": name})
    cfg["dataset"]["synthetic"] = {"length": 64, "valid_length": 24, "num_leads": 1}
    cfg["train"].update(epochs=2, warmup_epochs=1)
    cfg["dataloader"]["num_workers"] = 0
    return cfg


@pytest.mark.parametrize("name", ["scratch", "fixmatch", "mean_teacher", "cps"])
def test_train_then_test_and_inference(name, tmp_path):
    import algorithms
    cfg = _config(name, tmp_path)
    algo = algorithms.__dict__[cfg["algorithm"]]
    algo.train(cfg)
    out = os.path.join(str(tmp_path), name)
    rows = [json.loads(line) for line in open(os.path.join(out, "log.txt"))]
    assert [r["epoch"] for r in rows] == [0, 1]
    for r in rows:
        assert np.isfinite(r["valid_loss"]) and 0.0 <= r["MeanIoU"] <= 1.0
        assert np.isfinite(r["train_loss_total" if name != "scratch" else "train_loss"])
    assert os.path.exists(os.path.join(out, "best-loss.pth")) and os.path.exists(os.path.join(out, "best-MeanIoU.pth"))
    if name == "fixmatch":
        cfg["resume"] = os.path.join(out, "best-MeanIoU.pth")
        stats, metrics = algo.test(cfg)
        assert np.isfinite(stats["loss"]) and 0.0 <= metrics["MeanIoU"] <= 1.0
        import inference
        probs = inference.inference(cfg)
        assert probs.ndim == 3 and probs.shape[1] == 4 and probs.shape[2] == 2500
        assert np.allclose(probs.sum(axis=1), 1.0, atol=1e-5)
        assert np.array_equal(np.load(os.path.join(out, "test_outputs.npy")), probs)


def test_stpp_train_points_at_the_step(tmp_path):
    import algorithms
    with pytest.raises(NotImplementedError, match="train_one_epoch"):
        algorithms.stpp.train(_config("stpp", tmp_path))
