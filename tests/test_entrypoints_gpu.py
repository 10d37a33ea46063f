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
        # test(): <output_dir>/<exp_name>/best-<test.target_metric>.pth like the reference (base.py:455-468); a missing
        # checkpoint is an error, never a freshly initialised model scored
        cfg.setdefault("test", {})["target_metric"] = "no-such-metric"
        with pytest.raises(AssertionError, match="Checkpoint not found"):
            algo.test(cfg)
        cfg["test"]["target_metric"] = "MeanIoU"
        stats, metrics = algo.test(cfg)
        assert np.isfinite(stats["loss"]) and 0.0 <= metrics["MeanIoU"] <= 1.0
        row = open(os.path.join(out, "test_metrics.csv")).read().splitlines()
        assert row[0].split(",")[-1] == "loss" and len(row) == 2
        rows2 = [json.loads(line) for line in open(os.path.join(out, "log.txt"))]
        best = max(r["MeanIoU"] for r in rows2)
        assert abs(metrics["MeanIoU"] - best) < 0.5      # the trained checkpoint, evaluated on another split
        import inference
        probs = inference.inference(cfg)
        assert probs.ndim == 3 and probs.shape[1] == 4 and probs.shape[2] == 2500
        assert np.allclose(probs.sum(axis=1), 1.0, atol=1e-5)
        assert np.array_equal(np.load(os.path.join(out, "test_outputs.npy")), probs)


def test_stpp_train_points_at_the_step(tmp_path):
    import algorithms
    with pytest.raises(NotImplementedError, match="train_one_epoch"):
        algorithms.stpp.train(_config("stpp", tmp_path))


def test_mean_teacher_resume_keeps_the_teacher(tmp_path):
    """Resuming a Mean-Teacher run: the teacher restored from `model_ema` keeps its EMA history -- the first step after
    the resume continues the average, it does not re-initialise the teacher from the student."""
    import algorithms
    from algorithms.base import build_model_and_optimizer
    from algorithms.mean_teacher import init_teacher, train_one_epoch
    from helpers import batches
    from utils import misc
    cfg = _config("mean_teacher", tmp_path)
    cfg["train"]["ema_decay"] = 0.9
    algorithms.mean_teacher.train(cfg)
    path = os.path.join(str(tmp_path), "mean_teacher", "best-loss.pth")
    ck = torch.load(path, map_location="cpu", weights_only=False)
    assert "model_ema" in ck
    dev = torch.device("cuda")
    model, opt, scaler = build_model_and_optimizer(cfg, dev, cfg["seed"])
    teacher = init_teacher(cfg, model, dev)
    cfg["resume"] = path
    misc.load_model(cfg, model, opt, scaler, teacher)
    before = {k: v.detach().clone() for k, v in teacher.state_dict().items()}
    for k, v in ck["model_ema"].items():
        assert torch.allclose(before[k].cpu().float(), v.float(), atol=0, rtol=0), k
    assert any(float(v) != int(float(v)) for k, v in before.items() if "tracked" in k), "float EMA counters survive the resume"
    (lab, unl), = batches(5, 1, 16, 16, 1, 2500)
    train_one_epoch(model, teacher, [lab], [unl], opt, dev, 1, scaler, None, True, cfg["train"])
    d = 0.9
    student, after = model.state_dict(), teacher.state_dict()
    for k in before:
        want = before[k].float() * d + student[k].float() * (1.0 - d)
        assert torch.allclose(after[k].float(), want, rtol=1e-5, atol=1e-6), k
