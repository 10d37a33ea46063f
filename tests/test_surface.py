"""Drop-in surface: configs parse through init_model_from_cfg, state_dict keys /
shapes / dtypes equal the reference's 128 entries, seeded init is bit-identical to the
reference's, LR schedule, optimizer facade, loud failure without CUDA."""
import glob
import os

import numpy as np
import pytest
import torch
import yaml

from helpers import O, REPO, TRAIN_CFG, group, model_cfg

CONFIGS = os.path.join(REPO, "semi-seg-ecg_b200", "configs")


def test_every_base_config_builds_a_model():
    from algorithms.base import init_model_from_cfg
    from utils.config import load_config
    files = sorted(glob.glob(os.path.join(CONFIGS, "base", "**", "*.yaml"), recursive=True))
    assert len(files) >= 10
    for f in files:
        cfg = load_config(f, os.path.join(CONFIGS, "bench", "ludb", "1over16.yaml"))
        assert cfg["exp_name"] == "ludb/1over16" and cfg["dataset"]["signal_length"] == 2500
        m = init_model_from_cfg(cfg)
        assert sum(p.numel() for p in m.parameters()) == 4041284
        assert len(m.state_dict()) == 128
    assert len(glob.glob(os.path.join(CONFIGS, "bench", "**", "*.yaml"), recursive=True)) >= 17


def test_algorithm_registry():
    import algorithms
    for name in ("base", "scratch", "fixmatch", "mean_teacher", "cps", "stpp"):
        mod = algorithms.__dict__[name]
        assert callable(mod.train) and callable(mod.test) and callable(mod.train_one_epoch)


def test_entry_point_signatures_follow_the_reference():
    """positional parameter names of the reference's entry points (base.py:83-93,184-191; fixmatch.py:28-39;
    mean_teacher.py:28-40; cps.py:28-41; stpp.py:91-103; inference.py:75)"""
    import inspect

    import algorithms
    import inference
    common = ["device", "epoch", "loss_scaler", "log_writer", "use_amp", "config"]
    expect = {
        "base": ["model", "data_loader", "optimizer"] + common,
        "fixmatch": ["model", "labeled_data_loader", "unlabeled_data_loader", "optimizer"] + common,
        "mean_teacher": ["model_student", "model_teacher", "labeled_data_loader", "unlabeled_data_loader", "optimizer"] + common,
        "cps": ["model_1", "model_2", "labeled_data_loader", "unlabeled_data_loader", "optimizer_1", "optimizer_2"] + common,
        "stpp": ["model_student", "model_teacher", "labeled_data_loader", "unlabeled_data_loader", "optimizer"] + common,
    }
    for name, params in expect.items():
        sig = inspect.signature(algorithms.__dict__[name].train_one_epoch)
        assert list(sig.parameters) == params, name
        assert sig.parameters["use_amp"].default is True and sig.parameters["log_writer"].default is None
    ev = inspect.signature(algorithms.base.evaluate)
    assert list(ev.parameters)[:5] == ["model", "data_loader", "device", "metric_fn", "use_amp"]
    assert list(inspect.signature(algorithms.stpp.select_reliable).parameters) == ["models", "dataloader", "device"]
    assert list(inspect.signature(algorithms.stpp.calculate_miou).parameters) == ["onehot_preds", "onehot_labels", "ignore_background"]
    assert list(inspect.signature(inference.inference).parameters) == ["config"]


def test_stpp_calculate_miou():
    """reference stpp.py:32-43: mean over classes of intersection / union on one-hot arrays, 0 for an empty union"""
    from algorithms.stpp import calculate_miou
    p = np.zeros((1, 3, 6), dtype=np.int64)
    t = np.zeros((1, 3, 6), dtype=np.int64)
    p[0, 0, :4] = 1; p[0, 1, 4:] = 1
    t[0, 0, :2] = 1; t[0, 1, 2:] = 1
    assert abs(calculate_miou(p, t) - (2 / 4 + 2 / 4 + 0.0) / 3) < 1e-12
    assert abs(calculate_miou(p, t, ignore_background=True) - (2 / 4 + 0.0) / 2) < 1e-12


def test_mean_iou_restatement_follows_torchmetrics_semantics():
    """oracle/eval_oracle.MeanIoU and semiseg_b200.evaluate.mean_iou_from_counts agree; per-sample IoU, empty union -> 0,
    batch mean accumulated per update() and divided by the number of batches (torchmetrics 1.5.2)"""
    from oracle.eval_oracle import MeanIoU
    from semiseg_b200.evaluate import mean_iou_from_counts
    rng = np.random.RandomState(0)
    m = MeanIoU(4)
    per_batch = []
    for n in (3, 1):
        pred, tgt = rng.randint(0, 3, (n, 50)), rng.randint(0, 3, (n, 50))     # class 3 never occurs
        oh = lambda a: torch.nn.functional.one_hot(torch.from_numpy(a), 4).movedim(-1, 1)   # noqa: E731
        m.update(oh(pred), oh(tgt))
        counts = torch.tensor([[[int(((pred[i] == c) & (tgt[i] == c)).sum()), int((pred[i] == c).sum()), int((tgt[i] == c).sum())]
                                for c in range(4)] for i in range(n)])
        per_batch.append(float(mean_iou_from_counts(counts).mean()))
    assert abs(float(m.compute()) - float(np.mean(per_batch))) < 1e-12
    assert float(m.compute()) < 0.75        # the absent class contributes a 0 to every sample's class mean


def test_mean_iou_known_answers_by_hand():
    """torchmetrics itself is not installable here, so the aggregation is held to answers worked out BY HAND from the
    published algorithm (the fractions below are the arithmetic, not output of any code): per sample and class
    intersection / union with 0 for an empty union, class mean per sample, batch mean per update(), mean of the batch
    means at compute() -- in the three configurations perf_metrics.py:14-20 can build."""
    from oracle.eval_oracle import MeanIoU
    from semiseg_b200.evaluate import aggregate_eval
    oh = lambda a: torch.nn.functional.one_hot(torch.tensor(a), 3).movedim(-1, 1)   # noqa: E731
    # batch 1, sample 0: pred 0 0 1 2 / target 0 1 1 2 -> class 0: 1/2, class 1: 1/2, class 2: 1/1
    #          sample 1: pred 1 1 1 1 / target 1 1 0 0 -> class 0: 0/2, class 1: 2/4, class 2: empty union -> 0
    # batch 2, sample 0: pred 2 2 0 0 / target 2 0 0 0 -> class 0: 2/3, class 1: empty -> 0, class 2: 1/2
    b1 = ([[0, 0, 1, 2], [1, 1, 1, 1]], [[0, 1, 1, 2], [1, 1, 0, 0]])
    b2 = ([[2, 2, 0, 0]], [[2, 0, 0, 0]])
    want = {
        (True, False): ((2 / 3 + 1 / 6) / 2 + (2 / 3 + 0 + 1 / 2) / 3) / 2,
        (False, False): ((3 / 4 + 1 / 4) / 2 + (0 + 1 / 2) / 2) / 2,
        (True, True): [((1 / 2 + 0) / 2 + 2 / 3) / 2, ((1 / 2 + 1 / 2) / 2 + 0) / 2, ((1 + 0) / 2 + 1 / 2) / 2],
    }
    for (bg, pc), w in want.items():
        m = MeanIoU(3, include_background=bg, per_class=pc)
        per_batch = []
        for p, t in (b1, b2):
            m.update(oh(p), oh(t))
            P, T = np.array(p), np.array(t)
            counts = torch.tensor([[[int(((P[i] == c) & (T[i] == c)).sum()), int((P[i] == c).sum()), int((T[i] == c).sum())]
                                    for c in range(3)] for i in range(len(p))], dtype=torch.int32)
            per_batch.append((torch.tensor([1.0, 1.0], dtype=torch.float64), counts, len(p)))
        got = m.compute()
        assert np.allclose(np.asarray(got, dtype=np.float64), np.asarray(w), atol=1e-12), (bg, pc, got, w)
        _, md = aggregate_eval(per_batch, include_background=bg, per_class=pc)      # the product's aggregation
        prod = [md[f"MeanIoU_{i}"] for i in range(3)] if pc else md["MeanIoU"]
        assert np.allclose(np.asarray(prod), np.asarray(w), atol=1e-12), (bg, pc, prod, w)


def test_state_dict_matches_reference_golden(golden):
    from algorithms.base import init_model_from_cfg
    ref = group(golden, "A/init")
    torch.manual_seed(0)
    m = init_model_from_cfg(model_cfg(2, 8, 8, 16, 0.0))
    sd = m.state_dict()
    assert list(sd.keys()) == list(ref.keys())
    for k, v in sd.items():
        assert tuple(v.shape) == tuple(ref[k].shape) and str(v.dtype).replace("torch.", "") == str(ref[k].dtype), k
        assert np.array_equal(v.numpy(), ref[k]), f"seeded init differs from the reference for {k}"
    # full-size model: per-tensor checksums of the reference init
    torch.manual_seed(0)
    m = init_model_from_cfg(model_cfg(1, 64, 64, 128, 0.0))
    cs = np.array([float(v.double().sum()) for v in m.state_dict().values()])
    assert np.array_equal(cs, golden["E/init_checksum"])


def test_resnet34_and_resnet50_init_match_reference(golden_arch):
    """seeded init of the deeper family members is bit-identical to the reference's (cases R, K), and the parameter
    arena follows the reference's registration order"""
    import dataclasses

    from algorithms.base import init_model_from_cfg
    from semiseg_b200.net import ParamLayout
    from semiseg_b200.runtime import spec_from_modules
    for name, tag, seed, head_in in (("resnet34", "R", 31, 64), ("resnet50", "K", 41, 256)):
        cfg = model_cfg(2, 8, 8, 16, 0.0)
        cfg["backbone"] = {name: cfg["backbone"]["resnet18"]}
        cfg["decode_head"]["FCNHead"]["in_channels"] = head_in
        torch.manual_seed(seed)
        m = init_model_from_cfg(cfg)
        ref = group(golden_arch, f"{tag}/init")
        sd = m.state_dict()
        assert list(sd.keys()) == list(ref.keys()), name
        for k, v in sd.items():
            assert np.array_equal(v.numpy(), ref[k]), f"{name}: seeded init differs from the reference for {k}"
        lay = ParamLayout(spec_from_modules(m.backbone, m.decode_head))
        assert lay.param_names() == [n for n, _ in m.named_parameters()]
        arch = dataclasses.replace(O.Arch(num_leads=2, stem_channels=8, base_channels=8, head_channels=16, dropout_ratio=0.0),
                                   stage_blocks=(3, 4, 6, 3), bottleneck=(name == "resnet50"))
        assert lay.param_names() == O.param_names(arch) and lay.buffer_names() == O.buffer_names(arch)


def test_layout_matches_oracle_names():
    from semiseg_b200.net import ParamLayout, SegNetSpec
    lay = ParamLayout(SegNetSpec())
    arch = O.Arch()
    assert lay.param_names() == O.param_names(arch)
    assert lay.buffer_names() == O.buffer_names(arch)
    assert lay.numel == 4041284 and len(lay.params) == 65 and len(lay.bns) == 21 and len(lay.convs) == 20
    for _, _, off in lay.params:
        assert off % 64 == 0


def test_plan_geometry():
    """Stage lengths and pitches: 2500 -> 1250 -> 625 -> 625 -> 313 -> 157 -> 79."""
    assert O.stage_lengths(O.Arch(), 2500) == [1250, 625, 625, 313, 157, 79]
    assert O.stage_lengths(O.Arch(), 5000) == [2500, 1250, 1250, 625, 313, 157]


def test_lr_schedule_matches_oracle():
    from utils.lr_sched import adjust_learning_rate
    opt = torch.optim.SGD([torch.nn.Parameter(torch.zeros(1))], lr=1.0)
    for e in (0.0, 0.5, 3.2, 10.0, 55.5, 99.9):
        lr = adjust_learning_rate(opt, e, TRAIN_CFG)
        assert lr == pytest.approx(O.lr_at(e, TRAIN_CFG), rel=0, abs=1e-15)
        assert opt.param_groups[0]["lr"] == lr
    assert adjust_learning_rate(opt, 0.0, TRAIN_CFG) == 0.0       # step 0 has lr = 0


def test_unsupported_variants_fail_loudly():
    import models.backbones as bb
    import models.decode_heads as dh
    from utils.optimizer import get_optimizer_from_config
    assert bb.resnet50(num_leads=1, stem_channels=8, base_channels=8).feat_dim == 8 * 8 * 4     # Bottleneck family is built
    with pytest.raises(NotImplementedError):
        bb.resnet18(num_leads=1, deep_stem=True)
    with pytest.raises(NotImplementedError):
        bb.vit_tiny()
    with pytest.raises(NotImplementedError):
        dh.FCNHead(512, 128, 4, num_convs=2, concat_input=False)
    with pytest.raises(NotImplementedError):
        get_optimizer_from_config(dict(TRAIN_CFG, optimizer="sgd"), [torch.nn.Parameter(torch.zeros(1))])


def test_cpu_forward_fails_loudly():
    from algorithms.base import init_model_from_cfg
    m = init_model_from_cfg(model_cfg(1, 8, 8, 16, 0.0))
    with pytest.raises(RuntimeError, match="no CPU fallback|CUDA"):
        m(torch.zeros(1, 1, 64))


def test_config_merge_and_cli_override(tmp_path):
    from utils.config import deep_merge, load_config
    a = {"x": {"y": 1, "z": 2}, "k": [1, 2]}
    b = {"x": {"y": 5}, "k": [3]}
    assert deep_merge(a, b) == {"x": {"y": 5, "z": 2}, "k": [3]} and a["x"]["y"] == 1
    p = tmp_path / "c.yaml"
    p.write_text(yaml.safe_dump({"exp_name": "a", "resume": None, "start_epoch": 0}))
    cfg = load_config(str(p), None, {"exp_name": "cli", "resume": "", "start_epoch": 0})
    assert cfg["exp_name"] == "cli" and cfg["resume"] is None      # only truthy CLI values overwrite


def test_synthetic_batches_follow_the_dataset_contract():
    from semiseg_b200 import synthetic
    lab, unl = synthetic.make_batch(0, 4, 6, 12, 5000, fs=500)
    assert lab["ecg"].shape == (4, 12, 5000) and lab["ecg"].dtype == np.float32
    assert lab["target"].shape == (4, 5000) and lab["target"].dtype == np.int64
    assert set(np.unique(lab["target"])) == {0, 1, 2, 3}
    assert unl["ecg"].shape == unl["ecg_aug"].shape == (6, 12, 5000)
    assert abs(float(lab["ecg"][0].mean())) < 1e-5 and abs(float(lab["ecg"][0].std()) - 1) < 1e-4
    lab2, _ = synthetic.make_batch(0, 4, 6, 12, 5000, fs=500)
    assert np.array_equal(lab["ecg"], lab2["ecg"])
