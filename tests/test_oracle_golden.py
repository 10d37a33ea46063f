"""Pins the CPU oracle (oracle/segnet_oracle.py) against outputs of the reference itself
(tests/golden/reference_vectors.npz, produced by tests/golden/make_golden.py).  The reference
has no tests / golden vectors of its own."""
import dataclasses
import os

import numpy as np
import pytest
import torch

from helpers import O, TINY_ARCH, TRAIN_CFG, batches, group, rel_err, sd_from

TOL = 2e-5  # fp32 reference vs fp64 oracle, relative L2 per tensor


def test_forward_backward_per_layer(golden):
    g = golden
    sd = sd_from(g, "A/init")
    tr = O.OracleTrainer(sd, TINY_ARCH, dict(TRAIN_CFG), dtype=torch.float64)
    x = torch.from_numpy(g["A/x"])
    y = torch.from_numpy(g["A/y"])
    st = tr.supervised_step(x, y, lr=0.0, want_taps=True)
    assert abs(st["loss"] - float(g["A/loss"])) < 1e-6
    acts, dacts = group(g, "A/act"), group(g, "A/dact")
    assert len(acts) == 22
    for name, ref in acts.items():
        assert rel_err(tr.taps[name], ref) < TOL, name
        assert rel_err(tr.taps[name].grad, dacts[name]) < TOL * 5, name
    for name, ref in group(g, "A/grad").items():
        assert rel_err(tr.grads[name], ref) < TOL * 5, name
    for name, ref in group(g, "A/after_train_fwd").items():
        assert rel_err(tr.sd[name], ref) < TOL, name
    # the reference's eval forward ran after the train forward: running stats updated once (lr=0: same weights)
    ev = O.forward(tr.sd, x.double(), TINY_ARCH, False)
    assert rel_err(ev["seg_logits"], g["A/seg_logits_eval"]) < TOL


def _run_fixmatch(g, tag, dropout):
    arch = dataclasses.replace(TINY_ARCH, dropout_ratio=dropout)
    cfg = dict(TRAIN_CFG, conf_thresh=float(g[f"{tag}/conf_thresh"]))
    tr = O.OracleTrainer(sd_from(g, f"{tag}/init"), arch, cfg, dtype=torch.float64)
    n, epoch = int(g[f"{tag}/nsteps"]), int(g[f"{tag}/epoch"])
    data = batches(int(g[f"{tag}/data_seed"]), n, 3, 3, 2, 300)
    masks = g[f"{tag}/dropout_masks"] if dropout > 0 else None
    stats = []
    for it, (lab, unl) in enumerate(data):
        lr = O.lr_at(it / n + epoch, cfg)
        dm = torch.from_numpy(masks[it]) if masks is not None else None
        stats.append(tr.fixmatch_step(lab["ecg"], lab["target"], unl["ecg"], unl["ecg_aug"], lr, dropout_mask=dm))
    return tr, stats


@pytest.mark.parametrize("tag,dropout", [("B", 0.0), ("D", 0.1)])
def test_fixmatch_steps(golden, tag, dropout):
    g = golden
    tr, stats = _run_fixmatch(g, tag, dropout)
    ref = group(g, f"{tag}/stats")
    for k in ("loss_total", "loss_x", "loss_u_s", "mask_ratio"):
        mean = float(np.mean([s[k] for s in stats]))
        assert abs(mean - float(ref[k])) < 2e-5 * max(1.0, abs(float(ref[k]))), (k, mean, float(ref[k]))
    assert 0.05 < float(ref["mask_ratio"]) < 0.95  # the masked branch is really exercised
    for name, refv in group(g, f"{tag}/final").items():
        if "num_batches_tracked" in name:
            assert int(tr.sd[name]) == int(refv)
        else:
            assert rel_err(tr.sd[name], refv) < 1e-4, name


@pytest.mark.parametrize("tag,dropout", [("B", 0.0), ("D", 0.1)])
def test_fixmatch_steps_fast_kernels(golden, tag, dropout, monkeypatch):
    """the timing arm's mode of the oracle (BatchNorm / max-pool / upsample / CE through the ATen kernels the reference's
    modules call, bench.py cpu_baseline / --impl reference) is held to the same golden vectors"""
    monkeypatch.setattr(O, "FAST_KERNELS", True)
    test_fixmatch_steps(golden, tag, dropout)


def test_fixmatch_steps_with_gradient_clipping(golden_clip):
    """case M: the reference's train_one_epoch with max_norm below the run's gradient norms (clip_grad_norm_ active at every
    step): the oracle's clipped update reproduces the final weights and the norms the reference's scaler returned; the
    unclipped oracle does not (Adam is nearly scale-invariant, so the difference is small but well above the tolerance)."""
    g = golden_clip
    n, epoch = int(g["M/nsteps"]), int(g["M/epoch"])
    data = batches(int(g["M/data_seed"]), n, 3, 3, 2, 300)
    drift = {}
    for clip in (True, False):
        cfg = dict(TRAIN_CFG, conf_thresh=float(g["M/conf_thresh"]), max_norm=float(g["M/max_norm"]) if clip else None)
        tr = O.OracleTrainer(sd_from(g, "M/init"), TINY_ARCH, cfg, dtype=torch.float64)
        norms = []
        for it, (lab, unl) in enumerate(data):
            tr.fixmatch_step(lab["ecg"], lab["target"], unl["ecg"], unl["ecg_aug"], O.lr_at(it / n + epoch, cfg))
            norms.append(tr.grad_norm())
        drift[clip] = max(rel_err(tr.sd[name], refv) for name, refv in group(g, "M/final").items()
                          if "num_batches_tracked" not in name)
        if clip:
            assert np.allclose(norms, g["M/grad_norms"], rtol=2e-4), (norms, g["M/grad_norms"])
            assert min(norms) > float(g["M/max_norm"])
    assert drift[True] < 1e-4, drift
    assert drift[False] > 10 * drift[True], drift


def test_mean_teacher_steps(golden):
    g = golden
    cfg = dict(TRAIN_CFG, ema_decay=0.99)
    tr = O.OracleTrainer(sd_from(g, "C/init"), TINY_ARCH, cfg, dtype=torch.float64)
    tr.init_teacher(sd_from(g, "C/teacher_init_buffers"))
    n, epoch = int(g["C/nsteps"]), int(g["C/epoch"])
    stats = []
    for it, (lab, unl) in enumerate(batches(int(g["C/data_seed"]), n, 3, 3, 2, 300)):
        stats.append(tr.mean_teacher_step(lab["ecg"], lab["target"], unl["ecg"], unl["ecg_aug"], O.lr_at(it / n + epoch, cfg)))
    ref = group(g, "C/stats")
    for k in ("loss_total", "loss_x", "loss_u_s"):
        assert abs(float(np.mean([s[k] for s in stats])) - float(ref[k])) < 2e-5
    for name, refv in group(g, "C/final").items():
        if "num_batches_tracked" not in name:
            assert rel_err(tr.sd[name], refv) < 1e-4, name
    teacher = tr.teacher_state()
    for name, refv in group(g, "C/teacher_final").items():
        if "num_batches_tracked" in name:
            # reference quirk: the int64 counter becomes a float32 EMA (0.01 -> 0.0299 -> 0.0596)
            assert refv.dtype == np.float32 and abs(float(teacher[name]) - float(refv)) < 1e-6, name
        else:
            assert rel_err(teacher[name], refv) < 1e-4, name


def _check_final(sd, g, prefix, tol=1e-4):
    for name, refv in group(g, prefix).items():
        if "num_batches_tracked" in name:
            assert int(sd[name]) == int(refv), name
        else:
            assert rel_err(sd[name], refv) < tol, name


def test_cps_steps(golden_semi):
    """oracle cps_step vs the reference's cps.train_one_epoch (case F): both models' final states and the logged means"""
    g = golden_semi
    tr_1 = O.OracleTrainer(sd_from(g, "F/init_1"), TINY_ARCH, TRAIN_CFG, dtype=torch.float64)
    tr_2 = O.OracleTrainer(sd_from(g, "F/init_2"), TINY_ARCH, TRAIN_CFG, dtype=torch.float64)
    n, epoch = int(g["F/nsteps"]), int(g["F/epoch"])
    stats = []
    for it, (lab, unl) in enumerate(batches(int(g["F/data_seed"]), n, 3, 3, 2, 300)):
        stats.append(O.cps_step(tr_1, tr_2, lab["ecg"], lab["target"], unl["ecg"], O.lr_at(it / n + epoch, TRAIN_CFG)))
    ref = group(g, "F/stats")
    for k in ("loss_total", "loss_x", "loss_u_s"):
        assert abs(float(np.mean([s[k] for s in stats])) - float(ref[k])) < 2e-5, k
    _check_final(tr_1.sd, g, "F/final_1")
    _check_final(tr_2.sd, g, "F/final_2")
    # the two models really differ (different init) and really moved
    assert rel_err(tr_1.sd["backbone.layer1.0.conv1.weight"], g["F/final_2/backbone.layer1.0.conv1.weight"]) > 0.1


def test_stpp_steps(golden_semi):
    """oracle stpp_step vs the reference's stpp.train_one_epoch (case G): frozen teacher, hard labels"""
    g = golden_semi
    tr = O.OracleTrainer(sd_from(g, "G/init"), TINY_ARCH, TRAIN_CFG, dtype=torch.float64)
    teacher = {k: (v.double() if v.is_floating_point() else v) for k, v in sd_from(g, "G/teacher").items()}
    n, epoch = int(g["G/nsteps"]), int(g["G/epoch"])
    stats = []
    for it, (lab, unl) in enumerate(batches(int(g["G/data_seed"]), n, 3, 3, 2, 300)):
        stats.append(tr.stpp_step(lab["ecg"], lab["target"], unl["ecg"], teacher, O.lr_at(it / n + epoch, TRAIN_CFG)))
    ref = group(g, "G/stats")
    for k in ("loss_total", "loss_x", "loss_u_s"):
        assert abs(float(np.mean([s[k] for s in stats])) - float(ref[k])) < 2e-5, k
    _check_final(tr.sd, g, "G/final")
    for name, refv in group(g, "G/teacher_final").items():      # the teacher is untouched by the step
        assert np.array_equal(g[f"G/teacher/{name}"], refv), name


def test_evaluate_oracle_matches_reference(golden_eval):
    """oracle/eval_oracle.evaluate vs the reference's base.evaluate (case H): loss, soft-max outputs, one-hot labels and
    -- with the restated torchmetrics MeanIoU on both sides -- the metric in its three configurations."""
    from helpers import eval_batches
    from oracle import eval_oracle
    g = golden_eval
    sd = sd_from(g, "H/model")
    bs = eval_batches(g)
    for tag, kw in (("H", {}), ("Hnb", {"include_background": False}), ("Hpc", {"per_class": True})):
        stats, metrics, outputs, preds = eval_oracle.evaluate(sd, TINY_ARCH, bs, **kw)
        assert abs(stats["loss"] - float(g[f"{tag}/stats/loss"])) < 1e-5
        ref = group(g, f"{tag}/metrics")
        assert set(metrics) == set(ref)
        for k in ref:
            assert abs(metrics[k] - float(ref[k])) < 1e-9, (tag, k, metrics[k], float(ref[k]))
    assert rel_err(outputs, g["H/outputs"]) < 1e-5
    assert np.array_equal(preds.numpy(), g["H/outputs"].argmax(1))
    onehot = torch.nn.functional.one_hot(torch.cat([b["target"] for b in bs]), 4).movedim(-1, 1).numpy()
    assert np.array_equal(onehot, g["H/labels_onehot"])
    # the aggregation really is a mean of per-batch means of per-sample IoUs: it differs from the pooled IoU
    conf = np.zeros((4, 4))
    for yt, yp in zip(torch.cat([b["target"] for b in bs]).numpy().ravel(), preds.numpy().ravel()):
        conf[yt, yp] += 1
    pooled = np.mean(np.diag(conf) / np.maximum(conf.sum(0) + conf.sum(1) - np.diag(conf), 1))
    assert abs(pooled - float(g["H/metrics/MeanIoU"])) > 1e-3


R34_ARCH = dataclasses.replace(TINY_ARCH, stage_blocks=(3, 4, 6, 3))


def test_resnet34_oracle_matches_reference(golden_arch):
    """the oracle's BasicBlock stack at depth 3-4-6-3 vs the reference's resnet34 (case R): train forward + loss +
    every gradient + running statistics, eval logits, two FixMatch steps"""
    g = golden_arch
    sd = {k: (v.double() if v.is_floating_point() else v) for k, v in sd_from(g, "R/init").items()}
    (lab, _), = batches(int(g["R/data_seed"]), 1, 4, 1, 2, 300)
    tr = O.OracleTrainer(sd, R34_ARCH, TRAIN_CFG, dtype=torch.float64)
    st = tr.supervised_step(lab["ecg"], lab["target"], 0.0)
    assert abs(st["loss"] - float(g["R/loss"])) < 1e-5
    for n, refv in group(g, "R/grad").items():
        assert rel_err(tr.grads[n], refv) < 2e-5, n
    for n, refv in group(g, "R/after_train_fwd").items():
        if "tracked" in n:
            assert int(tr.sd[n]) == int(refv)
        else:
            assert rel_err(tr.sd[n], refv) < 1e-5, n
    with torch.no_grad():     # (lr 0: the step above left the parameters alone and updated the running statistics once)
        ev = O.forward(tr.sd, lab["ecg"].double(), R34_ARCH, False)["seg_logits"]
    assert rel_err(ev, g["R/seg_logits_eval"]) < 1e-5
    cfg = dict(TRAIN_CFG, conf_thresh=float(g["R2/conf_thresh"]))
    tr = O.OracleTrainer(sd_from(g, "R2/init"), R34_ARCH, cfg, dtype=torch.float64)
    stats = []
    for it, (lab, unl) in enumerate(batches(int(g["R2/data_seed"]), 2, 3, 3, 2, 300)):
        stats.append(tr.fixmatch_step(lab["ecg"], lab["target"], unl["ecg"], unl["ecg_aug"], O.lr_at(it / 2 + 3, cfg)))
    ref = group(g, "R2/stats")
    for k in ("loss_total", "loss_x", "loss_u_s", "mask_ratio"):
        assert abs(float(np.mean([s[k] for s in stats])) - float(ref[k])) < 2e-5, k
    _check_final(tr.sd, g, "R2/final")


R50_ARCH = dataclasses.replace(TINY_ARCH, stage_blocks=(3, 4, 6, 3), bottleneck=True)


def test_resnet50_oracle_matches_reference(golden_arch):
    """the oracle's Bottleneck stack vs the reference's resnet50 (case K): parameter order, train forward + loss + every
    gradient + running statistics, eval logits"""
    g = golden_arch
    assert O.param_names(R50_ARCH) == [str(n) for n in g["K/param_order"]]
    sd = {k: (v.double() if v.is_floating_point() else v) for k, v in sd_from(g, "K/init").items()}
    assert set(O.param_names(R50_ARCH) + O.buffer_names(R50_ARCH)) == set(sd)
    (lab, _), = batches(int(g["K/data_seed"]), 1, 4, 1, 2, 300)
    tr = O.OracleTrainer(sd, R50_ARCH, TRAIN_CFG, dtype=torch.float64)
    st = tr.supervised_step(lab["ecg"], lab["target"], 0.0)
    assert abs(st["loss"] - float(g["K/loss"])) < 1e-5
    for n, refv in group(g, "K/grad").items():     # (the reference is fp32 through 53 layers: BN gradients are near-cancelling sums)
        assert rel_err(tr.grads[n], refv) < (2e-4 if ".bn" in n or "downsample.1" in n else 5e-5), n
    for n, refv in group(g, "K/after_train_fwd").items():
        if "tracked" in n:
            assert int(tr.sd[n]) == int(refv)
        else:
            assert rel_err(tr.sd[n], refv) < 1e-5, n
    with torch.no_grad():
        ev = O.forward(tr.sd, lab["ecg"].double(), R50_ARCH, False)["seg_logits"]
    assert rel_err(ev, g["K/seg_logits_eval"]) < 1e-5


def test_full_size_step_scalars(golden):
    """resnet18 @ 1x2500, one FixMatch step: losses, mask ratio and all 65 gradient norms."""
    import models.backbones  # product constructors give the seeded init (checked bit-exact in test_surface)
    from algorithms.base import init_model_from_cfg
    from helpers import model_cfg
    g = golden
    torch.manual_seed(0)
    model = init_model_from_cfg(model_cfg(1, 64, 64, 128, 0.0))
    arch = O.Arch(num_leads=1, dropout_ratio=0.0)
    cfg = dict(TRAIN_CFG, conf_thresh=float(g["E/conf_thresh"]))
    tr = O.OracleTrainer({k: v.detach() for k, v in model.state_dict().items()}, arch, cfg, dtype=torch.float32)
    (lab, unl), = batches(int(g["E/data_seed"]), 1, 2, 2, 1, 2500)
    s = tr.fixmatch_step(lab["ecg"], lab["target"], unl["ecg"], unl["ecg_aug"], O.lr_at(3.0, cfg))
    ref = group(g, "E/stats")
    for k in ("loss_total", "loss_x", "loss_u_s"):
        assert abs(s[k] - float(ref[k])) < 1e-4, (k, s[k], float(ref[k]))
    assert abs(s["mask_ratio"] - float(ref["mask_ratio"])) < 2e-3
    gn = np.array([float(tr.grads[n].double().norm()) for n in tr.pnames])
    assert np.allclose(gn, g["E/grad_norms"], rtol=2e-3, atol=1e-7)


def test_closed_form_loss_gradient():
    """The closed-form d(loss)/d(logits) used to check the fused CUDA loss kernel equals autograd."""
    torch.manual_seed(0)
    z = torch.randn(5, 4, 37, dtype=torch.float64, requires_grad=True)
    yx = torch.randint(0, 4, (2, 37))
    lab = torch.randint(0, 4, (3, 37))
    mask = torch.rand(3, 37) > 0.5
    loss = (O.ce_hard(z[:2], yx) + O.ce_masked(z[2:], lab, mask)) / 2
    loss.backward()
    assert rel_err(O.loss_grad_fullres(z.detach(), 2, yx, lab, mask, None), z.grad) < 1e-12
    z.grad = None
    p = torch.softmax(torch.randn(3, 4, 37, dtype=torch.float64), 1)
    loss = (O.ce_hard(z[:2], yx) + O.ce_soft(z[2:], p)) / 2
    loss.backward()
    assert rel_err(O.loss_grad_fullres(z.detach(), 2, yx, None, None, p), z.grad) < 1e-12


# ---- augmentation row: oracle pinned against the unmodified reference transforms ----
AUG_GOLDEN = os.path.join(os.path.dirname(__file__), "golden", "augment_vectors.npz")


def test_augment_oracle_matches_reference_transforms():
    """Seeding numpy like tests/golden/make_golden_aug.py did, the oracle must consume the same random
    stream and produce the reference's outputs (float32 items and int64 labels) exactly."""
    from oracle import augment_oracle as A
    g = np.load(AUG_GOLDEN)
    for seed, C, L in g["cases"]:
        pre = f"s{seed}"
        x, y = g[pre + "/x"], g[pre + "/y"]
        np.random.seed(int(seed))
        ecg, tgt, _ = A.labeled_item(x.copy(), y.copy(), int(L))
        assert np.array_equal(tgt, g[pre + "/lab_target"]), (seed, "labels")
        assert np.abs(ecg - g[pre + "/lab_ecg"]).max() <= 1e-6, (seed, "labeled ecg")
        np.random.seed(int(seed) + 1000)
        uw, us, _ = A.unlabeled_item(x.copy(), int(L))
        assert np.abs(uw - g[pre + "/unl_ecg"]).max() <= 1e-6, (seed, "weak view")
        assert np.abs(us - g[pre + "/unl_ecg_aug"]).max() <= 1e-6, (seed, "strong view")


def test_fourier_resample_is_scipy_resample():
    from scipy.signal import resample
    from oracle import augment_oracle as A
    rng = np.random.default_rng(0)
    for L, num in [(2500, 1250), (2500, 1251), (2500, 2500), (2500, 3708), (2500, 4999), (601, 433), (601, 900), (600, 600)]:
        x = rng.standard_normal((2, L))
        assert np.abs(A.fourier_resample(x, num) - resample(x, num, axis=1)).max() < 1e-10, (L, num)


def test_host_draws_follow_the_reference_order():
    """semiseg_b200.augment's draw functions (product side) consume numpy's stream exactly like the oracle's."""
    from oracle import augment_oracle as A
    from semiseg_b200 import augment as G
    cfg = G.AugConfig(target_length=600)
    np.random.seed(5)
    a = [A.draw_weak(600, 600), A.draw_strong(2, 600), A.draw_weak(600, 600)]
    np.random.seed(5)
    b = [G.draw_weak(600, cfg), G.draw_strong(2, 600, cfg, bulk=True), G.draw_weak(600, cfg)]
    assert (a[0]["size"], a[0]["start"]) == (b[0]["size"], b[0]["start"])
    assert (a[2]["size"], a[2]["start"]) == (b[2]["size"], b[2]["start"])
    for oa, ob in zip(a[1]["ops"], b[1]["ops"]):
        assert oa["op"] == ob["op"] and oa["apply"] == ob["apply"]
        if oa["apply"] and oa["op"] == "powerline":
            assert oa["freq"] == ob["a"]
        if oa["apply"] and oa["op"] in ("partial_white", "partial_sine"):
            assert (oa["count"], oa["start"]) == (ob["a"], ob["b"])
        if oa["apply"] and oa["op"] == "amplitude_scaling":
            assert np.array_equal(oa["scales"], ob["scales"])
