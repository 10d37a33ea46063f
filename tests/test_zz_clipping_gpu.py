"""Gradient clipping (`train.max_norm`) through the fused step, against a clipped run of the reference itself (golden case M,
tests/golden/make_golden_clip.py).  In its own file, collected after the other GPU suites."""
import pytest
import torch

from helpers import O, TRAIN_CFG, batches, group, model_cfg, rel_err, sd_from

pytestmark = pytest.mark.gpu

from semiseg_b200 import _lib  # noqa: E402
from semiseg_b200.trainer import get_engine  # noqa: E402

DEV = "cuda"


def build(cfg, sd=None, seed=0):
    from algorithms.base import init_model_from_cfg
    torch.manual_seed(seed)
    m = init_model_from_cfg(cfg)
    if sd is not None:
        m.load_state_dict(sd)
    return m.to(DEV)


def tiny_cfg(dropout=0.0):
    return model_cfg(2, 8, 8, 16, dropout)


@pytest.mark.parametrize("use_graph", [False, True])
def test_fixmatch_steps_with_gradient_clipping_golden(golden_clip, use_graph):
    """max_norm (loss_scaler(..., clip_grad=max_norm), fixmatch.py:129-136; misc.py:242-250): through the plugin call
    against the reference's own run with clipping active at every step (case M); the engine's global norm is the norm
    the reference's scaler returned, and without max_norm the same run ends measurably elsewhere."""
    from algorithms.fixmatch import train_one_epoch
    from utils.optimizer import get_optimizer_from_config
    g = golden_clip
    n, epoch = int(g["M/nsteps"]), int(g["M/epoch"])
    data = batches(int(g["M/data_seed"]), n, 3, 3, 2, 300)
    drift = {}
    for clip in (True, False):
        cfg = dict(TRAIN_CFG, conf_thresh=float(g["M/conf_thresh"]), max_norm=float(g["M/max_norm"]) if clip else None)
        model = build(tiny_cfg(0.0), sd_from(g, "M/init"))
        if clip and use_graph:      # the drop-in call
            opt = get_optimizer_from_config(cfg, model.parameters())
            stats = train_one_epoch(model, [d[0] for d in data], [d[1] for d in data], opt, torch.device(DEV), epoch,
                                    None, None, False, cfg)
            for k in ("loss_total", "loss_x", "loss_u_s", "mask_ratio"):
                assert abs(stats[k] - float(g[f"M/stats/{k}"])) < 5e-5 * max(1.0, abs(float(g[f"M/stats/{k}"]))), k
        else:
            eng = get_engine("fixmatch", model, None, 3, 3, 300, _lib.F32, cfg, use_graph=use_graph)
            for it, (lab, unl) in enumerate(data):
                eng.load_batch(lab["ecg"], lab["target"], unl["ecg"], unl["ecg_aug"])
                eng.step(O.lr_at(it / n + epoch, cfg))
                if clip:
                    assert abs(float(eng.gnorm) - float(g["M/grad_norms"][it])) < 2e-4 * float(g["M/grad_norms"][it])
            eng.read_stats()
        sd = model.state_dict()
        drift[clip] = max(rel_err(sd[name], refv) for name, refv in group(g, "M/final").items() if "tracked" not in name)
    assert drift[True] < 2e-4, drift
    assert drift[False] > 5 * drift[True], drift
