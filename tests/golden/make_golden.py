"""Generate golden vectors by running the UNMODIFIED reference (TEST INFRASTRUCTURE).

Run in the build container only (needs /root/reference or $SEMISEG_REF):

    python tests/golden/make_golden.py

Imports the reference's own modules through oracle/ref_harness.py (three out-of-tree shims,
SURVEY.md Appendix B), drives `models.*`, `algorithms.fixmatch.train_one_epoch` and
`algorithms.mean_teacher.train_one_epoch` on seeded synthetic batches (CPU, fp32, use_amp=False)
and stores INPUTS and OUTPUTS (numbers only -- no reference source) in tests/golden/*.npz.
The reference ships no tests or fixtures of its own, so these files are
what pins the oracle (tests/test_oracle_golden.py) and, through it, the CUDA path.
"""
import copy
import importlib.util
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
REPO = os.path.abspath(os.path.join(HERE, "..", ".."))
sys.path.insert(0, REPO)
from oracle.ref_harness import ListLoader, import_reference  # noqa: E402

spec = importlib.util.spec_from_file_location(
    "ssb_synthetic", os.path.join(REPO, "semi-seg-ecg_b200", "src", "semiseg_b200", "synthetic.py"))
synthetic = importlib.util.module_from_spec(spec)
spec.loader.exec_module(synthetic)

TINY = dict(num_leads=2, stem_channels=8, base_channels=8, head_channels=16, L=300, Bl=3, Bu=3)


def model_cfg(num_leads, stem, base, head_ch, dropout):
    return {
        "backbone": {"resnet18": {"num_leads": num_leads, "num_stages": 4, "out_indices": [0, 1, 2, 3],
                                  "dilations": [1, 1, 1, 1], "strides": [1, 2, 2, 2], "deep_stem": False,
                                  "avg_down": False, "contract_dilation": False, "stem_channels": stem,
                                  "base_channels": base}},
        "decode_head": {"FCNHead": {"in_channels": base * 8, "in_index": 3, "channels": head_ch, "num_convs": 1,
                                    "concat_input": False, "dropout_ratio": dropout, "num_classes": 4,
                                    "align_corners": False}},
    }


def train_cfg(**over):
    c = {"epochs": 100, "accum_iter": 1, "warmup_epochs": 10, "min_lr": 0.0001, "blr": None, "lr": 0.001,
         "weight_decay": 0.05, "max_norm": None, "layer_decay": None, "optimizer": "adamw",
         "optimizer_kwargs": {"betas": [0.9, 0.999]}}
    c.update(over)
    return c


def to_np(sd):
    return {k: v.detach().cpu().numpy().copy() for k, v in sd.items()}


def put(out, prefix, d):
    for k, v in d.items():
        out[f"{prefix}/{k}"] = np.array(v, copy=True)


def batches(seed0, n, Bl, Bu, C, L):
    lab, unl = [], []
    for i in range(n):
        a, b = synthetic.make_batch(seed0 + i, Bl, Bu, C, L)
        lab.append({k: torch.from_numpy(v) for k, v in a.items()})
        unl.append({k: torch.from_numpy(v) for k, v in b.items()})
    return ListLoader(lab), ListLoader(unl)


def main():
    R = import_reference()
    torch.set_num_threads(4)
    torch.use_deterministic_algorithms(True)
    out = {}
    T = TINY

    # ---- case A: forward / backward of the tiny model, per-layer activations + gradients ----
    torch.manual_seed(0)
    cfg = model_cfg(T["num_leads"], T["stem_channels"], T["base_channels"], T["head_channels"], 0.0)
    model = R.base.init_model_from_cfg(cfg)
    put(out, "A/init", to_np(model.state_dict()))
    lab, unl = synthetic.make_batch(100, T["Bl"] + T["Bu"], 1, T["num_leads"], T["L"])
    x = torch.from_numpy(lab["ecg"])
    y = torch.from_numpy(lab["target"])
    out["A/x"], out["A/y"] = lab["ecg"], lab["target"]
    acts, grads = {}, {}
    hooks = []
    for name, mod in model.named_modules():
        if isinstance(mod, torch.nn.Conv1d):
            def fh(m, i, o, name=name):
                acts[name] = o.detach().clone()
                o.register_hook(lambda g, name=name: grads.__setitem__(name, g.detach().clone()))
            hooks.append(mod.register_forward_hook(fh))
    model.train()
    res = model(x, y, return_loss=True)
    res["loss"].backward()
    for h in hooks:
        h.remove()
    out["A/seg_logits_train"] = res["seg_logits"].detach().numpy().copy()
    out["A/loss"] = np.float64(res["loss"].item())
    put(out, "A/act", to_np(acts))
    put(out, "A/dact", to_np(grads))
    put(out, "A/grad", {n: p.grad.detach().numpy().copy() for n, p in model.named_parameters()})
    put(out, "A/after_train_fwd", to_np({k: v for k, v in model.state_dict().items() if "running" in k or "tracked" in k}))
    model.eval()
    with torch.no_grad():
        out["A/seg_logits_eval"] = model(x)["seg_logits"].numpy()

    # ---- case B: 3 FixMatch steps through the reference's train_one_epoch ----
    def run_fixmatch(tag, dropout, seed, nsteps, thresh=None, capture_masks=False):
        torch.manual_seed(seed)
        cfg = model_cfg(T["num_leads"], T["stem_channels"], T["base_channels"], T["head_channels"], dropout)
        model = R.base.init_model_from_cfg(cfg)
        put(out, f"{tag}/init", to_np(model.state_dict()))
        labl, unll = batches(200 + seed, nsteps, T["Bl"], T["Bu"], T["num_leads"], T["L"])
        if thresh is None:  # choose a threshold that splits the positions roughly in half at init
            model.eval()
            with torch.no_grad():
                conf = model(unll[0]["ecg"])["seg_logits"].softmax(1).max(1)[0]
            thresh = float(np.round(conf.median().item(), 3))
        tc = train_cfg(conf_thresh=thresh)
        masks = []
        if capture_masks:
            def dh(m, i, o):
                if m.training:
                    masks.append((o != 0).to(torch.uint8) | (i[0] == 0).to(torch.uint8))
            model.decode_head.dropout.register_forward_hook(dh)
        opt = R.optimizer.get_optimizer_from_config(tc, model.parameters())
        scaler = R.misc.NativeScalerWithGradNormCount()
        stats = R.fixmatch.train_one_epoch(model, labl, unll, opt, torch.device("cpu"), 3, scaler, None, False, tc)
        out[f"{tag}/conf_thresh"] = np.float64(thresh)
        out[f"{tag}/epoch"] = np.int64(3)
        out[f"{tag}/nsteps"] = np.int64(nsteps)
        out[f"{tag}/data_seed"] = np.int64(200 + seed)
        put(out, f"{tag}/stats", {k: np.float64(v) for k, v in stats.items()})
        put(out, f"{tag}/final", to_np(model.state_dict()))
        if capture_masks:
            # only the student (train-mode) forwards apply dropout: one mask per step
            assert len(masks) == nsteps, len(masks)
            out[f"{tag}/dropout_masks"] = np.stack([m.numpy() for m in masks])
        return model

    run_fixmatch("B", 0.0, 1, 3)
    run_fixmatch("D", 0.1, 2, 2, capture_masks=True)

    # ---- case C: 3 Mean-Teacher steps (teacher aliasing + EMA over buffers) ----
    torch.manual_seed(3)
    cfg = model_cfg(T["num_leads"], T["stem_channels"], T["base_channels"], T["head_channels"], 0.0)
    student = R.base.init_model_from_cfg(cfg)
    teacher = R.base.init_model_from_cfg(cfg)
    put(out, "C/init", to_np(student.state_dict()))
    put(out, "C/teacher_init_buffers", to_np({k: v for k, v in teacher.state_dict().items() if "running" in k or "tracked" in k}))
    for p in teacher.parameters():
        p.requires_grad = False
    with torch.no_grad():
        for q, k in zip(student.parameters(), teacher.parameters()):
            k.data = q.data   # what mean_teacher.train does (mean_teacher.py:285-290)
    labl, unll = batches(300, 3, T["Bl"], T["Bu"], T["num_leads"], T["L"])
    tc = train_cfg(ema_decay=0.99)
    opt = R.optimizer.get_optimizer_from_config(tc, student.parameters())
    scaler = R.misc.NativeScalerWithGradNormCount()
    stats = R.mean_teacher.train_one_epoch(student, teacher, labl, unll, opt, torch.device("cpu"), 3, scaler, None, False, tc)
    put(out, "C/stats", {k: np.float64(v) for k, v in stats.items()})
    put(out, "C/final", to_np(student.state_dict()))
    put(out, "C/teacher_final", to_np(teacher.state_dict()))
    out["C/epoch"], out["C/nsteps"], out["C/data_seed"] = np.int64(3), np.int64(3), np.int64(300)

    # ---- case E: full-size resnet18 (1 x 2500), one FixMatch step, scalars + per-tensor norms only ----
    torch.manual_seed(0)
    cfg = model_cfg(1, 64, 64, 128, 0.0)
    model = R.base.init_model_from_cfg(cfg)
    out["E/init_checksum"] = np.array([float(v.double().sum()) for v in model.state_dict().values()])
    labl, unll = batches(400, 1, 2, 2, 1, 2500)
    model.eval()
    with torch.no_grad():
        conf = model(unll[0]["ecg"])["seg_logits"].softmax(1).max(1)[0]
    thresh = float(np.round(conf.median().item(), 3))
    tc = train_cfg(conf_thresh=thresh)
    opt = R.optimizer.get_optimizer_from_config(tc, model.parameters())
    scaler = R.misc.NativeScalerWithGradNormCount()
    gn = {}
    orig_step = opt.step

    def step_spy(*a, **k):
        for n, p in model.named_parameters():
            gn[n] = float(p.grad.double().norm())
        return orig_step(*a, **k)
    opt.step = step_spy
    stats = R.fixmatch.train_one_epoch(model, labl, unll, opt, torch.device("cpu"), 3, scaler, None, False, tc)
    out["E/conf_thresh"] = np.float64(thresh)
    put(out, "E/stats", {k: np.float64(v) for k, v in stats.items()})
    out["E/grad_norms"] = np.array([gn[n] for n, _ in model.named_parameters()])
    out["E/final_norms"] = np.array([float(v.double().norm()) for v in model.state_dict().values()])
    out["E/data_seed"] = np.int64(400)

    path = os.path.join(HERE, "reference_vectors.npz")
    np.savez_compressed(path, **out)
    print(f"wrote {path}: {len(out)} arrays, {os.path.getsize(path) / 1e6:.2f} MB")


if __name__ == "__main__":
    main()
