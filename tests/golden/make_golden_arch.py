"""Golden vectors for another member of the BasicBlock family, made by running the
UNMODIFIED reference (TEST INFRASTRUCTURE; build container only):

    python tests/golden/make_golden_arch.py

  case R: resnet34 (stage blocks 3-4-6-3, reference resnet.py:378-389) + FCNHead at tiny widths, two leads:
          train-mode forward + CE + backward (logits, loss, every parameter gradient, updated running statistics),
          eval-mode logits, and 2 steps of the reference's fixmatch.train_one_epoch.
  case K: resnet50 (Bottleneck blocks 3-4-6-3, expansion 4, resnet.py:75-132, 391-402) + FCNHead at tiny widths:
          train-mode forward + CE + backward, running statistics, eval logits, parameter registration order.
Stores numbers only (tests/golden/arch_vectors.npz)."""
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
REPO = os.path.abspath(os.path.join(HERE, "..", ".."))
sys.path.insert(0, REPO)
sys.path.insert(0, HERE)
from oracle.ref_harness import import_reference  # noqa: E402
import make_golden as mg  # noqa: E402


def cfg34(dropout=0.0):
    c = mg.model_cfg(2, 8, 8, 16, dropout)
    c["backbone"] = {"resnet34": c["backbone"]["resnet18"]}
    return c


def main():
    R = import_reference()
    torch.set_num_threads(4)
    torch.use_deterministic_algorithms(True)
    out = {}
    torch.manual_seed(int(os.environ.get("SEED_R", "31")))
    model = R.base.init_model_from_cfg(cfg34())
    assert sum(1 for k in model.state_dict() if k.endswith("conv1.weight")) == 16       # 3+4+6+3 BasicBlocks
    mg.put(out, "R/init", mg.to_np(model.state_dict()))
    lab, _ = mg.synthetic.make_batch(800, 4, 1, 2, 300)
    x, y = torch.from_numpy(lab["ecg"]), torch.from_numpy(lab["target"])
    out["R/data_seed"] = np.int64(800)
    model.train()
    res = model(x, y, return_loss=True)
    res["loss"].backward()
    out["R/seg_logits_train"] = res["seg_logits"].detach().numpy().copy()
    out["R/loss"] = np.float64(res["loss"].item())
    mg.put(out, "R/grad", {n: p.grad.detach().numpy().copy() for n, p in model.named_parameters()})
    mg.put(out, "R/after_train_fwd", mg.to_np({k: v for k, v in model.state_dict().items() if "running" in k or "tracked" in k}))
    model.eval()
    with torch.no_grad():
        out["R/seg_logits_eval"] = model(x)["seg_logits"].numpy()
    # two FixMatch steps
    torch.manual_seed(int(os.environ.get("SEED_R2", "32")))
    model = R.base.init_model_from_cfg(cfg34())
    mg.put(out, "R2/init", mg.to_np(model.state_dict()))
    labl, unll = mg.batches(810, 2, 3, 3, 2, 300)
    model.eval()
    with torch.no_grad():
        conf = model(unll[0]["ecg"])["seg_logits"].softmax(1).max(1)[0]
    # (this deep, narrow network saturates the soft-max at random init -- conf == 1.0f at most positions -- so the
    # median is not a usable threshold: take the first candidate that splits the positions and has no confidence
    # within 1e-6 of it, so that fp32 and fp64 restatements take the same decisions)
    thresh = None
    for cand in (float(np.round(conf.median().item(), 3)), 0.9999, 0.999, 0.99, 0.95, 0.9, 0.8):
        frac = float((conf >= cand).float().mean())
        if 0.1 < frac < 0.9 and float((conf - cand).abs().min()) > 1e-6:
            thresh = cand
            break
    assert thresh is not None
    print("threshold", thresh, "initial mask ratio", frac)
    tc = mg.train_cfg(conf_thresh=thresh)
    opt = R.optimizer.get_optimizer_from_config(tc, model.parameters())
    stats = R.fixmatch.train_one_epoch(model, labl, unll, opt, torch.device("cpu"), 3, R.misc.NativeScalerWithGradNormCount(),
                                       None, False, tc)
    out["R2/conf_thresh"], out["R2/data_seed"] = np.float64(thresh), np.int64(810)
    mg.put(out, "R2/stats", {k: np.float64(v) for k, v in stats.items()})
    mg.put(out, "R2/final", mg.to_np(model.state_dict()))
    # ---- case K: resnet50 (Bottleneck 3-4-6-3, reference resnet.py:75-132, 391-402) at tiny widths; head in = 8*8*4 ----
    def cfg50():
        c = mg.model_cfg(2, 8, 8, 16, 0.0)
        c["backbone"] = {"resnet50": c["backbone"]["resnet18"]}
        c["decode_head"]["FCNHead"]["in_channels"] = 8 * 8 * 4
        return c
    torch.manual_seed(int(os.environ.get("SEED_K", "41")))
    model = R.base.init_model_from_cfg(cfg50())
    assert sum(1 for k in model.state_dict() if k.endswith("conv3.weight")) == 16
    mg.put(out, "K/init", mg.to_np(model.state_dict()))
    lab, _ = mg.synthetic.make_batch(840, 4, 1, 2, 300)
    x, y = torch.from_numpy(lab["ecg"]), torch.from_numpy(lab["target"])
    out["K/data_seed"] = np.int64(840)
    model.train()
    res = model(x, y, return_loss=True)
    res["loss"].backward()
    out["K/seg_logits_train"] = res["seg_logits"].detach().numpy().copy()
    out["K/loss"] = np.float64(res["loss"].item())
    mg.put(out, "K/grad", {n: p.grad.detach().numpy().copy() for n, p in model.named_parameters()})
    mg.put(out, "K/after_train_fwd", mg.to_np({k: v for k, v in model.state_dict().items() if "running" in k or "tracked" in k}))
    model.eval()
    with torch.no_grad():
        out["K/seg_logits_eval"] = model(x)["seg_logits"].numpy()
    out["K/param_order"] = np.array([n for n, _ in model.named_parameters()])

    path = os.path.join(HERE, "arch_vectors.npz")
    np.savez_compressed(path, **out)
    print(f"wrote {path}: {len(out)} arrays, {os.path.getsize(path) / 1e6:.2f} MB")


if __name__ == "__main__":
    main()
