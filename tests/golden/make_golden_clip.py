"""Golden vectors for gradient clipping (TEST INFRASTRUCTURE; run in the build container only):

    python tests/golden/make_golden_clip.py

Three FixMatch steps of the UNMODIFIED reference (`algorithms.fixmatch.train_one_epoch`, CPU, fp32, use_amp=False) with
`max_norm` set below the gradient norm the run has, so `loss_scaler(..., clip_grad=max_norm)` (fixmatch.py:129-136 ->
misc.py:242-250 -> torch.nn.utils.clip_grad_norm_) really scales every update.  Stores inputs and outputs (numbers
only) in tests/golden/clip_vectors.npz; same harness as make_golden.py."""
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
import make_golden as G  # noqa: E402


def main():
    R = G.import_reference()
    torch.set_num_threads(4)
    torch.use_deterministic_algorithms(True)
    out, T = {}, G.TINY
    seed, nsteps = 7, 3
    torch.manual_seed(seed)
    model = R.base.init_model_from_cfg(G.model_cfg(T["num_leads"], T["stem_channels"], T["base_channels"], T["head_channels"], 0.0))
    G.put(out, "M/init", G.to_np(model.state_dict()))
    labl, unll = G.batches(200 + seed, nsteps, T["Bl"], T["Bu"], T["num_leads"], T["L"])
    model.eval()
    with torch.no_grad():
        conf = model(unll[0]["ecg"])["seg_logits"].softmax(1).max(1)[0]
    thresh = float(np.round(conf.median().item(), 3))
    max_norm = 0.25
    tc = G.train_cfg(conf_thresh=thresh, max_norm=max_norm)
    opt = R.optimizer.get_optimizer_from_config(tc, model.parameters())
    scaler = R.misc.NativeScalerWithGradNormCount()
    norms = []

    class Spy:                      # records the norm the reference's scaler returns (the norm BEFORE clipping)
        def __call__(self, *a, **k):
            n = scaler(*a, **k)
            norms.append(float(n))
            return n

        def __getattr__(self, k):
            return getattr(scaler, k)

    stats = R.fixmatch.train_one_epoch(model, labl, unll, opt, torch.device("cpu"), 3, Spy(), None, False, tc)
    assert len(norms) == nsteps and min(norms) > max_norm, norms     # clipping active at every step
    out["M/conf_thresh"], out["M/max_norm"] = np.float64(thresh), np.float64(max_norm)
    out["M/epoch"], out["M/nsteps"], out["M/data_seed"] = np.int64(3), np.int64(nsteps), np.int64(200 + seed)
    out["M/grad_norms"] = np.array(norms)
    G.put(out, "M/stats", {k: np.float64(v) for k, v in stats.items()})
    G.put(out, "M/final", G.to_np(model.state_dict()))
    path = os.path.join(HERE, "clip_vectors.npz")
    np.savez_compressed(path, **out)
    print("wrote", path, os.path.getsize(path), "bytes; grad norms", norms)


if __name__ == "__main__":
    main()
