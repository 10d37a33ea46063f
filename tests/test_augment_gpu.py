"""Augmentation kernels vs the oracle with INJECTED draws: the oracle makes the draws
in the reference's order (incl. the bulk noise arrays), the CUDA kernels are fed the very same draws.

Tolerances: the reference computes in float64 and casts the standardised item to float32; the kernels
compute in float32 (two dense 2500-term DFT passes) -> 2e-4 absolute on the unit-variance outputs;
resized labels (int64) are bit-exact.
"""
import os

import numpy as np
import pytest
import torch

from helpers import rel_err

pytestmark = pytest.mark.gpu

from oracle import augment_oracle as A  # noqa: E402
from semiseg_b200 import _lib, augment as G  # noqa: E402

DEV = "cuda"
GOLD = os.path.join(os.path.dirname(__file__), "golden", "augment_vectors.npz")
TOL = 2e-4


def strips(seed, B, C, L):
    g = np.load(GOLD)
    rng = np.random.default_rng(seed)
    t = np.arange(L) / 250.0
    xs, ys = [], []
    for b in range(B):
        x = np.zeros((C, L))
        for c in range(C):
            x[c] = (0.6 * np.sin(2 * np.pi * (0.9 + 0.2 * rng.random()) * t + rng.uniform(0, 6.28))
                    + 0.3 * np.sin(2 * np.pi * (5 + 5 * rng.random()) * t) + 0.15 * rng.standard_normal(L) + 0.2 * c)
        y = np.zeros((1, L), dtype=np.int64)
        pos = 0
        while pos < L:
            seg = int(rng.integers(5, 70))
            y[0, pos:pos + seg] = int(rng.integers(0, 4))
            pos += seg
        xs.append(x)
        ys.append(y)
    del g
    return xs, ys


@pytest.mark.parametrize("fft", ["0", "1"])
@pytest.mark.parametrize("C,L,B", [(1, 2500, 6), (2, 1000, 4), (3, 601, 3), (12, 500, 2), (1, 128, 5), (12, 5000, 2)])
def test_weak_resize_crop(C, L, B, fft, monkeypatch):
    """both evaluations of the Fourier resize: the dense DFT kernels (fft = 0) and the Bluestein / FFT kernels (fft = 1;
    the library picks by problem size when SSB_AUG_FFT is unset)"""
    monkeypatch.setenv("SSB_AUG_FFT", fft)
    np.random.seed(C * 1000 + L)
    xs, ys = strips(C + L, B, C, L)
    cfg = G.AugConfig(target_length=L)
    draws = [A.draw_weak(L, L) for _ in range(B)]
    draws[0] = {"size": L, "start": 0, "ratio": 1.0}                      # identity resize
    if B > 2:
        draws[1] = {"size": L // 2, "start": 0, "ratio": 0.5}             # smallest size, all padding around it
        draws[2] = {"size": 2 * L - 1, "start": L - 1, "ratio": 2.0}      # largest size, last crop window
    ref_x, ref_y = zip(*[A.weak_resize_crop(x, y.astype(np.float64), d, L) for x, y, d in zip(xs, ys, draws)])
    aug = G.GpuAugmenter(cfg, B, C, L, DEV)
    x_d = torch.from_numpy(np.stack(xs).astype(np.float32)).to(DEV)
    y_d = torch.from_numpy(np.stack(ys)[:, 0]).to(DEV)
    y_out = torch.full((B, L), -7, dtype=torch.int64, device=DEV)
    sz, st_ = G.weak_table(draws)
    out = aug.weak_resize_crop(x_d, y_d, sz, st_, labels_out=y_out)
    torch.cuda.synchronize()
    got = out.cpu().numpy().astype(np.float64)
    for b in range(B):
        scale = np.abs(ref_x[b]).max() + 1e-12
        assert np.abs(got[b] - ref_x[b]).max() / scale < TOL, (b, draws[b])
        assert np.array_equal(y_out[b].cpu().numpy(), ref_y[b][0].astype(np.int64)), (b, draws[b])
    # unlabeled call: no label pointers
    out2 = aug.weak_resize_crop(x_d, None, sz, st_, out=torch.empty_like(out))
    assert torch.equal(out2, out)
    with pytest.raises(ValueError):
        aug.weak_resize_crop(x_d, None, sz * 0 + 3 * L, st_)


@pytest.mark.parametrize("C,L,B", [(1, 2500, 8), (2, 1000, 6), (12, 500, 4)])
def test_strong_and_standardize(C, L, B):
    np.random.seed(77 + C)
    xs, _ = strips(3 * C + L, B, C, L)
    cfg = G.AugConfig(target_length=L)
    draws = [A.draw_strong(C, L) for _ in range(B)]
    # make sure every op is exercised at least once with apply = True
    forced = [{"ops": [{"op": "amplitude_scaling", "apply": True, "scales": np.random.normal(1, 0.5, size=(C, L))},
                       {"op": "powerline", "apply": True, "freq": 60},
                       {"op": "partial_white", "apply": True, "noise": np.random.randn(C, L), "count": L // 3, "start": 7}]},
              {"ops": [{"op": "partial_sine", "apply": True, "count": L // 2 - 1, "start": L // 2},
                       {"op": "powerline", "apply": True, "freq": 50},
                       {"op": "amplitude_scaling", "apply": False}]}]
    draws[:2] = forced
    ref_s = [A.standardize(A.strong_augment(x, d, fs=cfg.fs)) for x, d in zip(xs, draws)]
    ref_w = [A.standardize(x) for x in xs]
    scales = np.ones((B, C, L), dtype=np.float32)
    white = np.zeros((B, C, L), dtype=np.float32)
    gd = []
    for b, d in enumerate(draws):
        ops = []
        for op in d["ops"]:
            o = {"op": op["op"], "apply": op["apply"], "a": 0, "b": 0}
            if op["apply"]:
                if op["op"] == "amplitude_scaling":
                    scales[b] = op["scales"]
                elif op["op"] == "powerline":
                    o["a"] = op["freq"]
                else:
                    o["a"], o["b"] = op["count"], op["start"]
                    if op["op"] == "partial_white":
                        white[b] = op["noise"]
            ops.append(o)
        gd.append({"ops": ops})
    aug = G.GpuAugmenter(cfg, B, C, L, DEV)
    x_d = torch.from_numpy(np.stack(xs).astype(np.float32)).to(DEV)
    out_s, out_w = torch.empty_like(x_d), torch.empty_like(x_d)
    gd = G.ops_table(gd)
    aug.strong_standardize(x_d, out_s, gd, scales=torch.from_numpy(scales).to(DEV), white=torch.from_numpy(white).to(DEV))
    aug.strong_standardize(x_d, out_w)
    torch.cuda.synchronize()
    for b in range(B):
        assert np.abs(out_w[b].cpu().numpy() - ref_w[b]).max() < TOL, ("weak view", b)
        assert np.abs(out_s[b].cpu().numpy() - ref_s[b]).max() < TOL * 2, ("strong view", b, [(o["op"], o["apply"]) for o in draws[b]["ops"]])
    # device RNG mode: same ops without injected arrays -> finite, standardised, and different from the weak view
    out_r = torch.empty_like(x_d)
    aug.strong_standardize(x_d, out_r, gd)
    torch.cuda.synchronize()
    assert torch.isfinite(out_r).all()
    m = out_r.view(B, -1).mean(1).abs().max()
    s = (out_r.view(B, -1).std(1, unbiased=False) - 1).abs().max()
    assert float(m) < 1e-4 and float(s) < 1e-3
    # statistics of the device RNG: AmplitudeScaling factors ~ N(1, 0.5)
    ones = torch.ones(B, C, L, device=DEV)
    only_amp = G.ops_table([{"ops": [{"op": "amplitude_scaling", "apply": True, "a": 0, "b": 0}]} for _ in range(B)])
    raw = torch.empty_like(ones)
    aug.strong_standardize(ones, raw, only_amp)          # standardised N(1, .5) draws -> unit normal
    torch.cuda.synchronize()
    z = raw.flatten().double()
    assert abs(float((z ** 3).mean())) < 0.1 and abs(float((z ** 4).mean()) - 3.0) < 0.3


def test_fixmatch_batcher_fills_engine_arena():
    from algorithms.base import init_model_from_cfg
    from helpers import model_cfg
    from semiseg_b200.trainer import get_engine
    torch.manual_seed(0)
    np.random.seed(3)
    Bl = Bu = 3
    C, L = 2, 600
    model = init_model_from_cfg(model_cfg(C, 8, 8, 16, 0.0)).to(DEV)
    tcfg = {"epochs": 100, "warmup_epochs": 10, "min_lr": 1e-4, "lr": 1e-3, "weight_decay": 0.05, "optimizer": "adamw",
            "optimizer_kwargs": {"betas": [0.9, 0.999]}, "conf_thresh": 0.3}
    eng = get_engine("fixmatch", model, None, Bl, Bu, L, _lib.F32, tcfg, use_graph=False)
    cfg = G.AugConfig(target_length=L)
    xs, ys = strips(5, Bl + Bu, C, L)
    raw_l = torch.from_numpy(np.stack(xs[:Bl]).astype(np.float32)).to(DEV)
    lab_l = torch.from_numpy(np.stack(ys[:Bl])[:, 0]).to(DEV)
    raw_u = torch.from_numpy(np.stack(xs[Bl:]).astype(np.float32)).to(DEV)
    bat = G.FixMatchBatcher(eng, cfg, seed=1, exact_stream=True)
    bat.load(raw_l, lab_l, raw_u)
    torch.cuda.synchronize()
    # replay the same numpy stream through the oracle (scalar draws only: bulk arrays are device-side)
    np.random.seed(3)
    for b in range(Bl):
        d = A.draw_weak(L, L)
        xw, yw = A.weak_resize_crop(xs[b], ys[b].astype(np.float64), d, L)
        assert np.abs(eng.x_s[b].cpu().numpy() - A.standardize(xw)).max() < TOL
        assert np.array_equal(eng.y_l[b].cpu().numpy(), yw[0].astype(np.int64))
    for i in range(Bu):
        d = A.draw_weak(L, L)
        G.draw_strong(C, L, cfg)       # advances the stream like the batcher did
        xw = A.weak_resize_crop(xs[Bl + i], None, d, L)
        assert np.abs(eng.x_uw[i].cpu().numpy() - A.standardize(xw)).max() < TOL
    assert torch.isfinite(eng.x_s).all()
    eng.step(1e-3)
    s, = eng.read_stats()
    assert np.isfinite(s["loss_total"])
    # production mode: vectorised draws -- every draw inside its range, outputs standardised
    fast = G.FixMatchBatcher(eng, cfg, seed=2)
    for _ in range(20):
        sz, st_, ops = G.draw_batch(fast.rng, 64, L, cfg, True)
        assert sz.min() >= L // 2 and sz.max() <= 2 * L - 1 and (st_ <= np.maximum(sz, L) - L).all() and st_.min() >= 0
        part = (ops[..., 0] >= 2)
        assert ((ops[..., 2] + ops[..., 3])[part] <= L).all() and set(np.unique(ops[..., 2][ops[..., 0] == 1])) <= {50, 60}
        assert all(len(set(r)) == cfg.num_layers for r in ops[..., 0])       # ops drawn without replacement
    fast.load(raw_l, lab_l, raw_u)
    torch.cuda.synchronize()
    # overlapped mode delivers the same batches as the in-line mode (same draws: same private generator state)
    a1, a2 = G.FixMatchBatcher(eng, cfg, seed=9), G.FixMatchBatcher(eng, cfg, seed=9)
    ref = []
    for _ in range(3):
        a1.load(raw_l, lab_l, raw_u)
        torch.cuda.synchronize()
        ref.append((eng.x_s.clone(), eng.y_l.clone(), eng.x_uw.clone()))
    a2.aug_l.calls = a2.aug_u.calls = 0
    a1.aug_l.calls = a1.aug_u.calls = 0
    a2.prefetch(raw_l, lab_l, raw_u)
    for i in range(3):
        a2.commit()
        if i < 2:
            a2.prefetch(raw_l, lab_l, raw_u)
        torch.cuda.synchronize()
        assert torch.equal(eng.y_l, ref[i][1]) and torch.equal(eng.x_uw, ref[i][2])
        assert torch.equal(eng.x_s[:Bl], ref[i][0][:Bl])
    # captured mode (one graph launch per batch): same weak views / labels as the in-line mode for the same draws
    a3 = G.FixMatchBatcher(eng, cfg, seed=9)
    a3.prefetch_captured(raw_l, lab_l, raw_u)
    for i in range(3):
        a3.commit()
        if i < 2:
            a3.prefetch_captured(raw_l, lab_l, raw_u)
        torch.cuda.synchronize()
        assert torch.equal(eng.y_l, ref[i][1]) and torch.equal(eng.x_uw, ref[i][2]) and torch.equal(eng.x_s[:Bl], ref[i][0][:Bl])
        v = eng.x_s[Bl:].reshape(Bu, -1)
        assert float(v.mean(1).abs().max()) < 1e-4 and float((v.std(1, unbiased=False) - 1).abs().max()) < 1e-3
    for t in (eng.x_s, eng.x_uw):
        v = t.view(t.shape[0], -1)
        assert float(v.mean(1).abs().max()) < 1e-4 and float((v.std(1, unbiased=False) - 1).abs().max()) < 1e-3
