"""Synthetic LUDB-shaped batches: z-scored Gaussian strips, piecewise
constant 4-class labels {0 bg, 1 P, 2 QRS, 3 T}.  numpy PCG64 streams -> reproducible anywhere."""
from __future__ import annotations

from typing import Dict, Tuple

import numpy as np


def zscore(x: np.ndarray) -> np.ndarray:
    """Per-sample standardisation over (C, L), 0 where std == 0 (reference transforms.py:301-310)."""
    m = x.mean(axis=(-1, -2), keepdims=True)
    s = x.std(axis=(-1, -2), keepdims=True)
    return np.where(s > 0, (x - m) / np.where(s > 0, s, 1.0), 0.0).astype(np.float32)


def make_signals(rng: np.random.Generator, B: int, C: int, L: int) -> np.ndarray:
    return zscore(rng.standard_normal((B, C, L)).astype(np.float32))


def make_strong(rng: np.random.Generator, x_w: np.ndarray) -> np.ndarray:
    return zscore(x_w + 0.5 * rng.standard_normal(x_w.shape).astype(np.float32))


def make_labels(rng: np.random.Generator, B: int, L: int, fs: int = 250) -> np.ndarray:
    y = np.zeros((B, L), dtype=np.int64)
    for b in range(B):
        t = int(rng.uniform(0, 0.5) * fs)
        while t < L:
            p0 = t
            q0 = p0 + int(0.16 * fs)
            t0 = q0 + int(0.20 * fs)
            y[b, p0:min(L, p0 + int(0.10 * fs))] = 1
            y[b, min(L, q0):min(L, q0 + int(0.10 * fs))] = 2
            y[b, min(L, t0):min(L, t0 + int(0.20 * fs))] = 3
            t += int(rng.uniform(0.6, 1.2) * fs)
    return y


def make_batch(seed: int, Bl: int, Bu: int, C: int, L: int, fs: int = 250) -> Tuple[Dict, Dict]:
    """(labeled, unlabeled) batch dicts following the dataset contract
    (reference semi_dataset.py:235-244): {'ecg','target'} / {'ecg','ecg_aug'}."""
    rng = np.random.default_rng(seed)
    xl = make_signals(rng, Bl, C, L)
    yl = make_labels(rng, Bl, L, fs)
    xw = make_signals(rng, max(Bu, 1), C, L)[:Bu]
    xs = make_strong(rng, xw) if Bu else xw
    return {"ecg": xl, "target": yl}, {"ecg": xw, "ecg_aug": xs}
