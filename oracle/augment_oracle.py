"""CPU restatement (numpy/scipy, float64) of the reference's per-item augmentation pipeline -- TEST
INFRASTRUCTURE ONLY: imported by tests/, bench.py's cpu_baseline leg and __graft_entry__.smoke(),
never by the product path.

Row a15 of SURVEY.md section 8 (file:line relative to the reference tree):
  weak   RandomResizeCrop.__call__            src/utils/transforms.py:93-127
  strong RandAugment.__call__ / RandomApply   src/utils/transforms.py:574-583, 647-657
         AmplitudeScaling                      :340-351
         AdaptivePowerlineNoise                :480-502
         RandomPartialWhiteNoise / SineNoise   :504-509, 521-546
  Standardize                                  :301-310
  order inside __getitem__                     src/utils/semi_dataset.py:234-242
  shipped parameters                           configs/base/resnet18/fixmatch.yaml:56-79

Every function takes its random draws EXPLICITLY (a dict), so the CUDA kernels can be fed the very same
draws; `draw_*` reproduce the reference's `np.random` call ORDER, so that with the same global seed the
oracle consumes the same stream as the reference (that is what tests/test_oracle_golden.py pins against
outputs of the unmodified reference transforms, tests/golden/make_golden_aug.py).
"""
from __future__ import annotations

from typing import Dict, List, Optional, Tuple

import numpy as np
from scipy.signal import resample as _fourier_resample

OPS = ("amplitude_scaling", "powerline", "partial_white", "partial_sine")   # order of the YAML op list


# ---------------------------------------------------------------------------------------------
# draws (reference np.random call order)
# ---------------------------------------------------------------------------------------------
def draw_weak(L: int, target_length: int, scale_min: float = 0.5, scale_max: float = 2.0) -> Dict:
    """transforms.py:96-97,120: ratio ~ U(scale_min, scale_max); size = int(L*ratio); start ~ randint."""
    ratio = np.random.uniform(scale_min, scale_max)
    size = int(L * ratio)
    padded = max(size, target_length)
    start = int(np.random.randint(0, padded - target_length + 1))
    return {"ratio": float(ratio), "size": size, "start": start}


def draw_strong(C: int, L: int, num_layers: int = 3, prob: float = 0.5, level: int = 10) -> Dict:
    """RandAugment (transforms.py:647-657): choice of `num_layers` of the 4 ops without replacement, then
    per op RandomApply's rand() < prob (:575) followed by that op's own draws."""
    lv = level / 10.0
    order = np.random.choice(len(OPS), num_layers, replace=False)
    ops: List[Dict] = []
    for oi in order:
        name = OPS[int(oi)]
        d: Dict = {"op": name, "apply": bool(np.random.rand() < prob)}
        if d["apply"]:
            if name == "amplitude_scaling":            # :347: normal(1, sigma, x.shape), sigma = level/10*0.5
                d["scales"] = np.random.normal(1, lv * 0.5, size=(C, L))
            elif name == "powerline":                  # :497: rand() < 0.5 -> 50 Hz else 60 Hz
                d["freq"] = 50 if np.random.rand() < 0.5 else 60
            elif name == "partial_white":              # :523 randn(*x.shape) first, then :536-537
                d["noise"] = np.random.randn(C, L)
                d["count"] = int(np.random.uniform(0, lv * 0.5) * L)
                d["start"] = int(np.random.randint(0, L - d["count"]))
            elif name == "partial_sine":               # :536-537
                d["count"] = int(np.random.uniform(0, lv * 0.5) * L)
                d["start"] = int(np.random.randint(0, L - d["count"]))
        ops.append(d)
    return {"ops": ops}


# ---------------------------------------------------------------------------------------------
# transforms
# ---------------------------------------------------------------------------------------------
def fourier_resample(x: np.ndarray, num: int) -> np.ndarray:
    """scipy.signal.resample along the last axis (transforms.py:100), written out: keep the lowest
    min(num, L) frequencies of the real FFT, fix up the shared Nyquist bin, inverse real FFT of length num,
    scale by num/L."""
    L = x.shape[-1]
    X = np.fft.rfft(x, axis=-1)
    Y = np.zeros(x.shape[:-1] + (num // 2 + 1,), dtype=X.dtype)
    N = min(num, L)
    nyq = N // 2 + 1
    Y[..., :nyq] = X[..., :nyq]
    if N % 2 == 0:
        if num < L:
            Y[..., N // 2] *= 2.0
        elif L < num:
            Y[..., N // 2] *= 0.5
    return np.fft.irfft(Y, num, axis=-1) * (float(num) / float(L))


def nearest_label_resize(label: np.ndarray, size: int) -> np.ndarray:
    """interp1d(arange(L), label, kind='nearest')(linspace(0, L-1, size)) (transforms.py:104-111): scipy's
    'nearest' rounds half-way points DOWN (searchsorted on the midpoints, side='left')."""
    L = label.shape[-1]
    pos = np.linspace(0, L - 1, size)
    mid = np.arange(L - 1) + 0.5
    idx = np.searchsorted(mid, pos, side="left")
    return label[..., idx]


def weak_resize_crop(x: np.ndarray, label: Optional[np.ndarray], d: Dict, target_length: int):
    """RandomResizeCrop.__call__ with the draws given (transforms.py:93-127)."""
    size, start = d["size"], d["start"]
    xr = fourier_resample(x, size)
    lr = nearest_label_resize(label.astype(np.float64), size) if label is not None else None
    pad = target_length - size
    if pad > 0:
        left = pad // 2
        right = pad - left
        xr = np.pad(xr, ((0, 0), (left, right)), mode="constant")
        if lr is not None:
            lr = np.pad(lr, ((0, 0), (left, right)), mode="constant")
    xc = xr[:, start:start + target_length]
    if lr is not None:
        return xc, lr[:, start:start + target_length]
    return xc


def powerline_amplitude(x: np.ndarray) -> np.ndarray:
    """AdaptivePowerlineNoise._get_amplitude (transforms.py:487-491): (p95 - p5)/2 per lead, numpy's linear
    interpolation between order statistics."""
    return (np.percentile(x, 95, axis=1, keepdims=True) - np.percentile(x, 5, axis=1, keepdims=True)) / 2


def strong_augment(x: np.ndarray, d: Dict, fs: int = 250, level: int = 10) -> np.ndarray:
    """RandAugment over the four shipped ops with the draws given."""
    lv = level / 10.0
    C, L = x.shape
    x = x.copy()
    for op in d["ops"]:
        if not op["apply"]:
            continue
        if op["op"] == "amplitude_scaling":
            x = x * op["scales"]
        elif op["op"] == "powerline":
            t = np.expand_dims(np.arange(L) / fs, axis=0)
            x = x + powerline_amplitude(x) * np.sin(2 * np.pi * op["freq"] * t)
        elif op["op"] == "partial_white":
            noise = (lv * 1.0) * op["noise"]
            part = np.zeros_like(x)
            part[:, op["start"]:op["start"] + op["count"]] = noise[:, :op["count"]]
            x = x + part
        elif op["op"] == "partial_sine":
            t = np.expand_dims(np.arange(L) / L, axis=0)
            noise = (lv * 1.0) * np.sin(2 * np.pi * t / (0.5 / lv))
            part = np.zeros_like(x)
            part[:, op["start"]:op["start"] + op["count"]] = noise[:, :op["count"]]
            x = x + part
    return x


def standardize(x: np.ndarray) -> np.ndarray:
    """Standardize(axis=(-1,-2)) (transforms.py:301-310): population std, zeros where std == 0."""
    loc = np.mean(x, axis=(-1, -2), keepdims=True)
    scale = np.std(x, axis=(-1, -2), keepdims=True)
    return np.divide(x - loc, scale, out=np.zeros_like(x), where=scale != 0)


def labeled_item(x: np.ndarray, y: np.ndarray, target_length: int) -> Tuple[np.ndarray, np.ndarray, Dict]:
    """semi_dataset.py:193-197,235-239 for a labeled item (after resample/filter): weak -> standardize."""
    dw = draw_weak(x.shape[1], target_length)
    xw, yw = weak_resize_crop(x, y, dw, target_length)
    return standardize(xw).astype(np.float32), yw.astype(np.int64).squeeze(0), {"weak": dw}


def unlabeled_item(x: np.ndarray, target_length: int, fs: int = 250):
    """semi_dataset.py:193-197,235-242 for an unlabeled item: weak -> (standardize | strong -> standardize)."""
    dw = draw_weak(x.shape[1], target_length)
    xw = weak_resize_crop(x, None, dw, target_length)
    ds = draw_strong(xw.shape[0], xw.shape[1])
    xs = strong_augment(xw, ds, fs=fs)
    return standardize(xw).astype(np.float32), standardize(xs).astype(np.float32), {"weak": dw, "strong": ds}
