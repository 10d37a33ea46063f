"""Golden vectors for the augmentation row from the UNMODIFIED reference transforms
(TEST INFRASTRUCTURE; run in the build container only: needs /root/reference or $SEMISEG_REF):

    python tests/golden/make_golden_aug.py

Builds the reference's own weak / strong / transform pipelines from the shipped YAML parameters
(configs/base/resnet18/fixmatch.yaml:56-83) with `utils.transforms.get_transforms_from_config`-equivalent
constructors, seeds numpy's global RNG, pushes synthetic strips through them in the order of
`ECGSemiSegDataset.__getitem__` (semi_dataset.py:193-197, 235-242) and stores inputs, seeds and outputs
(numbers only) in tests/golden/augment_vectors.npz.  tests/test_oracle_golden.py replays the same seeds
through oracle/augment_oracle.py and requires equal results.
"""
import importlib.util
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
REPO = os.path.abspath(os.path.join(HERE, "..", ".."))
REF = os.environ.get("SEMISEG_REF", "/root/reference")

spec = importlib.util.spec_from_file_location("ref_transforms", os.path.join(REF, "src", "utils", "transforms.py"))
T = importlib.util.module_from_spec(spec)
spec.loader.exec_module(T)

CASES = [  # (seed, leads, L)
    (11, 1, 2500), (12, 1, 2500), (13, 1, 2500), (14, 1, 2500), (15, 1, 2500), (16, 1, 2500),
    (21, 2, 1000), (22, 2, 1000), (23, 3, 601), (24, 1, 600), (25, 1, 128), (26, 12, 500),
]


def synth(rng, C, L):
    t = np.arange(L) / 250.0
    x = np.zeros((C, L))
    for c in range(C):
        x[c] = (0.6 * np.sin(2 * np.pi * (1.1 + 0.1 * c) * t + rng.uniform(0, 6.28)) + 0.25 * np.sin(2 * np.pi * 7.3 * t)
                + 0.15 * rng.standard_normal(L)) + 0.3 * c
    y = np.zeros((1, L), dtype=np.int64)
    pos = 0
    while pos < L:
        seg = int(rng.integers(8, 60))
        y[0, pos:pos + seg] = int(rng.integers(0, 4))
        pos += seg
    return x, y


def main():
    out = {}
    for seed, C, L in CASES:
        rng = np.random.default_rng(seed)
        x, y = synth(rng, C, L)
        weak = T.RandomResizeCrop(target_length=L, scale_min=0.5, scale_max=2.0)
        strong = T.RandAugment(ops=[T.AmplitudeScaling(sigma=0.5), T.AdaptivePowerlineNoise(fs=250),
                                    T.RandomPartialWhiteNoise(amplitude=1, ratio=0.5),
                                    T.RandomPartialSineNoise(amplitude=1, ratio=0.5)], level=10, num_layers=3, prob=0.5)
        std = T.Standardize(axis=[-1, -2])
        pre = f"s{seed}"
        out[pre + "/x"], out[pre + "/y"] = x, y
        # labeled item
        np.random.seed(seed)
        xw, yw = weak(x.copy(), y.astype(np.float64).copy())
        out[pre + "/lab_ecg"] = std(xw).astype(np.float32)
        out[pre + "/lab_target"] = yw.astype(np.int64).squeeze(0)
        # unlabeled item
        np.random.seed(seed + 1000)
        xw = weak(x.copy())
        out[pre + "/unl_ecg"] = std(xw).astype(np.float32)
        xs = strong(xw)
        out[pre + "/unl_ecg_aug"] = std(xs).astype(np.float32)
        out[pre + "/unl_weak_raw"] = xw
        out[pre + "/unl_strong_raw"] = xs
    out["cases"] = np.array(CASES, dtype=np.int64)
    path = os.path.join(HERE, "augment_vectors.npz")
    np.savez_compressed(path, **out)
    print("wrote", path, os.path.getsize(path), "bytes")


if __name__ == "__main__":
    main()
