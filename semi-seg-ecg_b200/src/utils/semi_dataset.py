"""Dataset / loader surface (reference src/utils/semi_dataset.py).  The batch CONTRACT is part of
the hot path ({'ecg': f32[B,C,L], 'target': i64[B,L]} / {'ecg', 'ecg_aug'}); file I/O (pickle/CSV
index, Butterworth filters, CPU augmentation) is out of scope, so
`build_seg_dataset` serves synthetic LUDB-shaped strips when `dataset.synthetic` is set and
otherwise explains what is missing."""
import numpy as np
import torch
from torch.utils.data import DataLoader, Dataset, DistributedSampler, RandomSampler, SequentialSampler

from semiseg_b200 import synthetic


class SyntheticSemiSegDataset(Dataset):
    def __init__(self, length: int, num_leads: int, signal_length: int, labeled: bool, seed: int = 0, fs: int = 250):
        self.length, self.C, self.L, self.labeled, self.seed, self.fs = length, num_leads, signal_length, labeled, seed, fs

    def __len__(self):
        return self.length

    def __getitem__(self, idx):
        rng = np.random.default_rng((self.seed, idx, int(self.labeled)))
        x = synthetic.make_signals(rng, 1, self.C, self.L)
        if self.labeled:
            y = synthetic.make_labels(rng, 1, self.L, self.fs)[0]
            return {"ecg": torch.from_numpy(x[0]), "target": torch.from_numpy(y)}
        xs = synthetic.make_strong(rng, x)
        return {"ecg": torch.from_numpy(x[0]), "ecg_aug": torch.from_numpy(xs[0])}


def build_seg_dataset(cfg: dict, split: str, mode: str = None, num_unlabeled: int = None):
    syn = cfg.get("synthetic", None)
    if not syn:
        raise NotImplementedError(
            "file-backed ECG datasets (pickle + CSV index, reference semi_dataset.py:50-244) are outside the "
            "accelerated hot path; set dataset.synthetic: {length: N, num_leads: C} to train on synthetic "
            "LUDB-shaped strips, or feed your own loaders to algorithms.*.train_one_epoch")
    n = int(syn.get("length", 256))
    if split in ("valid", "test"):
        n = int(syn.get("valid_length", max(n // 8, 1)))
    labeled = split != "train_unlabeled"
    if split == "train_labeled" and num_unlabeled:
        n = num_unlabeled  # labeled set tiled to the unlabeled length (semi_dataset.py:86-95)
    return SyntheticSemiSegDataset(n, int(syn.get("num_leads", 1)), int(cfg.get("signal_length", 2500)), labeled,
                                   seed={"train_labeled": 11, "train_unlabeled": 23, "valid": 37, "test": 41}.get(split, 53))


def get_dataloader(dataset, is_distributed: bool = False, mode: str = "train", **kwargs):
    is_train = mode == "train"
    if is_distributed and is_train:
        sampler = DistributedSampler(dataset, shuffle=True)
    elif is_train:
        sampler = RandomSampler(dataset)
    else:
        sampler = SequentialSampler(dataset)
    kwargs = dict(kwargs)
    kwargs.setdefault("pin_memory", True)
    return DataLoader(dataset, sampler=sampler, drop_last=is_train, **kwargs)
