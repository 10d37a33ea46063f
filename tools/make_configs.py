"""Emit the YAML configs of the drop-in surface (schema and values of the reference's
configs/base/resnet18/{scratch,fixmatch,mean_teacher,cps,stpp}.yaml and configs/bench/**; SURVEY.md
sections 5 and 8d).  Written from the schema description, not copied; run once, output committed."""
import copy
import os

import yaml

ROOT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "semi-seg-ecg_b200", "configs")
FILL = "<FILL IN>"


def base(algorithm: str) -> dict:
    strong = [{"RandAugment": {
        "ops": [{"AmplitudeScaling": {"sigma": 0.5}}, {"AdaptivePowerlineNoise": {"fs": 250}},
                {"RandomPartialWhiteNoise": {"amplitude": 1, "ratio": 0.5}},
                {"RandomPartialSineNoise": {"amplitude": 1, "ratio": 0.5}}],
        "level": 10, "num_layers": 3, "prob": 0.5}}]
    dataset = {"ecg_dir": FILL, "label_dir": FILL, "index_dir": FILL}
    if algorithm != "base":
        dataset["train_unlabeled_csv"] = FILL
    dataset.update({"train_labeled_csv": FILL, "valid_csv": FILL, "test_csv": FILL,
                    "filename_col": "waveform", "label_filename_col": "label", "signal_length": 2500,
                    "filter": [{"highpass_filter": {"fs": 250, "cutoff": 0.67}},
                               {"lowpass_filter": {"fs": 250, "cutoff": 40}}],
                    "augmentations": [{"random_resize_crop": {"target_length": 2500, "scale_min": 0.5, "scale_max": 2.0}}]})
    if algorithm not in ("base", "cps"):      # CPS trains on the weak view only
        dataset["strong_augmentations"] = strong
    dataset["transforms"] = [{"standardize": {"axis": [-1, -2]}}, {"to_tensor": {"dtype": "float"}}]
    train = {"epochs": 100, "accum_iter": 1, "warmup_epochs": 10, "min_lr": 0.0001, "blr": None, "lr": 0.001,
             "weight_decay": 0.05, "max_norm": None, "layer_decay": None, "optimizer": "adamw",
             "optimizer_kwargs": {"betas": [0.9, 0.999]}, "auxiliary_loss_weight": [0.4]}
    if algorithm == "fixmatch":
        train["conf_thresh"] = 0.80
    if algorithm in ("mean_teacher", "stpp"):
        train["ema_decay"] = 0.99
    name = {"base": "scratch"}.get(algorithm, algorithm)
    return {
        "seed": 0, "output_dir": f"../exps/resnet18/{name}", "exp_name": FILL, "resume": None, "start_epoch": 0,
        "device": "cuda", "use_amp": True, "algorithm": algorithm, "mode": "scratch", "pretrained_backbone": None,
        "backbone": {"resnet18": {"num_leads": 1, "num_stages": 4, "out_indices": [0, 1, 2, 3],
                                  "dilations": [1, 1, 1, 1], "strides": [1, 2, 2, 2], "deep_stem": False,
                                  "avg_down": False, "contract_dilation": False}},
        "decode_head": {"FCNHead": {"in_channels": 512, "in_index": 3, "channels": 128, "num_convs": 1,
                                    "concat_input": False, "dropout_ratio": 0.1, "num_classes": 4,
                                    "align_corners": False}},
        "dataset": dataset,
        "dataloader": {"batch_size": 16, "num_workers": 2, "pin_memory": False},
        "train": train,
        "metric": {"task": "segmentation", "compute_on_cpu": True, "sync_on_compute": False, "num_classes": 4,
                   "include_background": True, "per_class": False, "input_format": "one-hot",
                   "target_metrics": ["MeanIoU"]},
        "test": {"target_metric": "MeanIoU"},
        "ddp": {"world_size": 1, "rank": -1, "gpu": 0, "dist_url": "env://", "dist_backend": "nccl",
                "distributed": False, "sync_bn": True},
    }


def bench(ds: str, frac: str) -> dict:
    prefix, ext = {"ludb": ("LUDB", "csv"), "qtdb": ("QTDB", "csv"), "isp": ("ISP", "csv"),
                   "zhejiang": ("Zhejiang", "pkl")}[ds]
    if ds == "zhejiang" and frac == "1over16":
        ext = "csv"  # the reference's own 1over16 override names .csv files (zhejiang/1over16.yaml:7-10)
    return {"exp_name": f"{ds}/{frac}",
            "dataset": {"ecg_dir": f"../data/{ds}/ecg", "label_dir": f"../data/{ds}/label",
                        "index_dir": f"../index/{ds}",
                        "train_unlabeled_csv": f"{prefix}_train_unlabeled.{ext}",
                        "train_labeled_csv": f"{prefix}_train_labeled_{frac}.{ext}",
                        "valid_csv": f"{prefix}_valid.{ext}", "test_csv": f"{prefix}_test.{ext}"}}


def dump(path: str, obj: dict):
    os.makedirs(os.path.dirname(path), exist_ok=True)
    with open(path, "w") as f:
        yaml.safe_dump(obj, f, default_flow_style=False, sort_keys=False)


if __name__ == "__main__":
    for algo, fname in (("base", "scratch"), ("fixmatch", "fixmatch"), ("mean_teacher", "mean_teacher"), ("cps", "cps"),
                        ("stpp", "stpp")):
        cfg = base(algo)
        dump(os.path.join(ROOT, "base", "resnet18", f"{fname}.yaml"), cfg)
        dump(os.path.join(ROOT, "base", f"{fname}.yaml"), cfg)  # README-style alias path
    for ds in ("ludb", "qtdb", "isp", "zhejiang"):
        for frac in ("1over2", "1over4", "1over8", "1over16"):
            dump(os.path.join(ROOT, "bench", ds, f"{frac}.yaml"), bench(ds, frac))
    dump(os.path.join(ROOT, "bench", "cross_domain", "merged.yaml"),
         {"exp_name": "cross_domain/merged",
          "dataset": {"ecg_dir": "../data", "label_dir": "../data", "index_dir": "../index/cross_domain",
                      "train_unlabeled_csv": "merged_unlabeled.csv", "train_labeled_csv": "merged_train_labeled.csv",
                      "valid_csv": "merged_valid.csv", "test_csv": "merged_test.csv"}})
    # synthetic-data override so `python train.py -f ... -o configs/bench/synthetic.yaml` runs end to end
    dump(os.path.join(ROOT, "bench", "synthetic.yaml"),
         {"exp_name": "synthetic", "dataset": {"synthetic": {"length": 256, "valid_length": 32, "num_leads": 1}}})
    print("configs written under", os.path.abspath(ROOT))
