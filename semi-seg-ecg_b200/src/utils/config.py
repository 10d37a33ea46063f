"""YAML base + deep-merged override + CLI overwrite (reference src/train.py:14-76), without the
`mergedeep` dependency."""
import copy

import yaml


def deep_merge(base: dict, override: dict) -> dict:
    out = copy.deepcopy(base)
    for k, v in (override or {}).items():
        if isinstance(v, dict) and isinstance(out.get(k), dict):
            out[k] = deep_merge(out[k], v)
        else:
            out[k] = copy.deepcopy(v)
    return out


def load_config(config_path: str, override_path: str = None, cli: dict = None) -> dict:
    with open(config_path, "r") as f:
        config = yaml.safe_load(f)
    if override_path:
        with open(override_path, "r") as f:
            config = deep_merge(config, yaml.safe_load(f))
    for k, v in (cli or {}).items():
        if v:
            config[k] = v
    return config
