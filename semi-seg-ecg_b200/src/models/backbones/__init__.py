"""Backbone registry: `backbones.__dict__[name](**kwargs)` as used by algorithms/base.py:34-37
of the reference (reference exports: backbones/__init__.py:1-2)."""
from .resnet import *  # noqa: F401,F403


def _out_of_scope(name):
    def ctor(*args, **kwargs):
        raise NotImplementedError(
            f"{name}: this repository accelerates the 1-D ResNet + FCNHead segmentor only; the ViT backbones of the "
            "reference (src/models/backbones/vision_transformer.py) are out of scope (SURVEY.md section 2).")
    ctor.__name__ = name
    return ctor


# the names a reference YAML can ask for: a clear message instead of a KeyError
vit_tiny, vit_small, vit_base = (_out_of_scope(n) for n in ("vit_tiny", "vit_small", "vit_base"))
