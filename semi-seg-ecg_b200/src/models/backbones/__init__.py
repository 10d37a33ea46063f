"""Backbone registry: `backbones.__dict__[name](**kwargs)` as used by algorithms/base.py:34-37
of the reference (reference exports: backbones/__init__.py:1-2)."""
from .resnet import *  # noqa: F401,F403
from .vision_transformer import *  # noqa: F401,F403
