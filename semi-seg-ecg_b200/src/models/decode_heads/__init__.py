"""Decode-head registry: `decode_heads.__dict__[name](**kwargs)` (reference algorithms/base.py:40-43)."""
from .fcn_head import FCNHead  # noqa: F401
