"""Algorithm registry: `algorithms.__dict__[config['algorithm']].train(config)` (reference
src/train.py:81-86, src/algorithms/__init__.py:1-6).  `base` is registered as `scratch`.
cps and stpp run on the same step engine; reco's contrastive head is not built."""
from . import base, cps, fixmatch, mean_teacher, stpp  # noqa: F401
from . import base as scratch  # noqa: F401
