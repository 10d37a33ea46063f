"""Algorithm registry: `algorithms.__dict__[config['algorithm']].train(config)` (reference
src/train.py:81-86, src/algorithms/__init__.py:1-6).  `base` is registered as `scratch`.
cps / reco / stpp are outside the accelerated hot path."""
from . import base, fixmatch, mean_teacher  # noqa: F401
from . import base as scratch  # noqa: F401
