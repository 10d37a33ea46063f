import os
import sys

import pytest

REPO = os.path.abspath(os.path.join(os.path.dirname(__file__), ".."))
SRC = os.path.join(REPO, "semi-seg-ecg_b200", "src")
for p in (REPO, SRC):
    if p not in sys.path:
        sys.path.insert(0, p)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (B200); run with -m gpu on the GPU box")


def pytest_collection_modifyitems(config, items):
    import torch
    if torch.cuda.is_available():
        return
    skip = pytest.mark.skip(reason="no CUDA device")
    for item in items:
        if "gpu" in item.keywords:
            item.add_marker(skip)


@pytest.fixture(scope="session")
def golden():
    import numpy as np
    return np.load(os.path.join(REPO, "tests", "golden", "reference_vectors.npz"))


@pytest.fixture(scope="session")
def golden_semi():
    """cases F (CPS) and G (ST++), tests/golden/make_golden_semi.py"""
    import numpy as np
    return np.load(os.path.join(REPO, "tests", "golden", "semi_vectors.npz"))


@pytest.fixture(scope="session")
def golden_eval():
    """case H (evaluate), tests/golden/make_golden_eval.py"""
    import numpy as np
    return np.load(os.path.join(REPO, "tests", "golden", "eval_vectors.npz"))


@pytest.fixture(scope="session")
def golden_arch():
    """case R (resnet34), tests/golden/make_golden_arch.py"""
    import numpy as np
    return np.load(os.path.join(REPO, "tests", "golden", "arch_vectors.npz"))


@pytest.fixture(scope="session")
def golden_clip():
    """case M (FixMatch with max_norm clipping), tests/golden/make_golden_clip.py"""
    import numpy as np
    return np.load(os.path.join(REPO, "tests", "golden", "clip_vectors.npz"))


def golden_group(g, prefix):
    """{'name': array} for all keys under 'prefix/'."""
    pre = prefix + "/"
    return {k[len(pre):]: g[k] for k in g.files if k.startswith(pre)}
