"""B200-native kernels + step engine for the SemiSegECG training hot path."""
from . import _lib  # noqa: F401
from ._lib import ALGO_SIMT, ALGO_TCGEN05, BF16, F32  # noqa: F401
