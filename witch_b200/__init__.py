"""witch_b200: B200-native eHMM score + align path for WITCH (CUDA sm_100a behind a C ABI)."""
from ._lib import WitchError, LIB_PATH  # noqa: F401
from .api import EHMM, Queries, score, score_dev, weights_topk, weights_topk_dev, align, graph_align, merge_rows, debug_fwdbwd, kernel_launches  # noqa: F401

__version__ = "0.1.0"
