"""hbird_b200 — B200-native backend for Open Hummingbird's dense nearest-neighbour evaluation.

Host-side mirror of the reference interface for ONE path (memory bank -> kNN -> label transfer ->
mIoU); all arithmetic runs in hand-written sm_100a kernels behind the C-ABI in
include/hbird_b200.h.  Importing this package loads libhbird_b200.so and fails if it is missing.
"""
from . import _capi  # noqa: F401  (loads the native library; raises if absent)
from .hbird_eval import HbirdEvaluation, hbird_evaluation  # noqa: F401
from .nn.search_b200 import NearestNeighborSearchB200  # noqa: F401
from .registry import NN_BACKENDS, nn_method_choices, parse_nn_params, register_nn_backend  # noqa: F401
from .utils.eval_metrics import PredsmIoU  # noqa: F401

__version__ = "0.1.0"
