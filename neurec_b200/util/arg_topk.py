"""arg_topk / dtype helpers of util/cython/{arg_topk,tools}.pyx on the GPU kernels."""
import numpy as np

from .. import ops

float_type = np.float32      # tools.pyx:26
int_type = np.int32          # tools.pyx:27


def is_ndarray(array, dtype):
    """tools.pyx:30-37: ndarray of that dtype that owns its data (views are rejected)."""
    return isinstance(array, np.ndarray) and array.dtype == dtype and array.base is None


def arg_topk(ranking_scores, top_k=50, thread_num=None):
    """Row-wise indices of the top_k scores, std::partial_sort_copy order (arg_topk.pyx:16-35).
    `thread_num` is accepted for signature compatibility and ignored."""
    if not is_ndarray(ranking_scores, float_type):
        ranking_scores = np.array(ranking_scores, dtype=float_type)
    return ops.arg_topk_host(ranking_scores, int(top_k))
