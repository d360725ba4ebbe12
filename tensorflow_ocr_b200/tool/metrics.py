"""Mirror of the reference's ``tool/metrics.py`` (lines 31-85): streaming TP/FP arrays, precision / recall, f-mean.

These are a handful of scalar operations on the outputs of ``tool.bboxes.bboxes_matching``; they run on the host in
fp32 exactly as the TF graph does (``tool/math.py:27-41`` safe_divide: 0 where the denominator is <= 0)."""
from __future__ import annotations

import numpy as np

__all__ = ["streaming_tp_fp_arrays", "precision_recall", "fmean", "StreamingTpFp"]


def _np(x):
    return x.detach().cpu().numpy() if hasattr(x, "detach") else np.asarray(x)


class StreamingTpFp:
    """The local variables of streaming_tp_fp_arrays (tool/metrics.py:46-49) and their update op (:52-57)."""

    def __init__(self):
        self.v_num_gbboxes = np.int32(0)
        self.v_tp = np.zeros((0,), bool)
        self.v_fp = np.zeros((0,), bool)

    def update(self, num_gbboxes, tp, fp):
        self.v_num_gbboxes = np.int32(self.v_num_gbboxes + np.sum(_np(num_gbboxes).astype(np.int32)))
        self.v_tp = np.concatenate([self.v_tp, _np(tp).astype(bool).reshape(-1)])
        self.v_fp = np.concatenate([self.v_fp, _np(fp).astype(bool).reshape(-1)])
        return self.value()

    def value(self):
        return self.v_num_gbboxes, self.v_tp, self.v_fp


def streaming_tp_fp_arrays(num_gbboxes, tp, fp, metrics_collections=None, updates_collections=None, name=None,
                           state=None):
    """tool/metrics.py:31-63.  TF returns (value tensors, update op) over graph-local variables; here the variables
    live in ``state`` (a StreamingTpFp, created when omitted): returns (value after this update, state)."""
    state = StreamingTpFp() if state is None else state
    return state.update(num_gbboxes, tp, fp), state


def precision_recall(num_gbboxes, tp, fp, scope=None):
    """tool/metrics.py:66-80 -> (precision, recall), float32."""
    tp = np.float32(np.sum(_np(tp).astype(np.float32), axis=0))
    fp = np.float32(np.sum(_np(fp).astype(np.float32), axis=0))
    n = np.float32(_np(num_gbboxes))
    recall = np.float32(tp / n) if n > 0 else np.float32(0)
    den = np.float32(tp + fp)
    precision = np.float32(tp / den) if den > 0 else np.float32(0)
    return precision, recall


def fmean(pre, rec):
    """tool/metrics.py:82-85: 2 p r / (p + r), unguarded like the reference (0/0 -> nan)."""
    pre, rec = np.float32(pre), np.float32(rec)
    with np.errstate(invalid="ignore", divide="ignore"):
        return np.float32(2) * pre * rec / (pre + rec)
