"""Time-grid helpers of experiments/model/misc/torch_utils.py:49-61 (the rest of that file -- checkpoint / seeding / numpy
conversion helpers -- is reference code outside the hot path and stays where it is)."""
import torch


def insert_zero_t0(ts):
    """Given a time span ts, insert an additional time zero in front (torch_utils.py:49-51)."""
    return torch.cat([torch.zeros(1, dtype=ts.dtype, device=ts.device), ts + ts[1] - ts[0]])


def compute_ts_dense(ts, ts_dense_scale):
    """Densify a time grid: every interval [t1, t2] becomes linspace(t1, t2, ts_dense_scale)[:-1], i.e. ts_dense_scale - 1 sub-steps
    (torch_utils.py:54-61).  ``dense[:: ts_dense_scale - 1]`` are the original points."""
    if ts_dense_scale > 1:
        pieces = [torch.linspace(float(t1), float(t2), ts_dense_scale, dtype=ts.dtype, device=ts.device)[:-1] for t1, t2 in zip(ts[:-1], ts[1:])]
        return torch.cat(pieces + [ts[-1:]])
    return ts
