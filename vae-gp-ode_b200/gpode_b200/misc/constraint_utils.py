"""Positive-parameter maps (reference: experiments/model/misc/constraint_utils.py:5-13)."""
import torch
import torch.nn.functional as F

LOWER = 1e-12


def softplus(x):
    """raw -> positive: log(1 + e^x) + 1e-12."""
    return F.softplus(x) + LOWER


def invsoftplus(y):
    """positive -> raw; clamps at machine eps like the reference so tiny values stay finite."""
    eps = torch.finfo(y.dtype).eps
    v = torch.clamp(y - LOWER, min=eps)
    return v + torch.log(-torch.expm1(-v))
