"""Param: an nn.Module holding one unconstrained tensor `optvar`; calling it returns the constrained
value (reference: experiments/model/misc/param.py:7-28 -- same constructor, same `optvar` state_dict key)."""
import numpy as np
import torch

from . import transforms
from .settings import settings


class Param(torch.nn.Module):
    def __init__(self, value, transform=None, name="var", device=None):
        super().__init__()
        self.transform = transform if transform is not None else transforms.Identity()
        self.name = name
        raw = np.asarray(self.transform.backward(value))
        self.optvar = torch.nn.Parameter(torch.tensor(raw, dtype=settings.torch_float, device=device or settings.device))

    def __call__(self):
        return self.transform.forward_tensor(self.optvar)

    def __repr__(self):
        return "{} parameter with {}".format(self.name, self.transform)
