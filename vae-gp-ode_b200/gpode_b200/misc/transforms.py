"""Constrained <-> unconstrained maps used by Param (reference: experiments/model/misc/transforms.py:8-81).

Same class names and methods (forward / backward on numpy, forward_tensor / backward_tensor on torch).
LowerTriangular scatters all D packed rows in one indexed assignment on the device instead of the
reference's per-matrix Python loop (transforms.py:71-77); the result is identical.
"""
import numpy as np
import torch
import torch.nn.functional as F

from .settings import settings


class Identity:
    def __str__(self):
        return "Identity transformation"

    def forward(self, x):
        return x

    def backward(self, y):
        return y

    def forward_tensor(self, x):
        return x

    def backward_tensor(self, y):
        return y


class SoftPlus:
    def __init__(self, lower=1e-12):
        self._lower = lower

    def __str__(self):
        return "Softplus transformation"

    def forward(self, x):
        return np.logaddexp(0, x) + self._lower

    def backward(self, y):
        v = np.maximum(y - self._lower, np.finfo(settings.numpy_float).eps)
        return v + np.log(-np.expm1(-v))

    def forward_tensor(self, x):
        return F.softplus(x) + self._lower

    def backward_tensor(self, y):
        v = torch.clamp(y - self._lower, min=torch.finfo(y.dtype).eps)
        return v + torch.log(-torch.expm1(-v))


class LowerTriangular:
    """packed rows (D, N(N+1)/2) in row-major tril order <-> (D, N, N) lower-triangular matrices."""

    def __init__(self, N, num_matrices=1, device="cpu"):
        self.N = N
        self.num_matrices = num_matrices
        self.device = device
        self._idx = {}

    def __str__(self):
        return "Lower cholesky transformation"

    def forward(self, x):
        out = np.zeros((self.num_matrices, self.N, self.N), dtype=settings.numpy_float)
        r, c = np.tril_indices(self.N)
        out[:, r, c] = x
        return out

    def backward(self, y):
        r, c = np.tril_indices(self.N)
        return np.asarray(y)[:, r, c]

    def _indices(self, device):
        key = str(device)
        if key not in self._idx:
            self._idx[key] = torch.tril_indices(self.N, self.N, 0, device=device)
        return self._idx[key]

    def forward_tensor(self, x):
        rc = self._indices(x.device)
        out = torch.zeros((self.num_matrices, self.N, self.N), dtype=x.dtype, device=x.device)
        out[:, rc[0], rc[1]] = x
        return out

    def backward_tensor(self, y):
        rc = self._indices(y.device)
        return y[:, rc[0], rc[1]]
