"""Covariance modules of the sparse GP -- drop-in for experiments/model/core/kernels.py.

Same class names, constructor arguments, parameter names (``unconstrained_lengthscales``,
``unconstrained_variance`` -> identical state_dict keys), properties (``lengthscales``, ``variance``)
and cache attributes (``rff_weights``, ``rff_omega``, ``rff_phase``, ``nu``).  One extra attribute,
``rff_eps``, keeps the standard-normal frequency draws so that the CUDA kernels can form
omega = eps / ell themselves and return the lengthscale gradient of the random features.

The per-STATE arithmetic (rff_forward + f_update at every solver stage) is NOT done here: it lives in
libgpode.so and is reached through SVGP_Layer.forward / Flow.forward.  The torch code below is the
per-ROLLOUT setup at the M inducing points only (Kzz, prior at Z, Cholesky solve): reference
kernels.py:98-110,140-153,155-172 (RBF) and :289-303,319-351,376-387 (DF).
"""
import math

import numpy as np
import torch
from torch import nn

from ..misc.constraint_utils import invsoftplus, softplus

jitter = 1e-5


def sample_normal(shape, seed=None):
    """host draw ~ N(0,1), fp32 (reference kernels.py:13-18; patched by the parity tests)."""
    rng = np.random.RandomState() if seed is None else np.random.RandomState(seed)
    return torch.tensor(rng.normal(size=shape).astype(np.float32))


def sample_uniform(shape, seed=None):
    """host draw ~ U(0,1), fp32 (reference kernels.py:20-26)."""
    rng = np.random if seed is None else np.random.RandomState(seed)
    return torch.tensor(rng.uniform(low=0.0, high=1.0, size=shape).astype(np.float32))


class RBF(nn.Module):
    """Squared-exponential kernel with random-Fourier-feature prior draws (reference kernels.py:29-195)."""

    variant_dimwise = "rbf_dimwise"
    variant_shared = "rbf_shared"

    def __init__(self, D_in, D_out=None, dimwise=False):
        super().__init__()
        self.D_in = D_in
        self.D_out = D_in if D_out is None else D_out
        self.dimwise = dimwise
        ell_shape = (self.D_out, self.D_in) if dimwise else (self.D_in,)
        var_shape = (self.D_out,) if dimwise else (1,)
        self.unconstrained_lengthscales = nn.Parameter(torch.full(ell_shape, invsoftplus(torch.tensor(0.2)).item()))
        self.unconstrained_variance = nn.Parameter(torch.full(var_shape, invsoftplus(torch.tensor(0.1)).item()))
        self.rff_weights = self.rff_omega = self.rff_phase = self.rff_eps = self.nu = None

    @property
    def variant(self):
        return self.variant_dimwise if self.dimwise else self.variant_shared

    @property
    def lengthscales(self):
        return softplus(self.unconstrained_lengthscales)

    @property
    def variance(self):
        return softplus(self.unconstrained_variance)

    # ---- setup-time covariance at the inducing points --------------------------------------------
    def K(self, X, X2=None):
        """dimwise: (D_out,N,M); shared: (N,M).  var * exp(-1/2 |(X - X2)/ell|^2)."""
        X2 = X if X2 is None else X2
        diff = X[:, None, :] - X2[None, :, :]
        if self.dimwise:
            sq = (diff[None] / self.lengthscales[:, None, None, :]).square().sum(-1)
            return self.variance[:, None, None] * torch.exp(-0.5 * sq)
        sq = (diff / self.lengthscales).square().sum(-1)
        return self.variance * torch.exp(-0.5 * sq)

    forward = K

    def sample_freq(self, S, seed=None, device="cpu"):
        """omega = eps / ell with eps ~ N(0,1) (D_in,S,D_out) or (D_in,S); keeps eps in self.rff_eps."""
        shape = (self.D_in, S, self.D_out) if self.dimwise else (self.D_in, S)
        self.rff_eps = sample_normal(shape, seed).to(device)
        ell = self.lengthscales.t().unsqueeze(1) if self.dimwise else self.lengthscales.unsqueeze(1)
        return self.rff_eps / ell

    def build_cache(self, S, device):
        """draw order of the reference (kernels.py:126-137): weights, frequencies, phases."""
        self.rff_weights = sample_normal((S, self.D_out)).to(device)
        self.rff_omega = self.sample_freq(S, device=device)
        shape = (1, S, self.D_out) if self.dimwise else (1, S)
        self.rff_phase = sample_uniform(shape).to(device) * 2 * np.pi

    def rff_forward(self, x, S):
        """prior sample at x (setup path, used at Z): sqrt(var/S) sum_s w_s cos(x.omega_s + b_s)."""
        if self.dimwise:
            theta = torch.einsum("nd,dsk->nsk", x, self.rff_omega) + self.rff_phase
            return torch.sqrt(self.variance / S) * (torch.cos(theta) * self.rff_weights).sum(1)
        theta = x @ self.rff_omega + self.rff_phase
        return torch.sqrt(self.variance / S) * (torch.cos(theta) @ self.rff_weights)

    def compute_nu(self, Ku, u_prior, inducing_val):
        """nu = L^-T (u - L^-1 f_p(Z)), L = chol(Ku + jitter I) (whitened pathwise update, eq. 13 of Wilson et al. 2020)."""
        M = Ku.shape[-1]
        L = torch.linalg.cholesky(Ku + jitter * torch.eye(M, device=Ku.device, dtype=Ku.dtype))
        if self.dimwise:
            a = torch.linalg.solve_triangular(L, u_prior.t().unsqueeze(2), upper=False)
            self.nu = torch.linalg.solve_triangular(L.transpose(1, 2), inducing_val.t().unsqueeze(2) - a, upper=True)
        else:
            a = torch.linalg.solve_triangular(L, u_prior, upper=False)
            self.nu = torch.linalg.solve_triangular(L.t(), inducing_val - a, upper=True)

    def f_update(self, x, x2):
        """setup-path pathwise update K(x2,x)^T nu -> (N,D_out)."""
        Kuf = self.K(x2, x)
        if self.dimwise:
            return torch.einsum("km,kmn->nk", self.nu.squeeze(2), Kuf)
        return Kuf.t() @ self.nu

    def extra_cache(self):
        return None


class DivergenceFreeKernel(RBF):
    """Matrix-valued divergence-free kernel (reference kernels.py:201-393): needs D_in == D_out = D,
    lengthscales (D,D) indexed by block entry, variance (D,) on the column index."""

    def __init__(self, D_in, D_out):
        super().__init__(D_in=D_in, D_out=D_out, dimwise=True)
        self.rff_B = None

    @property
    def variant(self):
        return "df"

    def K(self, X, X2=None):
        """(N*D, M*D), row n*D+i, col m*D+j:
        var_j exp(-r2/(2 l_ij^2))/l_ij^2 (d_i d_j/l_ij^2 + delta_ij((D-1) - r2/l_ij^2)), d = X2_m - X_n."""
        X2 = X if X2 is None else X2
        N, D = X.shape
        M = X2.shape[0]
        d = X2[None, :, :] - X[:, None, :]
        r2 = d.square().sum(-1)[:, :, None, None]
        c = self.lengthscales.pow(-2)
        E = torch.exp(-0.5 * r2 * c)
        H = d[:, :, :, None] * d[:, :, None, :] * c + torch.eye(D, device=X.device, dtype=X.dtype) * ((D - 1.0) - r2 * c)
        Kb = self.variance * E * H * c
        return Kb.permute(0, 2, 1, 3).reshape(N * D, M * D)

    forward = K

    @staticmethod
    def operator_B(omega):
        """B[s,a,c] = |omega[:,s,c]| delta_ac - sum_b omega[a,s,b] omega[c,s,b] / |omega[:,s,c]| (kernels.py:327-336);
        state independent, so it is built once per rollout instead of once per solver stage."""
        D = omega.shape[0]
        norm = omega.square().sum(0).sqrt()
        ww = torch.einsum("asb,csb->sac", omega, omega)
        return norm[:, None, :] * torch.eye(D, device=omega.device, dtype=omega.dtype) - ww / norm[:, None, :]

    @staticmethod
    def operator_B_batched(omega):
        """operator_B for a stack of samples: omega (L,D,S,D) -> (L,S,D,D), one einsum"""
        D = omega.shape[1]
        norm = omega.square().sum(1).sqrt()                                   # (L,S,D)
        ww = torch.einsum("lasb,lcsb->lsac", omega, omega)
        return norm[:, :, None, :] * torch.eye(D, device=omega.device, dtype=omega.dtype) - ww / norm[:, :, None, :]

    def build_cache(self, S, device):
        self.rff_weights = sample_normal((2 * S, self.D_out)).to(device)
        self.rff_omega = self.sample_freq(S, device=device)
        self.rff_phase = sample_uniform((1, S, self.D_out)).to(device) * 2 * np.pi
        self.rff_B = self.operator_B(self.rff_omega)

    def rff_forward(self, x, S):
        theta = torch.einsum("nd,dsa->nsa", x, self.rff_omega) + self.rff_phase
        u = torch.cos(theta) * self.rff_weights[:S] + torch.sin(theta) * self.rff_weights[S:]
        B = self.rff_B if self.rff_B is not None else self.operator_B(self.rff_omega)
        return torch.sqrt(self.variance / S) * torch.einsum("nsa,sac->nc", u, B)

    def compute_nu(self, Ku, u_prior, inducing_val):
        n = Ku.shape[0]
        L = torch.linalg.cholesky(Ku + jitter * torch.eye(n, device=Ku.device, dtype=Ku.dtype))
        a = torch.linalg.solve_triangular(L, u_prior.reshape(n, 1), upper=False)
        self.nu = torch.linalg.solve_triangular(L.t(), inducing_val.reshape(n, 1) - a, upper=True)

    def f_update(self, x, x2):
        return (self.nu[:, 0] @ self.K(x2, x)).reshape(x.shape)

    def extra_cache(self):
        return self.rff_B
