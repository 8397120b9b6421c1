"""SVGP_Layer -- drop-in for experiments/model/core/svpy.py:30-175 (decoupled sampling of a sparse GP
posterior, Wilson et al. 2020).  Same constructor, parameters (``inducing_loc``, ``Um``, ``Us_sqrt``
as Param modules -> same state_dict keys), ``build_cache`` / ``forward`` / ``kl`` /
``sample_inducing``.  ``forward`` evaluates the field with the CUDA kernels of libgpode.so."""
import sys

import numpy as np
import torch

from .. import functional as GF
from ..misc import transforms
from ..misc.param import Param
from .kernels import RBF, DivergenceFreeKernel

jitter = 1e-5


def sample_normal(shape, seed=None):
    """host draw ~ N(0,1) from the global numpy RNG (reference svpy.py:12-18; patched by the parity tests)."""
    rng = np.random if seed is None else np.random.RandomState(seed)
    return torch.tensor(rng.normal(size=shape).astype(np.float32))


_rng = {"mode": "host", "stream": None}


def set_rng(mode, seed=0):
    """Where the function-sample draws (w, eps, phase, eps_u) come from.
    "host" (default): the numpy helpers above and in core/kernels.py, like the reference -- and like the reference they are what the
    parity tests patch to replay recorded draws; every draw is a host array copied to the device.
    "device": a Philox4x32-10 stream on the GPU (functional.PhiloxStream, csrc/elbo_kernels.cu): all four tensors of all L samples in
    one launch, no host RNG time and no H2D copy; same distributions, different numbers (the reference's own draws are unseeded,
    kernels.py:17, so no run of it is reproducible either).  Only the fused CUDA setup paths use it."""
    if mode not in ("host", "device"):
        raise ValueError("rng mode must be 'host' or 'device'")
    _rng["mode"] = mode
    _rng["stream"] = GF.PhiloxStream(seed) if mode == "device" else None


class FieldSample:
    """The tensors that pin one function sample (what build_cache leaves on the kernel), with a leading
    sample axis so that several samples can share one launch."""

    def __init__(self, variant, Z, ell, var, eps, phase, w, nu, B=None):
        self.variant, self.Z, self.ell, self.var = variant, Z, ell, var
        self.eps, self.phase, self.w, self.nu, self.B = eps, phase, w, nu, B

    @staticmethod
    def stack(samples):
        s0 = samples[0]
        cat = lambda name: torch.cat([getattr(s, name) for s in samples], 0)
        return FieldSample(s0.variant, s0.Z, s0.ell, s0.var, cat("eps"), cat("phase"), cat("w"), cat("nu"),
                           cat("B") if s0.B is not None else None)


class SVGP_Layer(torch.nn.Module):
    def __init__(self, D_in, D_out, M, S, q_diag=False, dimwise=True, device="cpu", kernel="RBF"):
        super().__init__()
        if kernel == "RBF":
            self.kern = RBF(D_in, D_out, dimwise)
            self.dimwise = dimwise
        elif kernel == "DF":
            self.kern = DivergenceFreeKernel(D_in, D_out)
            self.dimwise = False
        else:
            sys.exit("Invalid kernel selection")
        self.kernel_n = kernel
        self.q_diag = q_diag
        self.D_out, self.D_in, self.M, self.S = D_out, D_in, M, S
        self.device = device
        # initial values come from the global numpy RNG in the reference's order (svpy.py:76-86)
        self.inducing_loc = Param(np.random.normal(size=(M, D_in)), name="Inducing locations", device=device)
        self.Um = Param(np.random.normal(size=(M, D_out)) * 1e-1, name="Inducing distribution (mean)", device=device)
        if q_diag:
            self.Us_sqrt = Param(np.ones((M, D_out)) * 1e-3, transform=transforms.SoftPlus(),
                                 name="Inducing distribution (scale)", device=device)
        else:
            self.Us_sqrt = Param(np.stack([np.eye(M)] * D_out) * 1e-3,
                                 transform=transforms.LowerTriangular(M, D_out, device=device),
                                 name="Inducing distribution (scale)", device=device)
        self.kern.to(device)
        self._cache = None
        # torch.linalg.cholesky raises in the reference when K(Z,Z) + jitter is not positive definite (kernels.py:163); the fused
        # setup reports it in a device flag instead.  True: read the flag after every build_cache (one 4-byte D2H, the same host
        # sync torch.linalg.cholesky performs) and raise torch.linalg.LinAlgError; False: no sync, call cholesky_ok() when wanted.
        self.check_cholesky = True
        self.chol_info = None

    def sample_inducing(self):
        """u = Lq eps + m, eps ~ N(0,I) (M,D_out) (whitened inducing sample, svpy.py:88-101)."""
        eps = sample_normal(shape=(self.M, self.D_out)).to(self.device)
        if self.q_diag:
            return self.Us_sqrt() * eps + self.Um()
        return torch.einsum("dnm,md->nd", self.Us_sqrt(), eps) + self.Um()

    def _fused_setup(self):
        """RBF kernels on a CUDA device run the per-rollout setup on the kernels of csrc/setup_kernels.cu (q_diag too: only the
        inducing sample differs, see _inducing_batched); CPU tensors (host-logic unit tests) use torch ops; DF see _df_batched_setup."""
        return self.kernel_n == "RBF" and self.inducing_loc.optvar.is_cuda and self.M <= 512

    def _df_fused_setup(self):
        return (self.kernel_n == "DF" and self.inducing_loc.optvar.is_cuda and self.D_in <= 8
                and self.M <= 512 and self.M * self.D_in <= 4096)

    def _inducing_batched(self, eps_u):
        """u = Lq eps + m for L samples (L,M,D_out).  Full q(u): the packed-parameter kernel (no (D_out,M,M) scatter); q_diag=True: the
        reference's own elementwise form `Us_sqrt() * eps + Um()` (svpy.py:96-97) broadcast over L -- M x D_out elements, nothing to fuse."""
        if self.q_diag:
            return self.Us_sqrt()[None] * eps_u + self.Um()[None]
        return GF.inducing_sample(self.Us_sqrt.optvar, self.Um(), eps_u)

    def _df_batched_setup(self, L):
        """DF kernel on a CUDA device: the L function samples of a rollout share Z, lengthscales and variance, hence ONE
        (M D x M D) Gram matrix and ONE Cholesky factor (the reference rebuilds and refactors it for every sample,
        kernels.py:376-387 called from svpy.py:118-121 inside the serial MC loop); the L right-hand sides are solved together.
        Gram matrix, factorisation (blocked, spread over the chip), both triangular solves and their closed-form backward run on
        the setup kernels of libgpode.so in float64 (csrc/setup_kernels.cu): K(Z,Z) has cond 1e4..1e6 and an
        fp32 factorisation is the largest single error of the reference's own gradients -- here nu is exact to fp32 rounding.
        Inducing sample, prior at Z (CUDA field kernel with nu = 0) and B(omega) are batched over L as well."""
        k = self.kern
        eps, phase, w, eps_u = self._draws(L)
        Z, ell, var = self.inducing_loc(), k.lengthscales, k.variance
        u = self._inducing_batched(eps_u)                                                   # (L,M,D)
        omega = eps / ell.t()[None, :, None, :]                                            # (L,D,S,D)  sample_freq, kernels.py:120-124
        B = k.operator_B_batched(omega)                                                    # (L,S,D,D)
        n = self.M * self.D_out
        nu0 = torch.zeros((L, n, 1), device=Z.device)
        u_prior, _ = GF.gp_field(Z[None].expand(L, -1, -1), Z, nu0, eps, phase, w, ell, var, k.variant, B)   # rff_forward(Z), (L,M,D)
        nu, self.chol_info = GF.compute_nu(Z, ell, var, u_prior, u, k.variant, return_info=True)             # (L, M D, 1)
        if self.check_cholesky:
            self.cholesky_ok(raise_error=True)
        k.rff_omega, k.rff_B, k.nu = omega[-1], B[-1], nu[-1]                               # the last sample stays on the kernel
        return FieldSample(k.variant, Z, ell, var, eps, phase, w, nu, B)

    def _draws(self, L):
        """(eps, phase, w, eps_u) of L function samples, each with a leading axis L"""
        if _rng["mode"] == "device":
            dev, k = self.inducing_loc.optvar.device, self.kern
            shared = self.kernel_n == "RBF" and not k.dimwise
            S2 = 2 * self.S if self.kernel_n == "DF" else self.S
            w = torch.empty((L, S2, self.D_out), device=dev)
            eps = torch.empty((L, self.D_in, self.S) if shared else (L, self.D_in, self.S, self.D_out), device=dev)
            phase = torch.empty((L, 1, self.S) if shared else (L, 1, self.S, self.D_out), device=dev)
            eps_u = torch.empty((L, self.M, self.D_out), device=dev)
            _rng["stream"].fill([w, eps, phase, eps_u], [GF.NORMAL, GF.NORMAL, GF.UNIFORM, GF.NORMAL])
            return eps, phase * (2 * np.pi), w, eps_u
        draws = [self._draw() for _ in range(L)]
        return tuple(torch.stack([d[i] for d in draws]) for i in range(4))

    def _draw(self):
        """host draws of one function sample in the reference's order (kernels.py:126-137, svpy.py:94): w, eps, phase, eps_u."""
        self.kern.build_cache(self.S, self.device)
        eps_u = sample_normal(shape=(self.M, self.D_out)).to(self.device)
        return self.kern.rff_eps, self.kern.rff_phase, self.kern.rff_weights, eps_u

    def build_cache_batched(self, L):
        """L function samples at once (reference: L serial build_cache calls, odegpvae.py:41-43): draws stay in the
        reference's order, the GPU work -- inducing sample, prior at Z, K(Z,Z) + Cholesky + whitened solves -- is one
        batched pass.  Returns a FieldSample with leading axis L."""
        self._cache = None      # a cache left by an earlier build_cache() is stale from here on
        if self._df_fused_setup():
            return self._df_batched_setup(L)
        if not self._fused_setup():
            samples = []
            for _ in range(L):
                self.build_cache()
                samples.append(self.field_sample())
            return FieldSample.stack(samples)
        eps, phase, w, eps_u = self._draws(L)
        k = self.kern
        Z, ell, var = self.inducing_loc(), k.lengthscales, k.variance
        u = self._inducing_batched(eps_u)                                                   # (L,M,D_out)
        nu0 = torch.zeros((L, self.D_out, self.M, 1) if k.dimwise else (L, self.M, self.D_out), device=Z.device)
        u_prior, _ = GF.gp_field(Z[None].expand(L, -1, -1), Z, nu0, eps, phase, w, ell, var, k.variant)   # rff_forward(Z)
        nu, self.chol_info = GF.compute_nu(Z, ell, var, u_prior, u, k.variant, return_info=True)
        if self.check_cholesky:
            self.cholesky_ok(raise_error=True)
        # leave the last sample on the kernel like L serial build_cache calls would
        k.rff_omega = eps[-1] / (ell.t().unsqueeze(1) if k.dimwise else ell.unsqueeze(1))
        k.nu = nu[-1]
        return FieldSample(k.variant, Z, ell, var, eps, phase, w, nu)

    def cholesky_ok(self, raise_error=False):
        """Status of the last fused setup's Cholesky factorisations (host sync).  The reference raises from
        torch.linalg.cholesky (kernels.py:163) -- with raise_error the same exception type is raised here."""
        if self.chol_info is None:
            return True
        bad = self.chol_info.nonzero()
        if bad.numel() == 0:
            return True
        if raise_error:
            k = int(bad[0, 0])
            raise torch.linalg.LinAlgError("gpode_b200 compute_nu: K(Z,Z) + jitter of output dimension %d is not positive definite "
                                           "(leading minor of order %d)" % (k, int(self.chol_info[k])))
        return False

    def build_cache(self):
        """Fix one function sample: feature draws, inducing sample, nu (svpy.py:103-121; same draw order)."""
        if self._fused_setup() or self._df_fused_setup():
            self._cache = self.build_cache_batched(1)
            return
        self._cache = None
        self.kern.build_cache(self.S, self.device)
        u = self.sample_inducing()
        Z = self.inducing_loc()
        Ku = self.kern.K(Z)
        u_prior = self.kern.rff_forward(Z, self.S)
        self.kern.compute_nu(Ku, u_prior, u)

    def field_sample(self):
        """Current cache as a FieldSample with L = 1."""
        k = self.kern
        if k.nu is None:
            raise RuntimeError("build_cache() must run before the layer is evaluated")
        if getattr(self, "_cache", None) is not None:
            return self._cache
        B = k.extra_cache()
        return FieldSample(k.variant, self.inducing_loc(), k.lengthscales, k.variance, k.rff_eps[None], k.rff_phase[None],
                           k.rff_weights[None], k.nu[None], None if B is None else B[None])

    def forward(self, x):
        """f(x) = prior draw + pathwise update for x (N,D_in) -> (N,D_out), on the GPU kernels."""
        s = self.field_sample()
        f, _ = GF.gp_field(x[None], s.Z, s.nu, s.eps, s.phase, s.w, s.ell, s.var, s.variant, s.B)
        return f[0]

    def kl(self):
        """whitened KL(q(u) || N(0,I)) = 1/2 sum_d(-log|Lq_d Lq_d^T| + |m_d|^2 + |Lq_d|_F^2 - M) (svpy.py:144-175)."""
        m = self.Um()
        if not self.q_diag and m.is_cuda:
            return GF.whitened_kl(self.Us_sqrt.optvar, m)     # packed parameter, no (D_out, M, M) scatter
        if self.q_diag:
            diag = self.Us_sqrt()
            trace = diag.square().sum(0)
        else:
            Lq = torch.tril(self.Us_sqrt())
            diag = torch.diagonal(Lq, dim1=1, dim2=2).t()
            trace = Lq.square().sum((1, 2))
        two_kl = -torch.log(diag.square()).sum(0) + m.square().sum(0) + trace - self.M
        return 0.5 * two_kl.sum()
