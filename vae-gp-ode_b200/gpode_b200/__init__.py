"""gpode_b200 -- B200-native (sm_100a) drop-in for the GP-ODE hot path of IlzeAmandaA/VAE-GP-ODE.

Mirrors the reference's module layout for this path only:
    gpode_b200.core.kernels   RBF, DivergenceFreeKernel      (reference: experiments/model/core/kernels.py)
    gpode_b200.core.svpy      SVGP_Layer                      (reference: experiments/model/core/svpy.py)
    gpode_b200.core.flow      ODEfunc, Flow                   (reference: experiments/model/core/flow.py)
    gpode_b200.misc.*         Param, transforms, softplus     (reference: experiments/model/misc/*)
    gpode_b200.core.odegpvae  elbo with the fused Bernoulli log-likelihood (reference: experiments/model/create_model.py:37-58)
    gpode_b200.functional     GPField / GPRollout / ComputeNu / InducingSample / WhitenedKL / BernoulliLhood autograd Functions
                              and the PhiloxStream draw source over the C ABI (include/gpode.h)

There is no CPU path and no fallback: every evaluation goes through libgpode.so on a CUDA device and
raises if the library or a GPU is missing.
"""
from . import _lib  # noqa: F401
from ._lib import kernel_flags, FLAG_FWD_MMA, FLAG_FWD_TCGEN05, FLAG_BWD_MMA, FLAG_BWD_TCGEN05  # noqa: F401
from .functional import (gp_field, gp_rollout, GPField, GPRollout, compute_nu, inducing_sample, whitened_kl,  # noqa: F401
                         ComputeNu, InducingSample, WhitenedKL, PhiloxStream, BernoulliLhood, bernoulli_lhood)
from .core.svpy import set_rng  # noqa: F401

__all__ = ["gp_field", "gp_rollout", "GPField", "GPRollout", "compute_nu", "inducing_sample", "whitened_kl", "ComputeNu",
           "InducingSample", "WhitenedKL", "PhiloxStream", "BernoulliLhood", "bernoulli_lhood", "set_rng"]
