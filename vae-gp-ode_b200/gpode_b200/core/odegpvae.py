"""ELBO pieces either side of the flow -- drop-ins for the functions of experiments/model/create_model.py (elbo) and
experiments/model/core/odegpvae.py (sample_trajectories, build_decoding) that touch the largest tensors of the model (SURVEY.md
section 8f rank 3).  The conv encoder / decoder, the KL of the encoder and the optimiser stay the
reference's own code.

    elbo(model, X, Xrec, s0_mu, s0_logv, v0_mu, v0_logv, L)     same signature and return value as create_model.py:37-58; the Bernoulli
        log-likelihood ``decoder.log_prob(X, Xrec, L).sum([2,3,4,5]).mean(0)`` (vae.py:136-153) is ONE fused reduction kernel
        (functional.BernoulliLhood): no X.repeat(L), no (L,N,T,1,28,28) intermediates -- 4 B of HBM traffic per element of Xrec
        instead of ~80.
    sample_trajectories(model, z0, T, L)    odegpvae.py:37-45 with all L function samples in one rollout launch.
    position_part(ztL, order)               the decoder input of build_decoding (odegpvae.py:18-35): positions only for order 2.
"""
import torch
from torch.distributions import kl_divergence as kl

from .. import functional as GF


def position_part(ztL, order):
    """what build_decoding feeds the decoder (odegpvae.py:27-34): the whole state for order 1, the position half for order 2"""
    if order == 1:
        return ztL
    return ztL[..., : ztL.shape[-1] // 2]


def sample_trajectories(model, z0, T, L=1):
    """(L,N,T,D_s): L function samples, ONE rollout launch (reference: serial loop over model.flow(z0, ts), odegpvae.py:37-45)"""
    ts = model.dt * torch.arange(T, dtype=torch.float).to(z0.device)
    return model.flow.forward_samples(z0, ts, L)


def elbo(model, X, Xrec, s0_mu, s0_logv, v0_mu, v0_logv, L):
    """(lhood.mean(), kl_reg.mean(), kl_u) exactly as create_model.py:37-58 returns them"""
    q = model.vae.encoder.q_dist(s0_mu, s0_logv, v0_mu, v0_logv)
    kl_reg = kl(q, model.vae.prior).sum(-1)                       # (N,)
    if model.vae.decoder.distribution != "bernoulli":
        raise ValueError("Currently only bernoulli dist implemented")
    lhood = GF.bernoulli_lhood(X, Xrec)                           # (N,) = log_prob(X, Xrec, L).sum([2,3,4,5]).mean(0)
    kl_u = model.flow.kl()
    return lhood.mean(), kl_reg.mean(), kl_u
