"""Drive the reference's FULL model (experiments/model/create_model.py:9-35 build_model, :61-73 compute_loss) on synthetic rotating
digits with every random draw pinned.  TEST INFRASTRUCTURE ONLY (see oracle/__init__.py).

Used twice: by oracle/make_golden_elbo.py with the reference's own SVGP_Layer / Flow (CPU, fp32 and fp64) to freeze
tests/golden/elbo_*.npz, and by tests/test_gpu_elbo.py with the drop-in classes swapped into ``create_model`` exactly as
INTEGRATION.md section 1 prescribes -- same ``build_model(args)``, same ``ODEGPVAE`` / ``VAE`` / ``elbo`` code, same inputs, same draws.

Random sources of one ``compute_loss`` call, all replaced by recorded / replayed arrays:
  * GP function draws: kernels.sample_normal / sample_uniform, svpy.sample_normal (reference_harness.DrawRecorder) -- per MC sample
    w, eps, phase, eps_u in this order (kernels.py:126-137 / 305-316, svpy.py:94);
  * the encoders' reparameterisation noise: Encoder.sample (vae.py:83-86) draws ``torch.randn_like``; replaced by a method with the
    same arithmetic that takes its noise from the replay list (position encoder first, then the velocity encoder, odegpvae.py:57-64).
"""
import argparse

import numpy as np
import torch

from oracle import glyph

# BASELINE.json configs 1-3: M = 100, S = 256, ell = 2.0, var = 1.0 (README.md:28), dt = 0.1, n_filt = 8,
# Ndata = 360 (main.py:37).  "solver" is varied by the callers (euler = main.py default, rk4 = north_star).
CONFIGS = {
    "cfg1": dict(kernel="RBF", ode=1, latent_dim=6, D_in=6, D_out=6, N=25, L=1, T=16, frames=5),
    "cfg2": dict(kernel="DF", ode=1, latent_dim=6, D_in=6, D_out=6, N=256, L=4, T=16, frames=5),
    "cfg3": dict(kernel="RBF", ode=2, latent_dim=3, D_in=6, D_out=3, N=25, L=1, T=16, frames=5),
}


def make_args(cfg, solver, device):
    c = CONFIGS[cfg]
    return argparse.Namespace(D_in=c["D_in"], D_out=c["D_out"], num_inducing=100, num_features=256, dimwise=True, q_diag=False,
                              device=device, kernel=c["kernel"], ode=c["ode"], solver=solver, use_adjoint=False, frames=c["frames"],
                              n_filt=8, latent_dim=c["latent_dim"], Ndata=360, dt=0.1, lengthscale=2.0, variance=1.0)


class Replay:
    """feeds recorded arrays, in order, to whoever asks for a draw of that shape"""

    def __init__(self, arrays, dtype=torch.float32, device="cpu"):
        self.arrays, self.i, self.dtype, self.device = list(arrays), 0, dtype, device

    def __call__(self, shape, seed=None):
        a = self.arrays[self.i]
        self.i += 1
        assert tuple(a.shape) == tuple(shape), (self.i - 1, a.shape, tuple(shape))
        return torch.tensor(a, dtype=self.dtype)        # host tensor, like the reference helpers; the callers move it

    def left(self):
        return len(self.arrays) - self.i


def patch_encoder_noise(vae_module, noise):
    """Encoder.sample (vae.py:83-86) with the noise taken from ``noise`` (a Replay or a recording callable)"""
    def sample(self, mu, logvar):
        std = torch.exp(0.5 * logvar)
        eps = noise(tuple(std.shape)).to(device=std.device, dtype=std.dtype)
        return mu + std * eps
    vae_module.Encoder.sample = sample


def init_model(create_model, initialization, args, seed):
    """main.py:141-153: seed, build_model, .to(device), initialize_and_fix_kernel_parameters"""
    np.random.seed(seed)
    torch.manual_seed(seed)
    model = create_model.build_model(args)
    model.to(args.device)
    model = initialization.initialize_and_fix_kernel_parameters(model, lengthscale_value=args.lengthscale, variance_value=args.variance, fix=False)
    model.train()
    return model


def inputs(cfg, seed=121):
    c = CONFIGS[cfg]
    return glyph.rotating_sequences(c["N"], c["T"], seed)       # (N,T,1,28,28) float32, normalised


def run_loss(create_model, model, X, L):
    """compute_loss + backward (main.py:204-210); returns the four scalars and the gradient of every named parameter"""
    for p in model.parameters():
        p.grad = None
    loss, nlhood, kl_reg, kl_gp = create_model.compute_loss(model, X, L)
    loss.backward()
    scal = dict(loss=loss.item(), nlhood=nlhood.item(), kl_reg=kl_reg.item(), kl_gp=kl_gp.item())
    grads = {n: (p.grad.detach().cpu().double().numpy() if p.grad is not None else np.zeros(tuple(p.shape))) for n, p in model.named_parameters()}
    return scal, grads
