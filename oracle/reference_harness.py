"""Drive the LIVE reference modules (read-only under /root/reference) with reproducible draws.

TEST INFRASTRUCTURE ONLY (see oracle/__init__.py).  Usable only where /root/reference exists (the
build container); the GPU box consumes the frozen vectors under tests/golden/ instead.

What it does
  * puts ``/root/reference/experiments`` on sys.path so ``model.core.*`` imports unmodified;
  * injects stub modules for the three absent third-party imports: ``torchsummary`` (only used by
    ``vae.py:25-30`` print_summary), ``matplotlib`` (plots) and ``torchdiffeq`` -- the latter mapped
    to ``oracle.solvers.odeint`` which restates torchdiffeq's fixed-grid solvers (flow.py:3-4,76-85);
  * replaces the three host RNG helpers (kernels.py:13-26, svpy.py:12-27) by a recording, seeded
    source: ``kernels.sample_normal(seed=None)`` otherwise builds a fresh unseeded RandomState
    (kernels.py:17) and no run would be reproducible.
"""
import os
import sys
import types

import numpy as np
import torch

_HERE = os.path.dirname(os.path.abspath(__file__))
# the live tree in the build container; elsewhere (GPU box) the verbatim copy made by oracle/fetch_ref.py (git-ignored oracle/_ref)
REFERENCE_ROOT = os.environ.get("GPODE_REFERENCE_ROOT") or ("/root/reference" if os.path.isdir("/root/reference/experiments/model/core")
                                                             else os.path.join(_HERE, "_ref"))


def reference_available():
    return os.path.isdir(os.path.join(REFERENCE_ROOT, "experiments", "model", "core"))


class DrawRecorder:
    """Seeded replacement for the reference's numpy draw helpers; keeps every draw in order."""

    def __init__(self, seed):
        self.rng = np.random.RandomState(seed)
        self.log = []

    def normal(self, shape, seed=None):
        v = self.rng.normal(size=shape).astype(np.float32)
        self.log.append(("normal", v))
        return torch.tensor(v)

    def uniform(self, shape, seed=None):
        v = self.rng.uniform(low=0.0, high=1.0, size=shape).astype(np.float32)
        self.log.append(("uniform", v))
        return torch.tensor(v)


_loaded = {}


def load_reference():
    """Import the reference packages once (with stubs) and return them in a dict."""
    if _loaded:
        return _loaded
    if not reference_available():
        raise RuntimeError("reference tree not present at %s" % REFERENCE_ROOT)
    from oracle import solvers

    if "torchsummary" not in sys.modules:
        m = types.ModuleType("torchsummary")
        m.summary = lambda *a, **k: None
        sys.modules["torchsummary"] = m
    if "torchdiffeq" not in sys.modules:
        m = types.ModuleType("torchdiffeq")
        m.odeint = solvers.odeint
        m.odeint_adjoint = solvers.odeint_adjoint
        sys.modules["torchdiffeq"] = m
    if "matplotlib" not in sys.modules:
        try:
            import matplotlib  # noqa: F401
        except Exception:
            mp = types.ModuleType("matplotlib")
            mp.pyplot = types.ModuleType("matplotlib.pyplot")
            sys.modules["matplotlib"] = mp
            sys.modules["matplotlib.pyplot"] = mp.pyplot
    exp = os.path.join(REFERENCE_ROOT, "experiments")
    if exp not in sys.path:
        sys.path.insert(0, exp)
    # the reference puts Param tensors on cuda:0 whenever CUDA is visible (param.py:20-22)
    import model.core.kernels as kernels
    import model.core.svpy as svpy
    import model.core.flow as flow
    import model.core.odegpvae as odegpvae
    import model.create_model as create_model
    import model.core.initialization as initialization
    import model.core.vae as vae

    _loaded.update(kernels=kernels, svpy=svpy, flow=flow, odegpvae=odegpvae,
                   create_model=create_model, initialization=initialization, vae=vae)
    return _loaded


def seed_draws(seed):
    """Patch the reference RNG helpers with one seeded recorder; returns the recorder."""
    ref = load_reference()
    rec = DrawRecorder(seed)
    ref["kernels"].sample_normal = rec.normal
    ref["kernels"].sample_uniform = rec.uniform
    ref["svpy"].sample_normal = rec.normal
    ref["svpy"].sample_uniform = rec.uniform
    return rec


def make_layer(D_in, D_out, M, S, kernel="RBF", dimwise=True, q_diag=False, ell=2.0, var=1.0,
               init_seed=0, perturb=False):
    """Build a reference SVGP_Layer on CPU with seeded numpy init (svpy.py:76-86 use np.random)."""
    ref = load_reference()
    np.random.seed(init_seed)
    gp = ref["svpy"].SVGP_Layer(D_in=D_in, D_out=D_out, M=M, S=S, q_diag=q_diag, dimwise=dimwise,
                                device="cpu", kernel=kernel)
    from model.misc.constraint_utils import invsoftplus
    k = gp.kern
    rs = np.random.RandomState(init_seed + 1000)
    lval = ell * torch.ones_like(k.unconstrained_lengthscales.data)
    vval = var * torch.ones_like(k.unconstrained_variance.data)
    if perturb:
        lval = lval + torch.tensor(rs.uniform(size=tuple(lval.shape)).astype(np.float32))
        vval = vval + torch.tensor(rs.uniform(size=tuple(vval.shape)).astype(np.float32))
    k.unconstrained_lengthscales.data = invsoftplus(lval)
    k.unconstrained_variance.data = invsoftplus(vval)
    if not q_diag:
        # give q(u) a non-trivial Cholesky factor so that Lq gradients are exercised
        with torch.no_grad():
            gp.Us_sqrt.optvar.add_(0.02 * torch.tensor(rs.normal(size=tuple(gp.Us_sqrt.optvar.shape)).astype(np.float32)))
    return gp


def extract_cache(gp):
    """Tensors that define the current function sample (after build_cache)."""
    k = gp.kern
    return dict(Z=gp.inducing_loc().detach().clone(), nu=k.nu.detach().clone(),
                omega=k.rff_omega.detach().clone(), phase=k.rff_phase.detach().clone(),
                w=k.rff_weights.detach().clone(), ell=k.lengthscales.detach().clone(),
                var=k.variance.detach().clone())
