"""ODEfunc / Flow -- drop-in for experiments/model/core/flow.py:7-102.

``Flow.forward(z0, ts)`` keeps the reference contract (fresh function sample per call, returns
(N,T,D_s)) but the whole fixed-grid solve -- every stage of every step, plus its reverse sweep -- is
one CUDA launch per direction instead of a torchdiffeq Python loop.  ``Flow.forward_samples`` is the
batched form of the serial MC loop in experiments/model/core/odegpvae.py:37-45: L function samples,
one launch.  Fixed-grid solvers only (euler, midpoint, rk4 = 3/8 rule); adaptive solvers raise.
``use_adjoint=True`` (flow.py:76, odeint_adjoint) keeps the fused forward launch and computes the gradients
by torchdiffeq's adjoint method on the field kernels (functional.GPRolloutAdjoint).  ``ts_dense_scale``
(main.py:83, misc/torch_utils.py:54-61) integrates on the densified grid and returns the states at ``ts``."""
import torch
import torch.nn as nn

from .. import functional as GF
from .._lib import METHODS, STAGES
from ..misc.torch_utils import compute_ts_dense
from .svpy import FieldSample


class ODEfunc(nn.Module):
    def __init__(self, diffeq, order):
        super().__init__()
        self.diffeq = diffeq
        self.order = order
        self.register_buffer("_num_evals", torch.tensor(0.))

    def before_odeint(self, rebuild_cache):
        self._num_evals.fill_(0)
        if rebuild_cache:
            self.diffeq.build_cache()

    def num_evals(self):
        return self._num_evals.item()

    def first_order(self, sv):
        return self.diffeq(sv)

    def second_order(self, sv):
        q = sv.shape[1] // 2
        return torch.cat([sv[:, q:], self.diffeq(sv)], 1)

    def forward(self, t, sv):
        """torchdiffeq-style right-hand side (autonomous: t is ignored), one field evaluation on the GPU."""
        self._num_evals += 1
        if self.order == 1:
            return self.first_order(sv)
        if self.order == 2:
            return self.second_order(sv)
        raise ValueError("order must be 1 or 2")


class Flow(nn.Module):
    def __init__(self, diffeq, order=2, solver="dopri5", atol=1e-6, rtol=1e-6, use_adjoint=False):
        super().__init__()
        self.odefunc = ODEfunc(diffeq, order)
        self.solver = solver
        self.atol = atol
        self.rtol = rtol
        self.use_adjoint = use_adjoint

    def _method(self):
        if self.solver not in METHODS:
            raise NotImplementedError("gpode_b200 implements the fixed-grid solvers %s; got %r (adaptive solvers are "
                                      "outside the CUDA hot path)" % (sorted(METHODS), self.solver))
        return METHODS[self.solver]

    def _rollout(self, z0, ts, sample):
        method = self._method()
        scale = int(getattr(self, "ts_dense_scale", 1) or 1)
        grid = compute_ts_dense(ts, scale) if scale > 2 else ts            # (scale = 2 adds no point: linspace(t1, t2, 2)[:-1] = [t1])
        traj = GF.gp_rollout(z0, grid, sample.Z, sample.nu, sample.eps, sample.phase, sample.w, sample.ell, sample.var,
                             sample.variant, self.odefunc.order, method, sample.B, adjoint=self.use_adjoint)
        self.odefunc._num_evals.fill_((grid.shape[0] - 1) * STAGES[method])
        if grid is not ts:
            traj = traj[:, :, :: scale - 1]                                # the states at the requested time points
        return traj

    def forward(self, z0, ts):
        """(N,D_s), (T,) -> (N,T,D_s) for one fresh function sample."""
        self.odefunc.before_odeint(rebuild_cache=True)
        gp = self.odefunc.diffeq
        sample = gp.field_sample()
        gp._cache = None    # consumed: later SVGP_Layer.forward calls read the sample from the kernel attributes, and the
        return self._rollout(z0, ts, sample)[0]   # layer does not keep this iteration's autograd graph alive

    def forward_samples(self, z0, ts, L):
        """(N,D_s) -> (L,N,T,D_s): L fresh function samples (caches drawn in order), a single launch."""
        self.odefunc.before_odeint(rebuild_cache=False)
        return self._rollout(z0, ts, self.odefunc.diffeq.build_cache_batched(L))

    def num_evals(self):
        return self.odefunc.num_evals()

    def kl(self):
        return self.odefunc.diffeq.kl()


def sample_trajectories(flow, z0, T, L=1, dt=0.1):
    """Batched equivalent of ODEGPVAE.sample_trajectories (odegpvae.py:37-45): (L,N,T,D_s)."""
    ts = dt * torch.arange(T, dtype=torch.float).to(z0.device)
    return flow.forward_samples(z0, ts, L)
