"""Time the REAL reference's CPU implementation of the hot path.  BASELINE INFRASTRUCTURE ONLY (see oracle/__init__.py).

Used by bench.py's ``cpu_baseline`` leg and ``--impl reference`` arm (``kind = "reference"``): the reference's own ``SVGP_Layer`` and
``Flow`` / ``ODEfunc`` (experiments/model/core/{svpy,kernels,flow}.py, unmodified -- from /root/reference in the build container, from the
verbatim copy oracle/_ref on the GPU box) driven exactly like ``ODEGPVAE.sample_trajectories`` drives them (core/odegpvae.py:37-45), with
``torchdiffeq.odeint`` replaced by the restated fixed-grid loop of oracle/solvers.py (torchdiffeq is absent from the image).
Timed region per pass = ``build_cache`` (draws, K(Z,Z), Cholesky, nu) + rollout + ``backward`` -- BASELINE.md section 3.
Run it in a process that cannot see a GPU (CUDA_VISIBLE_DEVICES=""): the reference's Param puts parameters on cuda:0 whenever CUDA is
visible (misc/param.py:20-22).
"""
import time

import numpy as np
import torch

from oracle import reference_harness as rh


def time_reference(variant, N, D_in, D_out, M, S, T, order, method, reps=1, warmup=0, threads=None, seed=0, ell=2.0, var=1.0):
    """seconds per forward+backward rollout pass (best of reps) of N trajectories, one function sample, on the host cores"""
    if threads:
        torch.set_num_threads(threads)
    ref = rh.load_reference()
    kernel = "DF" if variant == "df" else "RBF"
    gp = rh.make_layer(D_in, D_out, M, S, kernel=kernel, dimwise=variant != "rbf_shared", ell=ell, var=var, init_seed=seed)
    flow = ref["flow"].Flow(diffeq=gp, order=order, solver=method, use_adjoint=False)
    rs = np.random.RandomState(seed + 1)
    z0 = torch.tensor(rs.normal(size=(N, D_in)).astype(np.float32), requires_grad=True)
    G = torch.tensor(rs.normal(size=(N, T, D_in)).astype(np.float32))
    ts = 0.1 * torch.arange(T, dtype=torch.float)
    params = [p for p in flow.parameters()]
    best = float("inf")
    for it in range(warmup + reps):
        rh.seed_draws(seed + 2 + it)
        t0 = time.perf_counter()
        traj = flow(z0, ts)                       # build_cache + fixed-grid solve (flow.py:68-86)
        (traj * G).sum().backward()
        dt = time.perf_counter() - t0
        if it >= warmup:
            best = min(best, dt)
        z0.grad = None
        for p in params:
            p.grad = None
    return best
