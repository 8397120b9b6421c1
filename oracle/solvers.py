"""Fixed-grid ODE solvers restated from torchdiffeq's published algorithm.  TEST INFRASTRUCTURE ONLY.

The reference calls ``torchdiffeq.odeint(func, y0, t, rtol, atol, method)`` with no ``options``
(experiments/model/core/flow.py:76-85).  torchdiffeq is third-party, un-vendored and un-pinned (the
reference has no requirements file; README.md:16 names python 3.8 / PyTorch 1.13, contemporaneous
release 0.2.3) and is absent from this image.  Published algorithm followed here
(torchdiffeq/_impl/solvers.py FixedGridODESolver.integrate, fixed_grid.py Euler/Midpoint/RK4,
rk_common.py rk4_alt_step_func):
  * with no ``step_size`` option the integration grid IS the output grid ``t``; one step per interval,
    ``dt = t[i+1] - t[i]`` computed in the dtype of ``t``; the outputs are the grid states;
  * euler:    y1 = y0 + dt * f(t0, y0)
  * midpoint: y1 = y0 + dt * f(t0 + dt/2, y0 + dt/2 * f(t0, y0))
  * rk4:      the 3/8 rule -- k1 = f(t0,y0); k2 = f(t0+dt/3, y0 + dt*k1/3);
              k3 = f(t0+2dt/3, y0 + dt*(k2 - k1/3)); k4 = f(t1, y0 + dt*(k1 - k2 + k3));
              y1 = y0 + dt * (k1 + 3*(k2 + k3) + k4) / 8
``atol``/``rtol`` are ignored by fixed-grid solvers.  No reference test pins this boundary: parity
unpinned (oracle/__init__.py).
"""
import torch

FIXED_GRID_METHODS = ("euler", "midpoint", "rk4")
STAGES = {"euler": 1, "midpoint": 2, "rk4": 4}


def step(func, t0, dt, y0, method):
    if method == "euler":
        return y0 + dt * func(t0, y0)
    if method == "midpoint":
        half = 0.5 * dt
        return y0 + dt * func(t0 + half, y0 + half * func(t0, y0))
    if method == "rk4":
        third = 1.0 / 3.0
        k1 = func(t0, y0)
        k2 = func(t0 + dt * third, y0 + dt * k1 * third)
        k3 = func(t0 + dt * 2.0 * third, y0 + dt * (k2 - k1 * third))
        k4 = func(t0 + dt, y0 + dt * (k1 - k2 + k3))
        return y0 + (k1 + 3.0 * (k2 + k3) + k4) * dt * 0.125
    raise ValueError("oracle restates fixed-grid solvers only (euler, midpoint, rk4); got %r" % (method,))


def odeint(func, y0, t, rtol=None, atol=None, method="rk4", options=None, **unused):
    """Returns (T, *y0.shape): the state at every grid point of ``t`` (t[0] -> y0)."""
    ys = [y0]
    y = y0
    for i in range(t.shape[0] - 1):
        t0, t1 = t[i], t[i + 1]
        y = step(func, t0, t1 - t0, y, method)
        ys.append(y)
    return torch.stack(ys, 0)


def _axpy(y, h, ks, cs):
    """y + h * sum_j cs[j] * ks[j] on tuples of tensors"""
    return tuple(yi + h * sum(c * k[i] for c, k in zip(cs, ks)) for i, yi in enumerate(y))


def _step_tuple(F, t0, dt, y0, method):
    """one fixed-grid step on a tuple state (torchdiffeq flattens the tuple into one vector; componentwise is the same arithmetic)"""
    if method == "euler":
        return _axpy(y0, dt, [F(t0, y0)], [1.0])
    if method == "midpoint":
        k1 = F(t0, y0)
        return _axpy(y0, dt, [F(t0 + 0.5 * dt, _axpy(y0, dt, [k1], [0.5]))], [1.0])
    if method == "rk4":
        third = 1.0 / 3.0
        k1 = F(t0, y0)
        k2 = F(t0 + dt * third, _axpy(y0, dt, [k1], [third]))
        k3 = F(t0 + dt * 2.0 * third, _axpy(y0, dt, [k2, k1], [1.0, -third]))
        k4 = F(t0 + dt, _axpy(y0, dt, [k1, k2, k3], [1.0, -1.0, 1.0]))
        return _axpy(y0, dt, [k1, k2, k3, k4], [0.125, 0.375, 0.375, 0.125])
    raise ValueError(method)


class _OdeintAdjoint(torch.autograd.Function):
    """torchdiffeq/_impl/adjoint.py OdeintAdjointMethod restated for fixed-grid methods without options (the reference's call,
    core/flow.py:76-85): forward = odeint under no_grad, keeping only the solution at the grid points; backward = for i = T-1 .. 1
    one step of the SAME method from t[i] to t[i-1] on the augmented state (y, adj_y, adj_params) with dynamics
    (f, -vjp_y, -vjp_params), y reset to the stored y[i-1] and dL/dy[i-1] added to adj_y after every interval.  t carries no gradient."""

    @staticmethod
    def forward(ctx, func, method, y0, t, *params):
        with torch.no_grad():
            ys = odeint(func, y0, t, method=method)
        ctx.func, ctx.method = func, method
        ctx.save_for_backward(t, ys, *params)
        return ys

    @staticmethod
    def backward(ctx, grad_y):
        t, ys, *params = ctx.saved_tensors
        func, method = ctx.func, ctx.method
        params = tuple(params)

        def aug(tt, state):
            y, adj_y = state[0], state[1]
            with torch.enable_grad():
                y = y.detach().requires_grad_(True)
                fe = func(tt, y)
                vjps = torch.autograd.grad(fe, (y,) + params, -adj_y, allow_unused=True)
            vjps = [torch.zeros_like(v) if g is None else g for g, v in zip(vjps, (y,) + params)]
            return (fe.detach(), vjps[0]) + tuple(vjps[1:])

        state = (ys[-1], grad_y[-1]) + tuple(torch.zeros_like(p) for p in params)
        for i in range(t.shape[0] - 1, 0, -1):
            state = _step_tuple(aug, t[i], t[i - 1] - t[i], state, method)
            state = (ys[i - 1], state[1] + grad_y[i - 1]) + tuple(state[2:])
        return (None, None, state[1], None) + tuple(state[2:])


def odeint_adjoint(func, y0, t, rtol=None, atol=None, method="rk4", options=None, adjoint_params=None, **unused):
    """torchdiffeq.odeint_adjoint for the fixed-grid methods; adjoint_params defaults to func.parameters() like torchdiffeq"""
    if adjoint_params is None:
        adjoint_params = tuple(p for p in func.parameters() if p.requires_grad)
    return _OdeintAdjoint.apply(func, method, y0, t, *tuple(adjoint_params))
