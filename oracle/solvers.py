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
