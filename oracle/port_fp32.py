"""CPU port of the reference's hot path with the reference's OWN operation sequence, for timing.

TEST / BASELINE INFRASTRUCTURE ONLY (see oracle/__init__.py): used by bench.py's ``cpu_baseline`` leg
and ``--impl reference`` arm, and validated against oracle/field.py in tests.  The reference is Python
and cannot travel to the GPU box, so the CPU baseline is this port (``cpu_baseline.kind = "port"``).
Unlike oracle/field.py (written for clarity, direct differences), this file keeps the tensor-op
structure that determines the reference's CPU cost: scaled inputs, squared norms and a batched
matmul for the distances (kernels.py:64-96), einsum feature maps (kernels.py:147-152,340-349),
materialised (D_out,N,M) / (N,M,D,D) temporaries, and autograd through the unrolled fixed-grid solver
(flow.py:76-85 with use_adjoint=False).
"""
import math

import torch

from oracle import solvers


def rbf_sqdist_dimwise(X, X2, ell):
    Xs = X.unsqueeze(0) / ell.unsqueeze(1)                       # (D_out,N,D_in)
    X2s = X2.unsqueeze(0) / ell.unsqueeze(1)                     # (D_out,M,D_in)
    n1 = Xs.pow(2).sum(2)
    n2 = X2s.pow(2).sum(2)
    return -2 * torch.einsum("dnk,dmk->dnm", Xs, X2s) + n1.unsqueeze(-1) + n2.unsqueeze(1)


def rbf_sqdist_shared(X, X2, ell):
    Xs, X2s = X / ell, X2 / ell
    return -2 * Xs @ X2s.t() + Xs.pow(2).sum(1)[:, None] + X2s.pow(2).sum(1)[None, :]


def rbf_K(X, X2, ell, var, dimwise):
    if dimwise:
        return var[:, None, None] * torch.exp(-0.5 * rbf_sqdist_dimwise(X, X2, ell))
    return var * torch.exp(-0.5 * rbf_sqdist_shared(X, X2, ell))


def rbf_field(x, c):
    dimwise = c["variant"] == "rbf_dimwise"
    S = c["w"].shape[0]
    if dimwise:
        xo = torch.einsum("nd,dfk->nfk", x, c["omega"])
        phi = torch.cos(xo + c["phase"]) * torch.sqrt(c["var"] / S)
        f_prior = torch.einsum("nfk,fk->nk", phi, c["w"])
        Kuf = rbf_K(c["Z"], x, c["ell"], c["var"], True)         # (D_out,M,N)
        f_upd = torch.einsum("dm,dmn->nd", c["nu"].squeeze(2), Kuf)
    else:
        xo = x @ c["omega"]
        phi = torch.cos(xo + c["phase"]) * torch.sqrt(c["var"] / S)
        f_prior = phi @ c["w"]
        f_upd = torch.einsum("md,mn->nd", c["nu"], rbf_K(c["Z"], x, c["ell"], c["var"], False))
    return f_prior + f_upd


def df_K(X, X2, ell, var):
    N, D = X.shape
    M = X2.shape[0]
    n1, n2 = X.pow(2).sum(1), X2.pow(2).sum(1)
    sq = -2 * X @ X2.t() + n1[:, None] + n2[None, :]             # (N,M)
    l2 = ell.pow(2)
    rbf = var * torch.exp(-(1 / (2 * l2)) * sq[:, :, None, None])
    diff = X2.t()[:, None, :] - X.t()[:, :, None]                # (D,N,M)
    t1 = (1 / l2) * (diff[:, None] * diff[None]).permute(2, 3, 0, 1)
    t2 = ((D - 1.0) - (1 / l2) * sq[:, :, None, None]) * torch.eye(D, dtype=X.dtype)[None, None]
    K = rbf * (t1 + t2) / l2
    return K.permute(0, 2, 1, 3).reshape(N * D, M * D)


def df_field(x, c):
    om, S, D = c["omega"], c["omega"].shape[1], x.shape[1]
    # the reference rebuilds B(omega) at every call (kernels.py:327-337)
    norm = torch.sqrt(om.pow(2).sum(0))[:, None]
    b = norm * torch.eye(D, dtype=x.dtype)[None] - (om.permute(1, 0, 2) @ om.permute(1, 2, 0)) / norm
    Bm = torch.cat((b, b), 0)
    xo = torch.einsum("nd,dfk->nfk", x, om) + c["phase"]
    phi = torch.cat((torch.cos(xo), torch.sin(xo)), 1).unsqueeze(-1) * Bm.unsqueeze(0)
    phi = phi * torch.sqrt(c["var"] / S)
    f_prior = (phi * c["w"][None, :, :, None]).sum([1, 2])
    Kuf = df_K(c["Z"], x, c["ell"], c["var"])
    return f_prior + torch.einsum("md,mn->nd", c["nu"], Kuf).reshape(x.shape)


def field(x, c):
    return df_field(x, c) if c["variant"] == "df" else rbf_field(x, c)


def rollout(z0, ts, c, order, method):
    def rhs(t, sv):
        if order == 1:
            return field(sv, c)
        q = sv.shape[1] // 2
        return torch.cat([sv[:, q:], field(sv, c)], 1)
    return solvers.odeint(rhs, z0, ts, method=method).permute(1, 0, 2)


def make_cache(variant, D_in, D_out, M, S, ell=2.0, var=1.0, seed=0, requires_grad=True):
    """A random but well-formed function sample (leaf tensors Z, ell, var, nu), fp32 on CPU."""
    g = torch.Generator().manual_seed(seed)
    rn = lambda *s: torch.randn(*s, generator=g)
    dimwise = variant != "rbf_shared"
    c = dict(variant=variant, Z=rn(M, D_in))
    c["ell"] = torch.full((D_out, D_in) if dimwise else (D_in,), float(ell))
    c["var"] = torch.full((D_out,) if dimwise else (1,), float(var))
    for k in ("Z", "ell", "var"):
        c[k].requires_grad_(requires_grad)
    eps = rn(D_in, S, D_out) if dimwise else rn(D_in, S)
    c["eps"] = eps
    c["omega"] = eps / (c["ell"].t().unsqueeze(1) if dimwise else c["ell"].unsqueeze(1))
    c["phase"] = torch.rand((1, S, D_out) if dimwise else (1, S), generator=g) * 2 * math.pi
    c["w"] = rn(2 * S if variant == "df" else S, D_out)
    if variant == "rbf_dimwise":
        c["nu"] = rn(D_out, M, 1)
    elif variant == "rbf_shared":
        c["nu"] = rn(M, D_out)
    else:
        c["nu"] = rn(M * D_out, 1)
    c["nu"].requires_grad_(requires_grad)
    return c


def time_rollout(variant, N, D_in, D_out, M, S, T, order, method, reps=1, warmup=0, threads=None, seed=0):
    """Seconds per forward+backward rollout pass (best of reps) of N trajectories on the host cores."""
    import time
    if threads:
        torch.set_num_threads(threads)
    c = make_cache(variant, D_in, D_out, M, S, seed=seed)
    g = torch.Generator().manual_seed(seed + 1)
    z0 = torch.randn(N, D_in, generator=g).requires_grad_(True)
    ts = 0.1 * torch.arange(T, dtype=torch.float)
    G = torch.randn(N, T, D_in, generator=g)
    best = float("inf")
    dimwise = variant != "rbf_shared"
    for it in range(warmup + reps):
        t0 = time.perf_counter()
        # per-rollout part of build_cache that sits in the autograd graph: omega = eps / ell (kernels.py:120-124)
        c["omega"] = c["eps"] / (c["ell"].t().unsqueeze(1) if dimwise else c["ell"].unsqueeze(1))
        traj = rollout(z0, ts, c, order, method)
        (traj * G).sum().backward()
        dt = time.perf_counter() - t0
        if it >= warmup:
            best = min(best, dt)
        for v in (z0, c["Z"], c["ell"], c["var"], c["nu"]):
            v.grad = None
    return best
