"""Custom autograd Functions over the C ABI: one vector-field evaluation (GPField) and a whole
fixed-step solve (GPRollout).  Tensors carry a leading sample axis L: sample l uses its own
(eps, phase, w, nu[, B]) while Z, ell, var are shared -- one launch covers all MC samples
(reference: serial loop in experiments/model/core/odegpvae.py:41-43).

Gradients are returned for x / z0, Z, nu, ell, var (and B for the DF kernel); ts, eps, phase and w get
None, exactly the tensors that carry requires_grad in the reference graph.
"""
import ctypes

import torch

from . import _lib


def _dims(variant, Z, w, nu, eps):
    M, D_in = Z.shape
    L = w.shape[0]
    D_out = w.shape[2]
    S = eps.shape[2]
    return L, M, D_in, D_out, S


def _c(t):
    return None if t is None else t.contiguous()


class GPField(torch.autograd.Function):
    """f = SVGP_Layer.forward(x) (reference experiments/model/core/svpy.py:123-142) for x (L,N,D_in)."""

    @staticmethod
    def forward(ctx, x, Z, nu, eps, phase, w, ell, var, B, variant):
        lib = _lib.load()
        x, Z, nu, eps, phase, w, ell, var, B = map(_c, (x, Z, nu, eps, phase, w, ell, var, B))
        _lib.require_cuda(x, Z, nu, eps, phase, w, ell, var, B)
        L, M, D_in, D_out, S = _dims(variant, Z, w, nu, eps)
        if x.dim() != 3 or x.shape[0] != L or x.shape[2] != D_in:
            raise RuntimeError("x must be (L=%d, N, D_in=%d), got %s" % (L, D_in, tuple(x.shape)))
        N = x.shape[1]
        with torch.cuda.device(x.device):
            p = _lib.make_problem(variant, L, N, D_in, D_out, M, S, Z, ell, var, eps, phase, w, nu, B)
            ws, nbytes = _lib.workspace(p, 2, _lib.EULER, x.device)
            f = torch.empty((L, N, D_out), dtype=torch.float32, device=x.device)
            fp = torch.empty_like(f)
            rc = lib.gpode_field_fwd(ctypes.byref(p), _lib.ptr(x), _lib.ptr(f), _lib.ptr(fp), _lib.ptr(ws), nbytes,
                                     _lib.stream_handle(x.device))
        _lib.check(rc, "gpode_field_fwd")
        ctx.save_for_backward(x, Z, nu, eps, phase, w, ell, var, B if B is not None else x.new_empty(0), f, fp)
        ctx.variant = variant
        ctx.has_B = B is not None
        ctx.mark_non_differentiable(fp)
        return f, fp

    @staticmethod
    def backward(ctx, g, _gfp):
        lib = _lib.load()
        x, Z, nu, eps, phase, w, ell, var, B, f, fp = ctx.saved_tensors
        B = B if ctx.has_B else None
        g = g.contiguous()
        variant = ctx.variant
        L, M, D_in, D_out, S = _dims(variant, Z, w, nu, eps)
        N = x.shape[1]
        with torch.cuda.device(x.device):
            p = _lib.make_problem(variant, L, N, D_in, D_out, M, S, Z, ell, var, eps, phase, w, nu, B)
            ws, nbytes = _lib.workspace(p, 2, _lib.EULER, x.device)
            dx = torch.empty_like(x)
            dZ, dnu, dell, dvar = torch.empty_like(Z), torch.empty_like(nu), torch.empty_like(ell), torch.empty_like(var)
            dB = torch.empty_like(B) if B is not None else None
            grads = _lib.GpodeParamGrads(_lib.ptr(dZ), _lib.ptr(dell), _lib.ptr(dvar), _lib.ptr(dnu), _lib.ptr(dB))
            rc = lib.gpode_field_bwd(ctypes.byref(p), _lib.ptr(x), _lib.ptr(g), _lib.ptr(f), _lib.ptr(fp), _lib.ptr(dx),
                                     ctypes.byref(grads), _lib.ptr(ws), nbytes, _lib.stream_handle(x.device))
        _lib.check(rc, "gpode_field_bwd")
        return dx, dZ, dnu, None, None, None, dell, dvar, dB, None


class GPRollout(torch.autograd.Function):
    """traj = Flow.forward(z0, ts) for every sample (reference experiments/model/core/flow.py:68-86 +
    torchdiffeq fixed-grid euler / midpoint / rk4): z0 (N,D_s) shared by all samples or (L,N,D_s);
    returns (L,N,T,D_s)."""

    @staticmethod
    def forward(ctx, z0, ts, Z, nu, eps, phase, w, ell, var, B, variant, order, method):
        lib = _lib.load()
        z0, ts, Z, nu, eps, phase, w, ell, var, B = map(_c, (z0, ts, Z, nu, eps, phase, w, ell, var, B))
        _lib.require_cuda(z0, ts, Z, nu, eps, phase, w, ell, var, B)
        L, M, D_in, D_out, S = _dims(variant, Z, w, nu, eps)
        per_sample = z0.dim() == 3
        if per_sample and z0.shape[0] != L:
            raise RuntimeError("per-sample z0 must have leading dim L=%d" % L)
        if z0.shape[-1] != D_in:
            raise RuntimeError("z0 last dim %d != D_in %d" % (z0.shape[-1], D_in))
        N, T = z0.shape[-2], ts.shape[0]
        need_grad = any(ctx.needs_input_grad)
        with torch.cuda.device(z0.device):
            p = _lib.make_problem(variant, L, N, D_in, D_out, M, S, Z, ell, var, eps, phase, w, nu, B)
            ws, nbytes = _lib.workspace(p, T, method, z0.device)
            traj = torch.empty((L, N, T, D_in), dtype=torch.float32, device=z0.device)
            save = None
            if need_grad:
                save = torch.empty(lib.gpode_rollout_save_floats(ctypes.byref(p), T, method), dtype=torch.float32, device=z0.device)
            rc = lib.gpode_rollout_fwd(ctypes.byref(p), _lib.ptr(z0), int(per_sample), _lib.ptr(ts), T, method, order,
                                       _lib.ptr(traj), _lib.ptr(save), _lib.ptr(ws), nbytes, _lib.stream_handle(z0.device))
        _lib.check(rc, "gpode_rollout_fwd")
        ctx.save_for_backward(ts, Z, nu, eps, phase, w, ell, var, B if B is not None else ts.new_empty(0),
                              save if save is not None else ts.new_empty(0))
        ctx.cfg = (variant, order, method, per_sample, N, B is not None)
        return traj

    @staticmethod
    def backward(ctx, dtraj):
        lib = _lib.load()
        ts, Z, nu, eps, phase, w, ell, var, B, save = ctx.saved_tensors
        variant, order, method, per_sample, N, has_B = ctx.cfg
        B = B if has_B else None
        if save.numel() == 0:
            raise RuntimeError("GPRollout.backward called but the forward kept no saves")
        dtraj = dtraj.contiguous()
        L, M, D_in, D_out, S = _dims(variant, Z, w, nu, eps)
        T = ts.shape[0]
        dev = dtraj.device
        with torch.cuda.device(dev):
            p = _lib.make_problem(variant, L, N, D_in, D_out, M, S, Z, ell, var, eps, phase, w, nu, B)
            ws, nbytes = _lib.workspace(p, T, method, dev)
            dz0 = torch.empty((L, N, D_in), dtype=torch.float32, device=dev)
            dZ, dnu, dell, dvar = torch.empty_like(Z), torch.empty_like(nu), torch.empty_like(ell), torch.empty_like(var)
            dB = torch.empty_like(B) if B is not None else None
            grads = _lib.GpodeParamGrads(_lib.ptr(dZ), _lib.ptr(dell), _lib.ptr(dvar), _lib.ptr(dnu), _lib.ptr(dB))
            rc = lib.gpode_rollout_bwd(ctypes.byref(p), _lib.ptr(ts), T, method, order, None, _lib.ptr(save), _lib.ptr(dtraj),
                                       _lib.ptr(dz0), ctypes.byref(grads), _lib.ptr(ws), nbytes, _lib.stream_handle(dev))
        _lib.check(rc, "gpode_rollout_bwd")
        if not per_sample:
            dz0 = dz0.sum(0)
        return dz0, None, dZ, dnu, None, None, None, dell, dvar, dB, None, None, None


def gp_field(x, Z, nu, eps, phase, w, ell, var, variant, B=None):
    """x (L,N,D_in) -> (f, f_prior), both (L,N,D_out)."""
    v = _lib.VARIANTS[variant] if isinstance(variant, str) else variant
    return GPField.apply(x, Z, nu, eps, phase, w, ell, var, B, v)


def gp_rollout(z0, ts, Z, nu, eps, phase, w, ell, var, variant, order=1, method="rk4", B=None):
    """z0 (N,D_s) or (L,N,D_s), ts (T,) -> trajectories (L,N,T,D_s)."""
    v = _lib.VARIANTS[variant] if isinstance(variant, str) else variant
    m = _lib.METHODS[method] if isinstance(method, str) else method
    return GPRollout.apply(z0, ts, Z, nu, eps, phase, w, ell, var, B, v, order, m)
