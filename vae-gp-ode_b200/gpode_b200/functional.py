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
        ctx.flags = p.flags
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
            p = _lib.make_problem(variant, L, N, D_in, D_out, M, S, Z, ell, var, eps, phase, w, nu, B, flags=ctx.flags)
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
        ctx.flags = p.flags
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
            p = _lib.make_problem(variant, L, N, D_in, D_out, M, S, Z, ell, var, eps, phase, w, nu, B, flags=ctx.flags)
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


class GPRolloutAdjoint(torch.autograd.Function):
    """Flow.forward with ``use_adjoint=True`` (reference core/flow.py:76: torchdiffeq.odeint_adjoint): the forward is the same fused
    rollout launch and keeps NO stage saves; the backward integrates torchdiffeq's augmented system
        d/dt (z, a, g_theta) = (f(z), -J(z)^T a, -(df/dtheta)(z)^T a)
    backwards over every grid interval with ONE step of the same fixed-grid method, restarting z from the stored trajectory and
    adding dL/dz_i to a at every grid point (torchdiffeq/_impl/adjoint.py OdeintAdjointMethod.backward, restated in
    oracle/solvers.py odeint_adjoint).  Each stage is one field evaluation + one field VJP on the CUDA kernels (gpode_field_fwd /
    gpode_field_bwd); the stage combinations are a handful of small tensor ops.  Memory O(T) states instead of O(T stages) saves;
    gradients are those of the continuous adjoint discretised by the solver (they differ from the default mode's exact gradient
    of the discrete solve at O(dt^p), exactly as in the reference)."""

    @staticmethod
    def forward(ctx, z0, ts, Z, nu, eps, phase, w, ell, var, B, variant, order, method):
        with torch.no_grad():
            traj = GPRollout.apply(z0.detach(), ts, Z.detach(), nu.detach(), eps, phase, w, ell.detach(), var.detach(),
                                   None if B is None else B.detach(), variant, order, method)
        ctx.save_for_backward(ts, Z, nu, eps, phase, w, ell, var, B if B is not None else ts.new_empty(0), traj)
        ctx.cfg = (variant, order, method, z0.dim() == 3, B is not None)
        return traj

    @staticmethod
    def backward(ctx, dtraj):
        lib = _lib.load()
        ts, Z, nu, eps, phase, w, ell, var, B, traj = ctx.saved_tensors
        variant, order, method, per_sample, has_B = ctx.cfg
        B = B if has_B else None
        L, M, D_in, D_out, S = _dims(variant, Z, w, nu, eps)
        N, T = traj.shape[1], traj.shape[2]
        dev = traj.device
        dtraj = dtraj.contiguous()
        tsh = ts.detach().cpu().tolist()
        q = D_out
        with torch.cuda.device(dev):
            p = _lib.make_problem(variant, L, N, D_in, D_out, M, S, Z, ell, var, eps, phase, w, nu, B)
            ws, nbytes = _lib.workspace(p, 2, _lib.EULER, dev)
            stream = _lib.stream_handle(dev)
            f, fp = torch.empty((L, N, D_out), device=dev), torch.empty((L, N, D_out), device=dev)
            names = ["Z", "ell", "var", "nu"] + (["B"] if has_B else [])
            like = dict(Z=Z, ell=ell, var=var, nu=nu, B=B)
            acc = {k: torch.zeros_like(like[k]) for k in names}
            buf = {k: torch.empty_like(like[k]) for k in names}
            grads = _lib.GpodeParamGrads(_lib.ptr(buf["Z"]), _lib.ptr(buf["ell"]), _lib.ptr(buf["var"]), _lib.ptr(buf["nu"]),
                                         _lib.ptr(buf["B"]) if has_B else None)
            dx = torch.empty((L, N, D_in), device=dev)

            def F(z, a, weight):
                """(dz/dt, da/dt) at (z, a); adds weight * d g_theta / dt to acc"""
                z = z.contiguous()
                _lib.check(lib.gpode_field_fwd(ctypes.byref(p), _lib.ptr(z), _lib.ptr(f), _lib.ptr(fp), _lib.ptr(ws), nbytes, stream), "gpode_field_fwd")
                g = a if order == 1 else a[..., q:]
                g = g.contiguous()
                _lib.check(lib.gpode_field_bwd(ctypes.byref(p), _lib.ptr(z), _lib.ptr(g), _lib.ptr(f), _lib.ptr(fp), _lib.ptr(dx), ctypes.byref(grads),
                                               _lib.ptr(ws), nbytes, stream), "gpode_field_bwd")
                for k in names:
                    acc[k].add_(buf[k], alpha=-weight)
                if order == 1:
                    return f.clone(), -dx
                fz = torch.cat([z[..., q:], f], -1)
                ka = -dx
                ka[..., q:] -= a[..., :q]            # d/dv of the position rows dz_s/dt = v
                return fz, ka

            a = dtraj[:, :, T - 1].clone()
            for i in range(T - 1, 0, -1):
                h = tsh[i - 1] - tsh[i]
                z = traj[:, :, i]
                if method == _lib.EULER:
                    k1z, k1a = F(z, a, h)
                    a = a + h * k1a
                elif method == _lib.MIDPOINT:
                    k1z, k1a = F(z, a, 0.0)
                    _, k2a = F(z + 0.5 * h * k1z, a + 0.5 * h * k1a, h)
                    a = a + h * k2a
                else:   # rk4 = 3/8 rule (torchdiffeq rk4_alt_step_func)
                    k1z, k1a = F(z, a, h / 8)
                    k2z, k2a = F(z + h * k1z / 3, a + h * k1a / 3, 3 * h / 8)
                    k3z, k3a = F(z + h * (k2z - k1z / 3), a + h * (k2a - k1a / 3), 3 * h / 8)
                    _, k4a = F(z + h * (k1z - k2z + k3z), a + h * (k1a - k2a + k3a), h / 8)
                    a = a + h * (k1a + 3 * (k2a + k3a) + k4a) / 8
                a = a + dtraj[:, :, i - 1]
        dz0 = a if per_sample else a.sum(0)
        return dz0, None, acc["Z"], acc["nu"], None, None, None, acc["ell"], acc["var"], acc.get("B"), None, None, None


def gp_field(x, Z, nu, eps, phase, w, ell, var, variant, B=None):
    """x (L,N,D_in) -> (f, f_prior), both (L,N,D_out)."""
    v = _lib.VARIANTS[variant] if isinstance(variant, str) else variant
    return GPField.apply(x, Z, nu, eps, phase, w, ell, var, B, v)


def gp_rollout(z0, ts, Z, nu, eps, phase, w, ell, var, variant, order=1, method="rk4", B=None, adjoint=False):
    """z0 (N,D_s) or (L,N,D_s), ts (T,) -> trajectories (L,N,T,D_s).  adjoint: gradients by torchdiffeq's adjoint method (GPRolloutAdjoint)
    instead of the exact reverse sweep through the discrete solve."""
    v = _lib.VARIANTS[variant] if isinstance(variant, str) else variant
    m = _lib.METHODS[method] if isinstance(method, str) else method
    fn = GPRolloutAdjoint if adjoint else GPRollout
    return fn.apply(z0, ts, Z, nu, eps, phase, w, ell, var, B, v, order, m)


# ------------------------------------------------------------------------------------------------
# per-rollout setup at the inducing points (csrc/setup_kernels.cu)
# ------------------------------------------------------------------------------------------------
class ComputeNu(torch.autograd.Function):
    """nu = Lc^-T (u - Lc^-1 u_prior), Lc = chol(K(Z,Z) + 1e-5 I) for L samples at once -- RBF.compute_nu fused with K(Z)
    (reference experiments/model/core/kernels.py:98-110,155-172; svpy.py:118-121), and DivergenceFreeKernel.compute_nu fused with
    the (M D x M D) Gram matrix (kernels.py:289-303,376-387): the L samples share Z, lengthscales and variance, hence ONE matrix and
    ONE factorisation (spread over the chip, csrc/setup_kernels.cu k_chol_step) where the reference refactors it per sample.
    u_prior, u: (L,M,D_out); returns (nu, info): nu (L,D_out,M,1) [dimwise], (L,M,D_out) [shared] or (L,M*D,1) [DF]; info (Kc,) int32 on the device,
    0 or 1 + the index of the first non-positive pivot of K(Z,Z) + jitter -- where torch.linalg.cholesky raises in the reference.
    No host synchronisation happens here; SVGP_Layer reads info (one 4-byte D2H) when its check_cholesky attribute is set."""

    @staticmethod
    def forward(ctx, Z, ell, var, u_prior, u, variant):
        lib = _lib.load()
        Z, ell, var, u_prior, u = map(_c, (Z, ell, var, u_prior, u))
        _lib.require_cuda(Z, ell, var, u_prior, u)
        M, D_in = Z.shape
        L, _, D_out = u.shape
        if u_prior.shape != u.shape or u.shape[1] != M:
            raise RuntimeError("u_prior and u must both be (L, M=%d, D_out), got %s / %s" % (M, tuple(u_prior.shape), tuple(u.shape)))
        dev = Z.device
        with torch.cuda.device(dev):
            p = _lib.make_problem(variant, L, 1, D_in, D_out, M, 1, Z, ell, var, None, None, None, None, None)
            nbytes = lib.gpode_nu_workspace_bytes(ctypes.byref(p))
            nsave = lib.gpode_nu_save_floats(ctypes.byref(p))
            if nbytes == 0 or nsave == 0:
                raise RuntimeError("gpode_compute_nu: unsupported problem (M <= 512, D <= 16; DF: D <= 8 and M * D <= 4096)")
            ws = torch.empty(nbytes, dtype=torch.uint8, device=dev)
            save = torch.empty(nsave, dtype=torch.float32, device=dev)
            info = torch.zeros(D_out if variant == _lib.RBF_DIMWISE else 1, dtype=torch.int32, device=dev)
            shape = {_lib.RBF_DIMWISE: (L, D_out, M, 1), _lib.DF: (L, M * D_out, 1)}.get(variant, (L, M, D_out))
            nu = torch.empty(shape, dtype=torch.float32, device=dev)
            rc = lib.gpode_compute_nu_fwd(ctypes.byref(p), _lib.ptr(u_prior), _lib.ptr(u), _lib.ptr(nu), _lib.ptr(save), _lib.ptr(info),
                                          _lib.ptr(ws), nbytes, _lib.stream_handle(dev))
        _lib.check(rc, "gpode_compute_nu_fwd")
        ctx.save_for_backward(Z, ell, var, u, save)
        ctx.variant = variant
        ctx.mark_non_differentiable(info)
        return nu, info

    @staticmethod
    def backward(ctx, dnu, _dinfo=None):
        lib = _lib.load()
        Z, ell, var, u, save = ctx.saved_tensors
        variant = ctx.variant
        dnu = dnu.contiguous()
        M, D_in = Z.shape
        L, _, D_out = u.shape
        dev = Z.device
        with torch.cuda.device(dev):
            p = _lib.make_problem(variant, L, 1, D_in, D_out, M, 1, Z, ell, var, None, None, None, None, None)
            nbytes = lib.gpode_nu_workspace_bytes(ctypes.byref(p))
            ws = torch.empty(nbytes, dtype=torch.uint8, device=dev)
            dup, du = torch.empty_like(u), torch.empty_like(u)
            dZ, dell, dvar = torch.empty_like(Z), torch.empty_like(ell), torch.empty_like(var)
            rc = lib.gpode_compute_nu_bwd(ctypes.byref(p), _lib.ptr(u), _lib.ptr(save), _lib.ptr(dnu), _lib.ptr(dup), _lib.ptr(du),
                                          _lib.ptr(dZ), _lib.ptr(dell), _lib.ptr(dvar), _lib.ptr(ws), nbytes, _lib.stream_handle(dev))
        _lib.check(rc, "gpode_compute_nu_bwd")
        return dZ, dell, dvar, dup, du, None


class InducingSample(torch.autograd.Function):
    """u[l] = tril(Lq) eps_u[l] + Um on the packed parameter (reference svpy.py:88-101 + transforms.py:71-77)."""

    @staticmethod
    def forward(ctx, Lq_packed, Um, eps_u):
        lib = _lib.load()
        Lq_packed, Um, eps_u = map(_c, (Lq_packed, Um, eps_u))
        _lib.require_cuda(Lq_packed, Um, eps_u)
        L, M, D = eps_u.shape
        if Lq_packed.shape != (D, M * (M + 1) // 2) or Um.shape != (M, D):
            raise RuntimeError("Lq_packed must be (D_out, M(M+1)/2) and Um (M, D_out)")
        u = torch.empty_like(eps_u)
        with torch.cuda.device(u.device):
            rc = lib.gpode_inducing_sample_fwd(L, M, D, _lib.ptr(Lq_packed), _lib.ptr(Um), _lib.ptr(eps_u), _lib.ptr(u),
                                               _lib.stream_handle(u.device))
        _lib.check(rc, "gpode_inducing_sample_fwd")
        ctx.save_for_backward(eps_u)
        ctx.shape = (Lq_packed.shape, Um.shape)
        return u

    @staticmethod
    def backward(ctx, du):
        lib = _lib.load()
        (eps_u,) = ctx.saved_tensors
        du = du.contiguous()
        L, M, D = eps_u.shape
        dLq = torch.empty(ctx.shape[0], dtype=torch.float32, device=du.device)
        dUm = torch.empty(ctx.shape[1], dtype=torch.float32, device=du.device)
        with torch.cuda.device(du.device):
            rc = lib.gpode_inducing_sample_bwd(L, M, D, _lib.ptr(eps_u), _lib.ptr(du), _lib.ptr(dLq), _lib.ptr(dUm),
                                               _lib.stream_handle(du.device))
        _lib.check(rc, "gpode_inducing_sample_bwd")
        return dLq, dUm, None


class WhitenedKL(torch.autograd.Function):
    """SVGP_Layer.kl (reference svpy.py:144-175, q_diag=False) on the packed lower-triangular parameter."""

    @staticmethod
    def forward(ctx, Lq_packed, Um):
        lib = _lib.load()
        Lq_packed, Um = map(_c, (Lq_packed, Um))
        _lib.require_cuda(Lq_packed, Um)
        M, D = Um.shape
        kl = torch.empty((), dtype=torch.float32, device=Um.device)
        with torch.cuda.device(Um.device):
            rc = lib.gpode_kl_fwd(M, D, _lib.ptr(Lq_packed), _lib.ptr(Um), _lib.ptr(kl), _lib.stream_handle(Um.device))
        _lib.check(rc, "gpode_kl_fwd")
        ctx.save_for_backward(Lq_packed, Um)
        return kl

    @staticmethod
    def backward(ctx, dkl):
        lib = _lib.load()
        Lq_packed, Um = ctx.saved_tensors
        M, D = Um.shape
        dkl = dkl.contiguous().to(torch.float32)
        dLq, dUm = torch.empty_like(Lq_packed), torch.empty_like(Um)
        with torch.cuda.device(Um.device):
            rc = lib.gpode_kl_bwd(M, D, _lib.ptr(Lq_packed), _lib.ptr(Um), _lib.ptr(dkl), _lib.ptr(dLq), _lib.ptr(dUm),
                                  _lib.stream_handle(Um.device))
        _lib.check(rc, "gpode_kl_bwd")
        return dLq, dUm


def compute_nu(Z, ell, var, u_prior, u, variant, return_info=False):
    """nu for L samples; with return_info also the Cholesky status vector (see ComputeNu)."""
    v = _lib.VARIANTS[variant] if isinstance(variant, str) else variant
    nu, info = ComputeNu.apply(Z, ell, var, u_prior, u, v)
    return (nu, info) if return_info else nu


def inducing_sample(Lq_packed, Um, eps_u):
    return InducingSample.apply(Lq_packed, Um, eps_u)


def whitened_kl(Lq_packed, Um):
    return WhitenedKL.apply(Lq_packed, Um)


# ------------------------------------------------------------------------------------------------
# either side of the flow (csrc/elbo_kernels.cu): device-side draws, fused Bernoulli log-likelihood
# ------------------------------------------------------------------------------------------------
NORMAL, UNIFORM = 0, 1


class PhiloxStream:
    """Counter-based draw stream on one CUDA device (Philox4x32-10, include/gpode.h gpode_philox_fill): a pure function of
    (seed, offset) -- `offset` advances with every call, so a stream replays exactly from its seed.  Replaces the host numpy
    helpers of the reference (kernels.py:13-26, svpy.py:12-18) and the H2D copies of their results when selected with
    ``gpode_b200.set_rng("device", seed)``: same distributions, different numbers (the reference's own draws are unseeded)."""

    def __init__(self, seed=0):
        self.seed, self.offset = int(seed) & (2 ** 64 - 1), 0

    def fill(self, outs, kinds):
        """outs: up to four contiguous fp32 CUDA tensors filled in ONE launch; kinds: NORMAL / UNIFORM per tensor"""
        lib = _lib.load()
        n = len(outs)
        assert 1 <= n <= 4 and len(kinds) == n
        _lib.require_cuda(*outs)
        ptrs = (ctypes.c_void_p * n)(*[t.data_ptr() for t in outs])
        counts = (ctypes.c_uint64 * n)(*[t.numel() for t in outs])
        ks = (ctypes.c_int32 * n)(*[int(k) for k in kinds])
        dev = outs[0].device
        with torch.cuda.device(dev):
            rc = lib.gpode_philox_fill(n, ptrs, counts, ks, self.seed, self.offset, _lib.stream_handle(dev))
        _lib.check(rc, "gpode_philox_fill")
        self.offset += (max(t.numel() for t in outs) + 3) // 4
        return outs


class BernoulliLhood(torch.autograd.Function):
    """lhood[n] = mean_l sum_{t,pixels} log(z) x + log(1 - z)(1 - x): Decoder.log_prob (reference core/vae.py:136-153) fused with
    the reduction elbo() applies to it (create_model.py:51-53), without X.repeat(L) and without a (L,N,T,1,28,28) intermediate.
    x (N,T,...) targets, z (L,N,T,...) reconstructions; gradient for z only (the data carries none)."""

    @staticmethod
    def forward(ctx, x, z):
        lib = _lib.load()
        x, z = x.contiguous(), z.contiguous()
        _lib.require_cuda(x, z)
        L, N = z.shape[0], z.shape[1]
        P = z[0, 0].numel()
        if x.shape[0] != N or x[0].numel() != P:
            raise RuntimeError("x must be (N, ...) and z (L, N, ...) over the same trailing shape, got %s / %s" % (tuple(x.shape), tuple(z.shape)))
        dev = z.device
        lhood = torch.empty(N, dtype=torch.float32, device=dev)
        with torch.cuda.device(dev):
            nbytes = lib.gpode_bernoulli_workspace_bytes(N)
            ws = torch.empty(nbytes, dtype=torch.uint8, device=dev)
            rc = lib.gpode_bernoulli_lhood_fwd(L, N, P, _lib.ptr(z), _lib.ptr(x), _lib.ptr(lhood), _lib.ptr(ws), nbytes, _lib.stream_handle(dev))
        _lib.check(rc, "gpode_bernoulli_lhood_fwd")
        ctx.save_for_backward(x, z)
        return lhood

    @staticmethod
    def backward(ctx, g):
        lib = _lib.load()
        x, z = ctx.saved_tensors
        L, N = z.shape[0], z.shape[1]
        P = z[0, 0].numel()
        g = g.contiguous().to(torch.float32)
        dz = torch.empty_like(z)
        with torch.cuda.device(z.device):
            rc = lib.gpode_bernoulli_lhood_bwd(L, N, P, _lib.ptr(z), _lib.ptr(x), _lib.ptr(g), _lib.ptr(dz), _lib.stream_handle(z.device))
        _lib.check(rc, "gpode_bernoulli_lhood_bwd")
        return None, dz


def bernoulli_lhood(x, z):
    """(N,) per-trajectory Bernoulli log-likelihood of reconstructions z (L,N,T,...) for targets x (N,T,...), averaged over L"""
    return BernoulliLhood.apply(x, z)
