"""Index-explicit CPU restatement of the sparse-GP vector field and its per-rollout cache.

TEST INFRASTRUCTURE ONLY (see oracle/__init__.py).  dtype-generic torch code: run it in float64 for
the "truth" and in float32 for a like-for-like port; being plain torch it is differentiable, so
``torch.autograd`` over these functions is the gradient oracle.  Each function cites the reference
lines it restates (paths relative to /root/reference/experiments/model).

Variants:  "rbf_dimwise" (core/kernels.py RBF, dimwise=True -- the default, main.py:63),
           "rbf_shared"  (RBF, dimwise=False), "df" (DivergenceFreeKernel, core/kernels.py:201-393).
"""
import math

import torch

JITTER = 1e-5          # core/kernels.py:11, core/svpy.py:10
SOFTPLUS_FLOOR = 1e-12  # misc/constraint_utils.py:5-8


def constrain(raw):
    """lengthscales / variance properties: softplus(raw) + 1e-12 (kernels.py:56-62)."""
    return torch.nn.functional.softplus(raw) + SOFTPLUS_FLOOR


def unconstrain(value):
    """inverse of :func:`constrain` (constraint_utils.py:10-13)."""
    v = torch.clamp(value - SOFTPLUS_FLOOR, min=torch.finfo(value.dtype).eps)
    return v + torch.log(-torch.expm1(-v))


def tril_from_packed(packed, M):
    """LowerTriangular.forward_tensor: row-major tril scatter (D,M(M+1)/2)->(D,M,M) (transforms.py:71-77)."""
    D = packed.shape[0]
    rows, cols = torch.tril_indices(M, M, 0)
    out = packed.new_zeros((D, M, M))
    out[:, rows, cols] = packed
    return out


# ----------------------------------------------------------------------------------------------
# RBF
# ----------------------------------------------------------------------------------------------
def rbf_K(X, X2, ell, var, dimwise):
    """K(X, X2): dimwise -> (D_out,N,M) = var_k exp(-1/2 sum_d ((X_nd-X2_md)/ell_kd)^2); shared -> (N,M).
    (kernels.py:64-110; written with direct differences, algebraically equal to the reference's expansion)."""
    if X2 is None:
        X2 = X
    diff = X[:, None, :] - X2[None, :, :]                                  # (N,M,D_in)
    if dimwise:
        sq = (diff[None] / ell[:, None, None, :]).pow(2).sum(-1)           # (D_out,N,M)
        return var[:, None, None] * torch.exp(-0.5 * sq)
    sq = (diff / ell).pow(2).sum(-1)                                       # (N,M)
    return var * torch.exp(-0.5 * sq)


def rbf_rff(x, omega, phase, w, var, dimwise):
    """Prior sample f_p (N,D_out) (kernels.py:140-153).
    dimwise: omega (D_in,S,D_out), phase (1,S,D_out), w (S,D_out); shared: omega (D_in,S), phase (1,S)."""
    S = w.shape[0]
    if dimwise:
        theta = torch.einsum("nd,dsk->nsk", x, omega) + phase              # (N,S,D_out)
        return torch.sqrt(var / S)[None, :] * (torch.cos(theta) * w[None]).sum(1)
    theta = x @ omega + phase                                              # (N,S)
    return torch.sqrt(var / S) * (torch.cos(theta) @ w)


def rbf_f_update(x, Z, nu, ell, var, dimwise):
    """Pathwise update f_u (N,D_out) = K(Z,x)^T nu (kernels.py:174-181). nu: dimwise (D_out,M,1); shared (M,D_out)."""
    if dimwise:
        K = rbf_K(Z, x, ell, var, True)                                    # (D_out,M,N)
        return torch.einsum("km,kmn->nk", nu[..., 0], K)
    K = rbf_K(Z, x, ell, var, False)                                       # (M,N)
    return K.t() @ nu


# ----------------------------------------------------------------------------------------------
# divergence-free kernel (requires D_in == D_out == D; ell (D,D) indexed by block entry, var (D,))
# ----------------------------------------------------------------------------------------------
def df_K(X, X2, ell, var):
    """Block kernel laid out (N*D, M*D), row n*D+i, col m*D+j (kernels.py:289-303, 217-242, 259-262):
    Kb_ij = var_j exp(-r2/(2 ell_ij^2)) / ell_ij^2 * (d_i d_j/ell_ij^2 + delta_ij ((D-1) - r2/ell_ij^2)), d = X2_m - X_n."""
    if X2 is None:
        X2 = X
    N, D = X.shape
    M = X2.shape[0]
    d = X2[None, :, :] - X[:, None, :]                                     # (N,M,D)  second minus first
    r2 = d.pow(2).sum(-1)                                                  # (N,M)
    c = 1.0 / ell.pow(2)                                                   # (D,D)
    E = torch.exp(-0.5 * r2[:, :, None, None] * c)                         # (N,M,D,D)
    H = d[:, :, :, None] * d[:, :, None, :] * c \
        + torch.eye(D, dtype=X.dtype) * ((D - 1.0) - r2[:, :, None, None] * c)
    Kb = var[None, None, None, :] * E * H * c                              # (N,M,D,D)  var on column index j
    return Kb.permute(0, 2, 1, 3).reshape(N * D, M * D)


def df_B(omega):
    """B[s,a,c] = |omega[:,s,c]| delta_ac - (sum_b omega[a,s,b] omega[c,s,b]) / |omega[:,s,c]| (kernels.py:327-336)."""
    D = omega.shape[0]
    norm = torch.sqrt(omega.pow(2).sum(0))                                 # (S,D) indexed [s,c]
    ww = torch.einsum("asb,csb->sac", omega, omega)                        # (S,D,D)
    return norm[:, None, :] * torch.eye(D, dtype=omega.dtype)[None] - ww / norm[:, None, :]


def df_rff(x, omega, phase, w, var, B=None):
    """f_p[n,c] = sqrt(var_c/S) sum_{s,a} B[s,a,c] (cos th_nsa w[s,a] + sin th_nsa w[S+s,a]) (kernels.py:319-351).
    omega (D,S,D), phase (1,S,D), w (2S,D)."""
    S = omega.shape[1]
    if B is None:
        B = df_B(omega)
    theta = torch.einsum("nd,dsa->nsa", x, omega) + phase                  # (N,S,D)
    u = torch.cos(theta) * w[None, :S] + torch.sin(theta) * w[None, S:]    # (N,S,D)
    return torch.sqrt(var / S)[None, :] * torch.einsum("nsa,sac->nc", u, B)


def df_f_update(x, Z, nu, ell, var):
    """f_u[n,j] = sum_{m,i} nu[m*D+i] Kb(Z,x)[m*D+i, n*D+j] (kernels.py:390-393); nu (M*D,1)."""
    K = df_K(Z, x, ell, var)                                               # (M*D, N*D)
    return (nu[:, 0] @ K).reshape(x.shape)


# ----------------------------------------------------------------------------------------------
# cache (per rollout) -- core/svpy.py:88-121, core/kernels.py:112-137,155-172,305-316,376-387
# ----------------------------------------------------------------------------------------------
def make_omega(eps, ell, variant):
    """sample_freq: omega = eps / ell, ell broadcast as (D_in,1,D_out) (dimwise, DF) or (D_in,1) (kernels.py:120-124)."""
    if variant == "rbf_shared":
        return eps / ell[:, None]
    return eps / ell.t()[:, None, :]


def sample_inducing(Lq, eps_u, Um, q_diag=False):
    """u = Lq eps_u + m ('dnm,md->nd'), or diag (svpy.py:88-101). Lq (D_out,M,M) or (M,D_out) if q_diag."""
    if q_diag:
        return Lq * eps_u + Um
    return torch.einsum("dnm,md->nd", Lq, eps_u) + Um


def compute_nu(Ku, u_prior, u, variant):
    """nu = L^-T (u - L^-1 f_p(Z)),  L = chol(Ku + 1e-5 I) lower triangle only (kernels.py:155-172, 376-387)."""
    n = Ku.shape[-1]
    L = torch.linalg.cholesky(Ku + JITTER * torch.eye(n, dtype=Ku.dtype))
    if variant == "rbf_dimwise":
        a = torch.linalg.solve_triangular(L, u_prior.t()[:, :, None], upper=False)
        return torch.linalg.solve_triangular(L.transpose(1, 2), u.t()[:, :, None] - a, upper=True)   # (D_out,M,1)
    if variant == "rbf_shared":
        a = torch.linalg.solve_triangular(L, u_prior, upper=False)
        return torch.linalg.solve_triangular(L.t(), u - a, upper=True)                               # (M,D_out)
    a = torch.linalg.solve_triangular(L, u_prior.reshape(n)[:, None], upper=False)
    return torch.linalg.solve_triangular(L.t(), u.reshape(n)[:, None] - a, upper=True)               # (M*D,1)


def prior(x, c):
    v = c["variant"]
    if v == "df":
        return df_rff(x, c["omega"], c["phase"], c["w"], c["var"], c.get("B"))
    return rbf_rff(x, c["omega"], c["phase"], c["w"], c["var"], v == "rbf_dimwise")


def update(x, c):
    v = c["variant"]
    if v == "df":
        return df_f_update(x, c["Z"], c["nu"], c["ell"], c["var"])
    return rbf_f_update(x, c["Z"], c["nu"], c["ell"], c["var"], v == "rbf_dimwise")


def field(x, c):
    """SVGP_Layer.forward: f = rff_forward(x) + f_update(x, Z) (svpy.py:123-142)."""
    return prior(x, c) + update(x, c)


def gram(c):
    v = c["variant"]
    if v == "df":
        return df_K(c["Z"], None, c["ell"], c["var"])
    return rbf_K(c["Z"], None, c["ell"], c["var"], v == "rbf_dimwise")


def build_cache(variant, Z, raw_ell, raw_var, Um, Lq, draws, q_diag=False, nu_override=None):
    """SVGP_Layer.build_cache (svpy.py:103-121) with the host draws made explicit.
    draws = dict(w, eps, phase01, eps_u) in the reference's draw order (kernels.py:126-137 / 305-316, svpy.py:94):
    w ~ N(0,1) (S,D_out) [DF (2S,D_out)], eps ~ N(0,1) omega-shaped, phase01 ~ U(0,1) phase-shaped, eps_u ~ N(0,1) (M,D_out)."""
    ell, var = constrain(raw_ell), constrain(raw_var)
    c = dict(variant=variant, Z=Z, ell=ell, var=var, w=draws["w"],
             omega=make_omega(draws["eps"], ell, variant), phase=draws["phase01"] * 2 * math.pi)
    if variant == "df":
        c["B"] = df_B(c["omega"])
    u = sample_inducing(Lq, draws["eps_u"], Um, q_diag)
    c["u"] = u
    if nu_override is not None:
        c["nu"] = nu_override
    else:
        c["nu"] = compute_nu(gram(c), prior(Z, c), u, variant)
    return c


def kl_whitened(Um, Lq, q_diag=False):
    """SVGP_Layer.kl (svpy.py:144-175): 1/2 sum_d(-logdet(Lq_d Lq_d^T) + |m_d|^2 + |Lq_d|_F^2 - M)."""
    M = Um.shape[0]
    if q_diag:
        diag, trace = Lq, Lq.pow(2).sum(0)
    else:
        Lt = torch.tril(Lq)
        diag, trace = torch.diagonal(Lt, dim1=1, dim2=2).t(), Lt.pow(2).sum((1, 2))
    two_kl = -torch.log(diag.pow(2)).sum(0) + Um.pow(2).sum(0) + trace - M
    return 0.5 * two_kl.sum()


# ----------------------------------------------------------------------------------------------
# ODE right-hand side and rollout -- core/flow.py:30-45,68-86, core/odegpvae.py:37-45
# ----------------------------------------------------------------------------------------------
def rhs(sv, c, order):
    """order 1: f(sv); order 2: [sv[:,q:], f(sv)] (flow.py:30-38)."""
    if order == 1:
        return field(sv, c)
    q = sv.shape[1] // 2
    return torch.cat([sv[:, q:], field(sv, c)], 1)


def rollout(z0, ts, c, order, method):
    """Flow.forward for a fixed function sample: (N,T,D_s) (flow.py:68-86)."""
    from oracle import solvers
    zt = solvers.odeint(lambda t, y: rhs(y, c, order), z0, ts, method=method)
    return zt.permute(1, 0, 2)


def cast_cache(c, dtype):
    return {k: (v.to(dtype) if torch.is_tensor(v) else v) for k, v in c.items()}
