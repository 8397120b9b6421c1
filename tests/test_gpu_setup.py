"""GPU parity of the per-rollout setup kernels (csrc/setup_kernels.cu) through the C ABI: compute_nu (K(Z,Z) + blocked
Cholesky + whitened solves, batched over output dims and MC samples), inducing sample and KL on the packed
lower-triangular parameter -- against the fp64 oracle (oracle/field.py) and its autograd.

Tolerances: nu goes through a Cholesky of a matrix with cond up to ~1e5, so fp32 nu is compared
with the fp64 oracle at 1e-3 x cond-aware bars on well-conditioned problems (ell <= 1) and against the fp32 torch result
on the golden (ell = 2) cases; everything else at 1e-5 / 1e-4."""
import numpy as np
import pytest
import torch

from oracle import field as OF
from helpers import load_golden, oracle_cache, rel, t

pytestmark = pytest.mark.gpu


def _gp():
    import gpode_b200
    return gpode_b200


def _problem(variant, M, D_in, D_out, L, seed, ell0=0.7):
    rs = np.random.RandomState(seed)
    f64 = lambda a: torch.tensor(a, dtype=torch.float64)
    dimwise = variant == "rbf_dimwise"
    Z = f64(rs.normal(size=(M, D_in)))
    ell = f64(ell0 + 0.3 * rs.uniform(size=(D_out, D_in) if dimwise else (D_in,)))
    var = f64(0.5 + rs.uniform(size=(D_out,) if dimwise else (1,)))
    up = f64(rs.normal(size=(L, M, D_out)))
    u = f64(rs.normal(size=(L, M, D_out)))
    return Z, ell, var, up, u


def _oracle_nu(variant, Z, ell, var, up, u):
    Ku = OF.rbf_K(Z, None, ell, var, variant == "rbf_dimwise")
    return torch.stack([OF.compute_nu(Ku, up[l], u[l], variant) for l in range(u.shape[0])])


@pytest.mark.parametrize("variant", ["rbf_dimwise", "rbf_shared"])
@pytest.mark.parametrize("M,D_in,D_out,L", [(100, 6, 6, 1), (33, 3, 2, 3), (256, 6, 3, 2), (512, 16, 16, 4), (70, 16, 5, 8)])
def test_compute_nu_forward_backward(variant, M, D_in, D_out, L):
    Z, ell, var, up, u = _problem(variant, M, D_in, D_out, L, seed=M + D_out)
    leaves = [v.clone().requires_grad_(True) for v in (Z, ell, var, up, u)]
    nu64 = _oracle_nu(variant, *leaves)
    G = torch.tensor(np.random.RandomState(1).normal(size=tuple(nu64.shape)), dtype=torch.float64)
    want = torch.autograd.grad((nu64 * G).sum(), leaves)
    # the same computation in fp32 torch on the CPU (LAPACK): its distance from the fp64 result is the conditioning noise floor
    leaves32 = [v.clone().float().requires_grad_(True) for v in (Z, ell, var, up, u)]
    nu32 = _oracle_nu(variant, *leaves32)
    floor32 = torch.autograd.grad((nu32 * G.float()).sum(), leaves32)
    dev = [v.detach().float().cuda().requires_grad_(True) for v in (Z, ell, var, up, u)]
    nu = _gp().compute_nu(*dev, variant)
    assert nu.shape == nu64.shape
    e, e32 = rel(nu, nu64), rel(nu32, nu64)
    print("%s M=%d nu: %.2e (torch fp32: %.2e)" % (variant, M, e, e32))
    assert e < max(5 * e32, 2e-5), (e, e32)
    (nu * G.float().cuda()).sum().backward()
    for nm, a, b, f in zip(("dZ", "dell", "dvar", "du_prior", "du"), dev, want, floor32):
        e, e32 = rel(a.grad, b), rel(f, b)
        print("%s M=%d %s: %.2e (torch fp32: %.2e)" % (variant, M, nm, e, e32))
        assert e < max(5 * e32, 1e-4), (nm, e, e32)


@pytest.mark.parametrize("M,D,L,ell0,asym", [(100, 6, 4, 0.7, 0.3), (100, 6, 1, 2.0, 0.0), (33, 3, 2, 0.7, 0.3), (256, 6, 2, 1.0, 0.05),
                                                 (70, 8, 3, 0.7, 0.3), (10, 2, 1, 0.7, 0.3), (400, 8, 1, 0.5, 0.1), (100, 6, 12, 0.7, 0.3)])
def test_df_compute_nu_forward_backward(M, D, L, ell0, asym):
    """DivergenceFreeKernel.compute_nu fused with its (M D x M D) Gram matrix (kernels.py:289-303,376-387): K build, the
    multi-CTA blocked Cholesky, both whitened solves and the closed-form backward, against the fp64 oracle and its autograd.
    asym > 0: element-wise different lengthscales / per-column variances -> the Gram matrix is NOT symmetric and the reference
    semantics (torch.linalg.cholesky reads the lower triangle; its backward feeds the symmetrised gradient to both triangles)
    must be reproduced (asym is kept small enough for the lower-triangle matrix to stay positive definite).  Sizes: config 2 (600), config 4 (1,536), a ragged order (99), tiny (20) and a large one (3,200: 4-column solves)."""
    rs = np.random.RandomState(M + D)
    f64 = lambda a: torch.tensor(a, dtype=torch.float64)
    Z = f64(rs.normal(size=(M, D)))
    ell = f64(ell0 + asym * rs.uniform(size=(D, D)))
    var = f64(1.0 + asym * rs.uniform(size=(D,)))
    up, u = f64(rs.normal(size=(L, M, D))), f64(rs.normal(size=(L, M, D)))

    def oracle(Z, ell, var, up, u):
        Ku = OF.df_K(Z, None, ell, var)
        return torch.stack([OF.compute_nu(Ku, up[l], u[l], "df") for l in range(L)])
    big = M * D > 2000                                     # the fp64 CPU oracle of the largest case: forward and du only (autograd of a 4096^3 Cholesky is slow)
    leaves = [v.clone().requires_grad_(True) for v in (Z, ell, var, up, u)]
    nu64 = oracle(*leaves)
    G = torch.tensor(np.random.RandomState(1).normal(size=tuple(nu64.shape)), dtype=torch.float64)
    want = torch.autograd.grad((nu64 * G).sum(), leaves)
    floor32 = None
    e32 = 0.0
    if not big:
        try:                                               # (at ell = 2 the fp32 Gram matrix is not even positive definite any more)
            leaves32 = [v.clone().float().requires_grad_(True) for v in (Z, ell, var, up, u)]
            nu32 = oracle(*leaves32)
            floor32 = torch.autograd.grad((nu32 * G.float()).sum(), leaves32)
            e32 = rel(nu32, nu64)
        except torch.linalg.LinAlgError:
            floor32 = None
    dev = [v.detach().float().cuda().requires_grad_(True) for v in (Z, ell, var, up, u)]
    nu, info = _gp().compute_nu(*dev, "df", return_info=True)
    assert nu.shape == (L, M * D, 1) and int(info.abs().max()) == 0
    e = rel(nu, nu64)
    print("df M=%d D=%d (order %d) nu: %.2e (torch fp32: %.2e)" % (M, D, M * D, e, e32))
    assert e < max(5 * e32, 2e-5), (e, e32)
    (nu * G.float().cuda()).sum().backward()
    for i, nm in enumerate(("dZ", "dell", "dvar", "du_prior", "du")):
        e = rel(dev[i].grad, want[i])
        e32 = rel(floor32[i], want[i]) if floor32 is not None else 0.0
        print("df M=%d D=%d %s: %.2e (torch fp32: %.2e)" % (M, D, nm, e, e32))
        assert e < max(5 * e32, 1e-4), (nm, e, e32)


def test_df_cholesky_failure_is_reported():
    rs = np.random.RandomState(0)
    M, D = 40, 3
    f32 = lambda a: torch.tensor(a, dtype=torch.float32, device="cuda")
    Z, ell, var = f32(rs.normal(size=(M, D))), f32(0.7 + np.zeros((D, D))), f32(-np.ones(D))     # negative variance: not positive definite
    up, u = f32(rs.normal(size=(1, M, D))), f32(rs.normal(size=(1, M, D)))
    nu, info = _gp().compute_nu(Z, ell, var, up, u, "df", return_info=True)
    assert int(info[0]) >= 1


@pytest.mark.parametrize("name", ["rbf_dimwise_o1", "rbf_shared_o1", "rbf_dimwise_d16"])
def test_compute_nu_matches_reference_golden(name):
    """reference nu (fp32 LAPACK, ell = 2 -> cond ~1e4..1e6): same inputs through the CUDA kernels; the bar is the
    reference's own fp32 noise against the fp64 oracle."""
    g = load_golden(name)
    m = g["meta"]
    c64 = oracle_cache(g, shared_nu=False)
    up64 = OF.prior(c64["Z"], c64)
    nu = _gp().compute_nu(c64["Z"].float().cuda(), c64["ell"].float().cuda(), c64["var"].float().cuda(), up64.float().cuda()[None],
                          c64["u"].float().cuda()[None], m["variant"])
    e_new, e_ref = rel(nu[0], c64["nu"]), rel(g["field_nu"], c64["nu"])
    print("%s nu: new-vs-fp64 %.2e ref-vs-fp64 %.2e" % (name, e_new, e_ref))
    assert e_new < max(5 * e_ref, 1e-4)


def test_cholesky_failure_is_reported():
    Z, ell, var, up, u = _problem("rbf_dimwise", 40, 3, 2, 1, seed=0)
    Z[1] = Z[0]                       # duplicate inducing point: K + 1e-5 I stays SPD (jitter) -> info == 0
    from gpode_b200 import functional as GF
    nu, info = GF.ComputeNu.apply(Z.float().cuda(), ell.float().cuda(), var.float().cuda(), up.float().cuda(), u.float().cuda(), 1)
    assert torch.isfinite(nu).all() and int(info.abs().max()) == 0
    var_bad = -var                    # negative variance: not positive definite -> NaN factor, reported like LAPACK's info
    nu, info = _gp().compute_nu(Z.float().cuda(), ell.float().cuda(), var_bad.float().cuda(), up.float().cuda(), u.float().cuda(), "rbf_dimwise",
                                return_info=True)
    assert not torch.isfinite(nu).all() and int(info.min()) >= 1      # 1 + index of the first non-positive pivot, per matrix


@pytest.mark.parametrize("M,D,L", [(10, 3, 1), (100, 6, 4), (512, 16, 2), (257, 5, 3)])
def test_inducing_sample_and_kl(M, D, L):
    rs = np.random.RandomState(M)
    P = M * (M + 1) // 2
    packed = torch.tensor(0.05 * rs.normal(size=(D, P)), dtype=torch.float64)
    rows, cols = torch.tril_indices(M, M)
    packed[:, rows == cols] = torch.tensor(0.5 + rs.uniform(size=(D, M)), dtype=torch.float64)
    Um = torch.tensor(rs.normal(size=(M, D)), dtype=torch.float64)
    eps = torch.tensor(rs.normal(size=(L, M, D)), dtype=torch.float64)
    p64, m64 = packed.clone().requires_grad_(True), Um.clone().requires_grad_(True)
    Lq = OF.tril_from_packed(p64, M)
    u64 = torch.stack([OF.sample_inducing(Lq, eps[l], m64) for l in range(L)])
    kl64 = OF.kl_whitened(m64, Lq)
    G = torch.tensor(rs.normal(size=(L, M, D)), dtype=torch.float64)
    want = torch.autograd.grad((u64 * G).sum() + 0.7 * kl64, [p64, m64])
    pc, mc = packed.float().cuda().requires_grad_(True), Um.float().cuda().requires_grad_(True)
    u = _gp().inducing_sample(pc, mc, eps.float().cuda())
    kl = _gp().whitened_kl(pc, mc)
    assert rel(u, u64) < 1e-5
    assert abs(kl.item() - kl64.item()) < 1e-5 * abs(kl64.item())
    ((u * G.float().cuda()).sum() + 0.7 * kl).backward()
    assert rel(pc.grad, want[0]) < 1e-5
    assert rel(mc.grad, want[1]) < 1e-5
