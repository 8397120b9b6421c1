"""GPU parity of the divergence-free (DF) vector-field kernels through the C ABI (ctypes -> libgpode.so).

Same bars as the RBF kernels (north_star, fp32, norm-wise): field 1e-5, trajectories 1e-4, gradients 1e-4.
Kernel-level gradients are checked against autograd through the fp64 oracle with Z, ell, var, nu AND B as
independent leaves -- the tensors the C ABI differentiates (include/gpode.h: d_ell is the direct dependence
only, the dependence through B = B(eps/ell) is returned as d_B and carried by PyTorch)."""
import numpy as np
import pytest
import torch

from oracle import field as OF
from helpers import DF_CASES, gpu_sample, load_golden, oracle_cache, rel, t

pytestmark = pytest.mark.gpu

FIELD_TOL = 1e-5
TRAJ_TOL = 1e-4
GRAD_TOL = 1e-4
LEAVES = ("Z", "nu", "ell", "var", "B")


def _gp():
    import gpode_b200
    return gpode_b200


def _leaf_cache(g):
    c = oracle_cache(g, leaves=True)
    c["B"] = c["B"].detach().clone().requires_grad_(True)
    return c


def _rollout(s, z0, ts, method):
    return _gp().gp_rollout(z0, ts, s["Z"], s["nu"], s["eps"], s["phase"], s["w"], s["ell"], s["var"], "df", 1, method, s["B"])


@pytest.mark.parametrize("name", DF_CASES)
def test_field_forward(name):
    g = load_golden(name)
    c64 = oracle_cache(g)
    s = gpu_sample(c64)
    x = t(g["x"], device="cuda")[None]
    f, fp = _gp().gp_field(x, s["Z"], s["nu"], s["eps"], s["phase"], s["w"], s["ell"], s["var"], "df", s["B"])
    truth = OF.field(t(g["x"], torch.float64), c64)
    e_ref, e_truth, e_floor = rel(f[0], g["field_f"]), rel(f[0], truth), rel(g["field_f"], truth)
    print("%s field: new-vs-ref %.2e new-vs-fp64 %.2e ref-vs-fp64 %.2e" % (name, e_ref, e_truth, e_floor))
    assert e_truth < FIELD_TOL, (e_ref, e_truth, e_floor)
    assert e_ref < FIELD_TOL + e_floor, (e_ref, e_truth, e_floor)
    assert rel(fp[0], OF.prior(t(g["x"], torch.float64), c64)) < FIELD_TOL


@pytest.mark.parametrize("name", DF_CASES)
def test_field_backward_kernel_level(name):
    g = load_golden(name)
    c = _leaf_cache(g)
    x64 = t(g["x"], torch.float64).requires_grad_(True)
    gout = t(g["g"], torch.float64)
    want = torch.autograd.grad((OF.field(x64, c) * gout).sum(), [x64] + [c[k] for k in LEAVES])
    s = gpu_sample(c)
    for k in LEAVES:
        s[k].requires_grad_(True)
    x = t(g["x"], device="cuda")[None].requires_grad_(True)
    f, _ = _gp().gp_field(x, s["Z"], s["nu"], s["eps"], s["phase"], s["w"], s["ell"], s["var"], "df", s["B"])
    (f[0] * t(g["g"], device="cuda")).sum().backward()
    got = [x.grad[0], s["Z"].grad, s["nu"].grad[0], s["ell"].grad, s["var"].grad, s["B"].grad[0]]
    for nm, a, b in zip(("dx", "dZ", "dnu", "dell", "dvar", "dB"), got, want):
        e = rel(a, b)
        print("%s field-bwd %s: %.2e" % (name, nm, e))
        assert e < GRAD_TOL, (nm, e)
    assert rel(x.grad[0], g["field_dx"]) < 5 * GRAD_TOL  # the reference's own fp32 dx


@pytest.mark.parametrize("name", DF_CASES)
@pytest.mark.parametrize("method", ["euler", "midpoint", "rk4"])
def test_rollout_forward(name, method):
    g = load_golden(name)
    m = g["meta"]
    c64 = oracle_cache(g)
    s = gpu_sample(c64)
    traj = _rollout(s, t(g["z0"], device="cuda"), t(g["ts"], device="cuda"), method)
    assert traj.shape == (1, m["N"], m["T"], m["D_in"])
    truth = OF.rollout(t(g["z0"], torch.float64), t(g["ts"], torch.float64), c64, 1, method)
    e_truth = rel(traj[0], truth)
    print("%s rollout %s: new-vs-fp64 %.2e" % (name, method, e_truth))
    assert e_truth < TRAJ_TOL
    assert torch.equal(traj[0, :, 0].cpu(), t(g["z0"]))
    if method != "midpoint":
        assert rel(traj[0], g["traj_" + method]) < TRAJ_TOL


@pytest.mark.parametrize("name", DF_CASES)
@pytest.mark.parametrize("method", ["euler", "rk4"])
def test_rollout_backward_kernel_level(name, method):
    g = load_golden(name)
    c = _leaf_cache(g)
    z64 = t(g["z0"], torch.float64).requires_grad_(True)
    G = t(g["G"], torch.float64)
    loss = (OF.rollout(z64, t(g["ts"], torch.float64), c, 1, method) * G).sum()
    want = torch.autograd.grad(loss, [z64] + [c[k] for k in LEAVES])
    s = gpu_sample(c)
    for k in LEAVES:
        s[k].requires_grad_(True)
    z0 = t(g["z0"], device="cuda").requires_grad_(True)
    traj = _rollout(s, z0, t(g["ts"], device="cuda"), method)
    (traj[0] * t(g["G"], device="cuda")).sum().backward()
    got = [z0.grad, s["Z"].grad, s["nu"].grad[0], s["ell"].grad, s["var"].grad, s["B"].grad[0]]
    for nm, a, b in zip(("dz0", "dZ", "dnu", "dell", "dvar", "dB"), got, want):
        e = rel(a, b)
        print("%s rollout-bwd %s %s: %.2e" % (name, method, nm, e))
        assert e < GRAD_TOL, (nm, e)


@pytest.mark.parametrize("D,M,S,N", [(2, 7, 5, 100), (3, 9, 16, 257), (5, 33, 31, 1000), (7, 20, 12, 300), (8, 64, 32, 4100)])
def test_shapes_odd_sizes(D, M, S, N):
    """every compiled D, odd M and S (padded pairs), ragged N: field + VJP + parameter gradients vs the fp64 oracle."""
    rs = np.random.RandomState(D * 1000 + M)
    f64 = lambda a: torch.tensor(a, dtype=torch.float64)
    c = dict(variant="df", Z=f64(rs.normal(size=(M, D))), ell=f64(1.5 + rs.uniform(size=(D, D))), var=f64(0.5 + rs.uniform(size=D)),
             nu=f64(rs.normal(size=(M * D, 1))), eps=f64(rs.normal(size=(D, S, D))), phase=f64(rs.uniform(size=(1, S, D)) * 2 * np.pi),
             w=f64(rs.normal(size=(2 * S, D))))
    for k in ("Z", "ell", "var", "nu"):
        c[k].requires_grad_(True)
    c["omega"] = OF.make_omega(c["eps"], c["ell"], "df")
    c["B"] = OF.df_B(c["omega"]).detach().clone().requires_grad_(True)
    x64 = f64(1.5 * rs.normal(size=(N, D))).requires_grad_(True)
    gout = f64(rs.normal(size=(N, D)))
    f64v = OF.field(x64, c)
    want = torch.autograd.grad((f64v * gout).sum(), [x64] + [c[k] for k in LEAVES])
    s = gpu_sample(c)
    for k in LEAVES:
        s[k].requires_grad_(True)
    x = x64.detach().float().cuda()[None].requires_grad_(True)
    f, _ = _gp().gp_field(x, s["Z"], s["nu"], s["eps"], s["phase"], s["w"], s["ell"], s["var"], "df", s["B"])
    assert rel(f[0], f64v) < FIELD_TOL
    (f[0] * gout.float().cuda()).sum().backward()
    got = [x.grad[0], s["Z"].grad, s["nu"].grad[0], s["ell"].grad, s["var"].grad, s["B"].grad[0]]
    for nm, a, b in zip(("dx", "dZ", "dnu", "dell", "dvar", "dB"), got, want):
        e = rel(a, b)
        assert e < GRAD_TOL, (nm, e)


def test_divergence_free_property():
    """the update part with uniform lengthscales is a divergence-free field: trace of its Jacobian vanishes
    (checked through the VJP with unit upstream vectors); holds at any size -- run at 50k states."""
    g = load_golden("df_o1")
    c = oracle_cache(g)
    s = gpu_sample(c)
    s["w"] = torch.zeros_like(s["w"])          # drop the prior part
    D, N = 6, 50000
    x = (1.5 * torch.randn(1, N, D, device="cuda")).requires_grad_(True)
    f, _ = _gp().gp_field(x, s["Z"], s["nu"], s["eps"], s["phase"], s["w"], s["ell"], s["var"], "df", s["B"])
    div = torch.zeros(N, device="cuda")
    for k in range(D):
        e = torch.zeros(1, N, D, device="cuda")
        e[..., k] = 1.0
        (gk,) = torch.autograd.grad(f, x, e, retain_graph=True)
        div += gk[0, :, k]
    scale = f.detach().abs().mean().item()
    assert div.abs().max().item() < 1e-3 * max(scale, 1e-3), (div.abs().max().item(), scale)


def test_batched_samples():
    g = load_golden("df_o1_pert")
    m = g["meta"]
    c = oracle_cache(g)
    s1 = gpu_sample(c)
    L = 3
    rs = np.random.RandomState(0)
    s = dict(s1)
    for k in ("eps", "phase", "w", "nu", "B"):
        s[k] = torch.cat([s1[k]] + [s1[k] * float(1.0 + 0.1 * rs.normal()) for _ in range(L - 1)], 0).contiguous()
    z0 = t(g["z0"], device="cuda").requires_grad_(True)
    ts = t(g["ts"], device="cuda")
    G = torch.randn(L, m["N"], m["T"], m["D_in"], device="cuda")
    traj = _rollout(s, z0, ts, "rk4")
    (traj * G).sum().backward()
    dz_sum = torch.zeros_like(z0)
    for l in range(L):
        sl = {k: (v[l:l + 1] if k in ("eps", "phase", "w", "nu", "B") else v) for k, v in s.items()}
        zl = t(g["z0"], device="cuda").requires_grad_(True)
        tl = _rollout(sl, zl, ts, "rk4")
        assert torch.equal(tl[0], traj[l])
        (tl[0] * G[l]).sum().backward()
        dz_sum += zl.grad
    assert rel(z0.grad, dz_sum) < 1e-6
