"""GPU: the two remaining solver modes of Flow.

  * use_adjoint=True (reference core/flow.py:76, torchdiffeq.odeint_adjoint): functional.GPRolloutAdjoint -- fused forward launch,
    backward = torchdiffeq's augmented system integrated backwards interval by interval with one step of the same method, every
    stage one field evaluation + one field VJP on the CUDA kernels -- against oracle/solvers.py odeint_adjoint (the published
    algorithm restated; torchdiffeq itself is absent: parity unpinned at this boundary like the forward solvers) on the fp64 oracle
    field.  Bars: the kernel-level gradient bar 1e-4 (measured ~1e-6).
  * ts_dense_scale (main.py:83, misc/torch_utils.py:54-61): integration on the densified grid, states returned at ts."""
import numpy as np
import pytest
import torch

from oracle import field as OF
from oracle import solvers
from helpers import gpu_sample, load_golden, oracle_cache, rel, t

pytestmark = pytest.mark.gpu
GRAD_TOL = 1e-4


def _gp():
    import gpode_b200
    return gpode_b200


@pytest.mark.parametrize("name", ["rbf_dimwise_o1", "rbf_shared_o1", "rbf_dimwise_o2", "df_o1"])
@pytest.mark.parametrize("method", ["euler", "midpoint", "rk4"])
def test_adjoint_mode_matches_restated_torchdiffeq_adjoint(name, method):
    g = load_golden(name)
    m = g["meta"]
    c = oracle_cache(g, leaves=True)
    df = m["variant"] == "df"
    if df:
        c["B"] = OF.df_B(c["omega"]).detach().clone().requires_grad_(True)
    params = [c["Z"], c["nu"], c["ell"], c["var"]] + ([c["B"]] if df else [])

    def func(tt, y):
        cc = dict(c)
        cc["omega"] = OF.make_omega(c["eps"], c["ell"], m["variant"])         # omega = eps / ell stays inside the graph of every evaluation
        return OF.rhs(y, cc, m["order"])
    z64 = t(g["z0"], torch.float64).requires_grad_(True)
    ts64, G = t(g["ts"], torch.float64), t(g["G"], torch.float64)
    ys = solvers.odeint_adjoint(func, z64, ts64, method=method, adjoint_params=tuple(params))            # (T,N,D)
    (ys.permute(1, 0, 2) * G).sum().backward()
    want = [z64.grad] + [p.grad for p in params]
    s = gpu_sample(c)
    for k in ("Z", "nu", "ell", "var"):
        s[k].requires_grad_(True)
    if df:
        s["B"] = c["B"].detach().float().cuda()[None].requires_grad_(True)
    z0 = t(g["z0"], device="cuda").requires_grad_(True)
    traj = _gp().gp_rollout(z0, t(g["ts"], device="cuda"), s["Z"], s["nu"], s["eps"], s["phase"], s["w"], s["ell"], s["var"], m["variant"], m["order"],
                            method, s["B"], adjoint=True)
    assert rel(traj[0], ys.detach().permute(1, 0, 2)) < 1e-5
    (traj[0] * t(g["G"], device="cuda")).sum().backward()
    got = [z0.grad, s["Z"].grad, s["nu"].grad[0], s["ell"].grad, s["var"].grad] + ([s["B"].grad[0]] if df else [])
    for nm, a, b in zip(("dz0", "dZ", "dnu", "dell", "dvar", "dB"), got, want):
        e = rel(a, b)
        print("%s adjoint %s %s: %.2e" % (name, method, nm, e))
        assert e < GRAD_TOL, (nm, e)
    # and it is a different gradient from the exact one of the discrete solve (O(dt^p) apart), as in the reference
    z1 = t(g["z0"], device="cuda").requires_grad_(True)
    tr = _gp().gp_rollout(z1, t(g["ts"], device="cuda"), s["Z"].detach(), s["nu"].detach(), s["eps"], s["phase"], s["w"], s["ell"].detach(), s["var"].detach(),
                          m["variant"], m["order"], method, None if s["B"] is None else s["B"].detach())
    (tr[0] * t(g["G"], device="cuda")).sum().backward()
    print("%s %s: adjoint vs discrete dz0 %.2e" % (name, method, rel(z0.grad, z1.grad)))
    assert rel(z0.grad, z1.grad) < (0.2 if method == "euler" else 2e-2)


def test_flow_adjoint_and_dense_substepping(monkeypatch):
    from gpode_b200.core.flow import Flow
    from gpode_b200.core.svpy import SVGP_Layer
    np.random.seed(3)
    layer = SVGP_Layer(6, 6, 50, 64, dimwise=True, device="cuda", kernel="RBF")
    z0 = torch.randn(40, 6, device="cuda")
    ts = 0.1 * torch.arange(6, dtype=torch.float, device="cuda")
    flow = Flow(layer, order=1, solver="rk4", use_adjoint=True)
    np.random.seed(7)
    za = z0.clone().requires_grad_(True)
    out = flow(za, ts)
    assert out.shape == (40, 6, 6)
    out.square().sum().backward()
    assert za.grad is not None and torch.isfinite(za.grad).all() and layer.Um.optvar.grad is not None
    # dense sub-stepping: the same function sample (same draws) integrated with 3 sub-steps per interval; oracle on the dense grid
    flow2 = Flow(layer, order=1, solver="euler")
    flow2.ts_dense_scale = 4
    np.random.seed(11)
    zb = z0.clone().requires_grad_(True)
    dense = flow2(zb, ts)
    assert dense.shape == (40, 6, 6) and flow2.num_evals() == 15
    k = layer.kern
    c = dict(variant="rbf_dimwise", Z=layer.inducing_loc().detach().double().cpu(), ell=k.lengthscales.detach().double().cpu(), var=k.variance.detach().double().cpu(),
             nu=k.nu.detach().double().cpu(), phase=k.rff_phase.double().cpu(), w=k.rff_weights.double().cpu())
    c["omega"] = k.rff_omega.detach().double().cpu()
    from gpode_b200.misc.torch_utils import compute_ts_dense
    truth = OF.rollout(z0.double().cpu(), compute_ts_dense(ts.cpu(), 4).double(), c, 1, "euler")[:, ::3]
    assert rel(dense, truth) < 1e-5
    coarse = OF.rollout(z0.double().cpu(), ts.double().cpu(), c, 1, "euler")
    assert rel(dense, coarse) > 1e-4                      # the sub-steps do change the solution
    dense.sum().backward()
    assert torch.isfinite(zb.grad).all()
