"""GPU: the drop-in modules (SVGP_Layer / Flow with the reference's constructor signatures and
state_dict keys) end to end against the golden vectors of the live reference: own nu (cuSOLVER Cholesky
on the GPU instead of LAPACK), leaf gradients through build_cache."""
import numpy as np
import pytest
import torch

from helpers import ALL_CASES, RBF_CASES, load_golden, rel, t
from oracle import field as OF

pytestmark = pytest.mark.gpu


class Draws:
    """Feeds the golden draws to the drop-in's host RNG helpers in the reference's draw order."""

    def __init__(self, g):
        self.q = [g["draw_w"], g["draw_eps"], g["draw_phase01"], g["draw_eps_u"]]
        self.i = 0

    def __call__(self, shape, seed=None):
        v = self.q[self.i % 4]
        self.i += 1
        assert tuple(v.shape) == tuple(shape), (v.shape, shape)
        return torch.tensor(v)


def build_flow(g, method, monkeypatch):
    from gpode_b200.core import kernels as K
    from gpode_b200.core import svpy as SV
    from gpode_b200.core.flow import Flow
    m = g["meta"]
    np.random.seed(0)
    gp = SV.SVGP_Layer(D_in=m["D_in"], D_out=m["D_out"], M=m["M"], S=m["S"], q_diag=False, dimwise=m["dimwise"], device="cuda",
                       kernel=m["kernel"])
    flow = Flow(diffeq=gp, order=m["order"], solver=method, use_adjoint=False)
    sd = {"odefunc._num_evals": torch.tensor(0.),
          "odefunc.diffeq.kern.unconstrained_lengthscales": t(g["p_raw_ell"]),
          "odefunc.diffeq.kern.unconstrained_variance": t(g["p_raw_var"]),
          "odefunc.diffeq.inducing_loc.optvar": t(g["p_Z"]),
          "odefunc.diffeq.Um.optvar": t(g["p_Um"]),
          "odefunc.diffeq.Us_sqrt.optvar": t(g["p_Us_sqrt"])}
    flow.load_state_dict(sd, strict=True)   # the reference's checkpoint keys, verbatim
    d = Draws(g)
    monkeypatch.setattr(K, "sample_normal", d)
    monkeypatch.setattr(K, "sample_uniform", d)
    monkeypatch.setattr(SV, "sample_normal", d)
    return flow, gp


LEAVES = ("raw_ell", "raw_var", "Z", "Um", "Us_sqrt")


def fp64_truth(g, method):
    """The whole path in float64 with its OWN nu (fp64 K(Z,Z), Cholesky, solves -- oracle.field.build_cache without nu_override), the
    rollout, the KL and autograd down to the leaf parameters: the truth both fp32 implementations are measured against."""
    m = g["meta"]
    f64 = torch.float64
    leaves = {k: t(g["p_" + k], f64).requires_grad_(True) for k in LEAVES}
    Lq = OF.tril_from_packed(leaves["Us_sqrt"], m["M"])
    draws = dict(w=t(g["draw_w"], f64), eps=t(g["draw_eps"], f64), phase01=t(g["draw_phase01"], f64), eps_u=t(g["draw_eps_u"], f64))
    c = OF.build_cache(m["variant"], leaves["Z"], leaves["raw_ell"], leaves["raw_var"], leaves["Um"], Lq, draws)
    z0 = t(g["z0"], f64).requires_grad_(True)
    traj = OF.rollout(z0, t(g["ts"], f64), c, m["order"], method)
    loss = (traj * t(g["G"], f64)).sum() + OF.kl_whitened(leaves["Um"], Lq)
    grads = torch.autograd.grad(loss, [z0] + [leaves[k] for k in LEAVES])
    return traj.detach(), dict(zip(("z0",) + LEAVES, grads))


@pytest.mark.parametrize("name", ALL_CASES)
@pytest.mark.parametrize("method", ["euler", "rk4"])
def test_flow_end_to_end(name, method, monkeypatch):
    g = load_golden(name)
    m = g["meta"]
    flow, gp = build_flow(g, method, monkeypatch)
    z0 = t(g["z0"], device="cuda").requires_grad_(True)
    ts = t(g["ts"], device="cuda")
    traj = flow(z0, ts)
    assert traj.shape == (m["N"], m["T"], m["D_in"])
    assert flow.num_evals() == float(g["nevals_" + method])
    e = rel(traj, g["traj_" + method])
    print("%s %s traj (own nu): %.2e" % (name, method, e))
    assert e < 1e-4
    kl = flow.kl()
    assert abs(kl.item() - float(g["kl"])) < 1e-5 * abs(float(g["kl"]))
    loss = (traj * t(g["G"], device="cuda")).sum() + kl
    loss.backward()
    got = {"z0": z0.grad, "raw_ell": gp.kern.unconstrained_lengthscales.grad, "raw_var": gp.kern.unconstrained_variance.grad,
           "Z": gp.inducing_loc.optvar.grad, "Um": gp.Um.optvar.grad, "Us_sqrt": gp.Us_sqrt.optvar.grad}
    # north_star bar 3 (parameter gradients to rel 1e-4) as three numbers per leaf: the reference's own fp32
    # gradients carry the noise of its fp32 Cholesky of an ill-conditioned K(Z,Z) (cond 1e4..1e6), so "equal to
    # the reference to 1e-4" is only meaningful where the reference itself is that close to the truth.  Required of the new path:
    # at least as close to the fp64 truth as 1e-4, or as the reference is.
    traj64, want = fp64_truth(g, method)
    e_new, e_ref = rel(traj, traj64), rel(g["traj_" + method], traj64)
    print("%s %s traj: new-vs-fp64 %.2e  ref-vs-fp64 %.2e  new-vs-ref %.2e" % (name, method, e_new, e_ref, rel(traj, g["traj_" + method])))
    assert e_new <= max(1e-4, e_ref)
    for k, v in got.items():
        ref = g["roll_%s_d%s" % (method, k)]
        e_new, e_ref, e_nr = rel(v, want[k]), rel(ref, want[k]), rel(v, ref)
        print("%s %s d%s: new-vs-fp64 %.2e  ref-vs-fp64 %.2e  new-vs-ref %.2e" % (name, method, k, e_new, e_ref, e_nr))
        assert e_new <= max(1e-4, e_ref), (k, e_new, e_ref)
        assert e_nr <= max(1e-4, 2.0 * (e_new + e_ref)), (k, e_nr)      # and the two fp32 paths differ by no more than their noise


@pytest.mark.parametrize("name", ["rbf_dimwise_o1", "rbf_dimwise_o2"])
def test_layer_forward_and_batched_flow(name, monkeypatch):
    g = load_golden(name)
    m = g["meta"]
    flow, gp = build_flow(g, "rk4", monkeypatch)
    gp.build_cache()
    f = gp(t(g["x"], device="cuda"))
    assert rel(f, g["field_f"]) < 5e-5
    # torchdiffeq-style callable contract
    sv = t(g["z0"], device="cuda")
    dy = flow.odefunc(torch.tensor(0.0), sv)
    assert dy.shape == sv.shape
    # batched MC samples: same draws fed twice -> both samples equal the single-sample golden trajectory
    trajL = flow.forward_samples(t(g["z0"], device="cuda"), t(g["ts"], device="cuda"), 2)
    assert trajL.shape == (2, m["N"], m["T"], m["D_in"])
    assert rel(trajL[0], g["traj_rk4"]) < 1e-4 and rel(trajL[1], g["traj_rk4"]) < 1e-4


def test_unsupported_solver_raises(monkeypatch):
    g = load_golden("rbf_dimwise_o1")
    flow, _ = build_flow(g, "dopri5", monkeypatch)
    with pytest.raises(NotImplementedError):
        flow(t(g["z0"], device="cuda"), t(g["ts"], device="cuda"))


@pytest.mark.parametrize("name", ["rbf_dimwise_o1", "df_o1"])
def test_single_time_point_rollout_has_zero_parameter_gradients(name, monkeypatch):
    """T = 1 (e.g. T_custom = 1): no step is taken, traj == z0, dz0 == dtraj and every parameter gradient is exactly zero --
    written by the library, not left as uninitialised memory (round-1 advisor finding)."""
    g = load_golden(name)
    flow, gp = build_flow(g, "rk4", monkeypatch)
    z0 = t(g["z0"], device="cuda").requires_grad_(True)
    traj = flow(z0, t(g["ts"][:1], device="cuda"))
    assert traj.shape[1] == 1 and torch.equal(traj[:, 0], z0.detach())
    G = torch.randn_like(traj)
    (traj * G).sum().backward()
    assert torch.equal(z0.grad, G[:, 0])
    for p in (gp.kern.unconstrained_lengthscales, gp.kern.unconstrained_variance, gp.inducing_loc.optvar, gp.Um.optvar, gp.Us_sqrt.optvar):
        assert p.grad is None or (torch.isfinite(p.grad).all() and p.grad.abs().max().item() == 0.0)


def test_non_positive_definite_gram_raises_like_the_reference(monkeypatch):
    """torch.linalg.cholesky raises in the reference (kernels.py:163); the fused setup reports the failed pivot in a device flag that
    SVGP_Layer checks after build_cache (check_cholesky=True) or on request (cholesky_ok())."""
    g = load_golden("rbf_dimwise_o1")
    flow, gp = build_flow(g, "rk4", monkeypatch)
    with torch.no_grad():
        gp.inducing_loc.optvar[3, 2] = float("nan")
    with pytest.raises(torch.linalg.LinAlgError):
        gp.build_cache()
    gp.check_cholesky = False
    gp.build_cache()                      # no host sync, no exception ...
    assert gp.cholesky_ok() is False      # ... but the status is there to be asked for


def test_stale_cache_is_not_served_after_batched_samples(monkeypatch):
    """SVGP_Layer.forward after Flow.forward_samples must evaluate the LAST batched sample (what the kernel attributes hold), not a
    cache left behind by an earlier build_cache() (round-1 advisor finding)."""
    g = load_golden("rbf_dimwise_o1")
    flow, gp = build_flow(g, "rk4", monkeypatch)
    x = t(g["x"], device="cuda")
    gp.build_cache()
    f_first = gp(x).detach().clone()
    assert rel(f_first, g["field_f"]) < 5e-5
    # a different sample: scale the inducing means, then draw batched samples (same host draws)
    with torch.no_grad():
        gp.Um.optvar.mul_(3.0)
    flow.forward_samples(t(g["z0"], device="cuda"), t(g["ts"], device="cuda"), 2)
    assert gp._cache is None
    f_after = gp(x).detach()
    assert rel(f_after, f_first) > 1e-2                      # not the stale sample
    gp.build_cache()
    assert rel(gp(x), f_after) < 1e-5                        # the same draws and parameters reproduce it


@pytest.mark.parametrize("name", ["rbf_dimwise_o1", "df_o1"])
def test_q_diag_layer_runs_the_fused_setup(name, monkeypatch):
    """q_diag=True (reference svpy.py:79-82,96-97,152-167): diagonal q(u).  Only the inducing sample and the KL differ (both elementwise,
    M x D_out); K(Z,Z), the factorisation and the whitened solves run on the same setup kernels as the full-covariance layer.  Trajectories,
    KL and every leaf gradient against the fp64 oracle run with its own nu (no golden of the reference exists for q_diag)."""
    from gpode_b200.core import kernels as K
    from gpode_b200.core import svpy as SV
    from gpode_b200.core.flow import Flow
    g = load_golden(name)
    m = g["meta"]
    np.random.seed(1)
    gp = SV.SVGP_Layer(D_in=m["D_in"], D_out=m["D_out"], M=m["M"], S=m["S"], q_diag=True, dimwise=m["dimwise"], device="cuda", kernel=m["kernel"])
    assert gp._fused_setup() or gp._df_fused_setup()
    flow = Flow(diffeq=gp, order=m["order"], solver="rk4", use_adjoint=False)
    raw_s = np.random.RandomState(5).normal(size=(m["M"], m["D_out"])) * 0.3 - 1.0     # unconstrained diagonal scales
    with torch.no_grad():
        gp.kern.unconstrained_lengthscales.copy_(t(g["p_raw_ell"]))
        gp.kern.unconstrained_variance.copy_(t(g["p_raw_var"]))
        gp.inducing_loc.optvar.copy_(t(g["p_Z"]))
        gp.Um.optvar.copy_(t(g["p_Um"]))
        gp.Us_sqrt.optvar.copy_(t(raw_s))
    d = Draws(g)
    monkeypatch.setattr(K, "sample_normal", d)
    monkeypatch.setattr(K, "sample_uniform", d)
    monkeypatch.setattr(SV, "sample_normal", d)
    z0 = t(g["z0"], device="cuda").requires_grad_(True)
    traj = flow(z0, t(g["ts"], device="cuda"))
    kl = flow.kl()
    ((traj * t(g["G"], device="cuda")).sum() + kl).backward()
    # fp64 truth
    f64 = torch.float64
    lv = {k: t(g["p_" + k], f64).requires_grad_(True) for k in ("raw_ell", "raw_var", "Z", "Um")}
    rs = t(raw_s, f64).requires_grad_(True)
    sq = torch.nn.functional.softplus(rs) + 1e-12
    draws = dict(w=t(g["draw_w"], f64), eps=t(g["draw_eps"], f64), phase01=t(g["draw_phase01"], f64), eps_u=t(g["draw_eps_u"], f64))
    c = OF.build_cache(m["variant"], lv["Z"], lv["raw_ell"], lv["raw_var"], lv["Um"], sq, draws, q_diag=True)
    z64 = t(g["z0"], f64).requires_grad_(True)
    want = OF.rollout(z64, t(g["ts"], f64), c, m["order"], "rk4")
    kl64 = OF.kl_whitened(lv["Um"], sq, q_diag=True)
    gr = torch.autograd.grad((want * t(g["G"], f64)).sum() + kl64, [z64, lv["raw_ell"], lv["raw_var"], lv["Z"], lv["Um"], rs])
    got = [z0.grad, gp.kern.unconstrained_lengthscales.grad, gp.kern.unconstrained_variance.grad, gp.inducing_loc.optvar.grad, gp.Um.optvar.grad,
           gp.Us_sqrt.optvar.grad]
    errs = [rel(traj, want), abs(kl.item() - kl64.item()) / abs(kl64.item())] + [rel(a, b) for a, b in zip(got, gr)]
    print(name, "q_diag: traj %.2e kl %.2e dz0 %.2e dell %.2e dvar %.2e dZ %.2e dUm %.2e dUs %.2e" % tuple(errs))
    assert max(errs) < 1e-4
