"""CPU: the oracle restatement against the golden vectors frozen from the live reference."""
import numpy as np
import pytest
import torch

from oracle import field as OF
from oracle import solvers
from helpers import ALL_CASES, load_golden, oracle_cache, rel, t


@pytest.mark.parametrize("name", ALL_CASES)
def test_field_matches_reference(name):
    g = load_golden(name)
    c32 = oracle_cache(g, torch.float32)
    c64 = oracle_cache(g, torch.float64)
    x = g["x"]
    f32 = OF.field(t(x), c32)
    f64 = OF.field(t(x, torch.float64), c64)
    # like-for-like fp32 port and fp64 truth both sit at the reference's fp32 noise floor (shared nu)
    assert rel(f32, g["field_f"]) < 2e-5
    assert rel(g["field_f"], f64) < 2e-5
    assert rel(OF.prior(t(x, torch.float64), c64), g["field_fprior"]) < 1e-5


@pytest.mark.parametrize("name", ALL_CASES)
def test_nu_and_kl(name):
    g = load_golden(name)
    c64 = oracle_cache(g, torch.float64, shared_nu=False)
    # nu goes through an ill-conditioned Cholesky: compare through f_update, which is what the path uses
    c_ref = oracle_cache(g, torch.float64, shared_nu=True)
    x = t(g["x"], torch.float64)
    assert rel(OF.update(x, c64), OF.update(x, c_ref)) < 5e-4
    M = g["meta"]["M"]
    kl = OF.kl_whitened(t(g["p_Um"], torch.float64), OF.tril_from_packed(t(g["p_Us_sqrt"], torch.float64), M))
    assert abs(kl.item() - float(g["kl"])) < 1e-4 * abs(float(g["kl"]))


@pytest.mark.parametrize("name", ALL_CASES)
@pytest.mark.parametrize("method", ["euler", "rk4"])
def test_rollout_matches_reference(name, method):
    g = load_golden(name)
    m = g["meta"]
    c64 = oracle_cache(g, torch.float64)
    traj = OF.rollout(t(g["z0"], torch.float64), t(g["ts"], torch.float64), c64, m["order"], method)
    assert rel(traj, g["traj_" + method]) < 1e-4
    stages = solvers.STAGES[method]
    assert float(g["nevals_" + method]) == (m["T"] - 1) * stages


@pytest.mark.parametrize("name", ["rbf_dimwise_o1", "rbf_dimwise_o2", "df_d4"])
def test_rollout_gradients_match_reference(name):
    """autograd through the oracle (fp64, own nu) reproduces the reference's leaf gradients."""
    g = load_golden(name)
    m = g["meta"]
    M = m["M"]
    dt = torch.float64
    leaves = {k: t(g["p_" + k], dt).requires_grad_(True) for k in ("raw_ell", "raw_var", "Z", "Um", "Us_sqrt")}
    z0 = t(g["z0"], dt).requires_grad_(True)
    draws = dict(w=t(g["draw_w"], dt), eps=t(g["draw_eps"], dt), phase01=t(g["draw_phase01"], dt), eps_u=t(g["draw_eps_u"], dt))
    Lq = OF.tril_from_packed(leaves["Us_sqrt"], M)
    c = OF.build_cache(m["variant"], leaves["Z"], leaves["raw_ell"], leaves["raw_var"], leaves["Um"], Lq, draws)
    traj = OF.rollout(z0, t(g["ts"], dt), c, m["order"], "rk4")
    loss = (traj * t(g["G"], dt)).sum() + OF.kl_whitened(leaves["Um"], Lq)
    grads = torch.autograd.grad(loss, [z0] + list(leaves.values()))
    assert rel(grads[0], g["roll_rk4_dz0"]) < 2e-3
    for k, gv in zip(leaves.keys(), grads[1:]):
        assert rel(gv, g["roll_rk4_d" + k]) < 5e-3, k


def test_solver_restatement_properties():
    """fixed-grid solvers: exact on dy/dt = const, rk4 4th order on dy/dt = -y, grid = output grid."""
    ts = 0.1 * torch.arange(16, dtype=torch.float64)
    y0 = torch.ones(3, 2, dtype=torch.float64)
    for method in solvers.FIXED_GRID_METHODS:
        out = solvers.odeint(lambda tt, y: torch.full_like(y, 2.0), y0, ts, method=method)
        assert out.shape == (16, 3, 2)
        assert torch.allclose(out[-1], y0 + 2.0 * ts[-1])
    exact = torch.exp(-ts[-1])
    errs = []
    for n in (16, 31):
        tt = torch.linspace(0, 1.5, n, dtype=torch.float64)
        out = solvers.odeint(lambda s, y: -y, y0, tt, method="rk4")
        errs.append(abs(out[-1, 0, 0].item() - np.exp(-1.5)))
    assert errs[0] / errs[1] > 12  # ~2^4
    with pytest.raises(ValueError):
        solvers.odeint(lambda s, y: -y, y0, ts, method="dopri5")


@pytest.mark.parametrize("name", ALL_CASES)
def test_port_is_bit_identical_to_the_reference(name):
    """oracle/port_fp32.py (bench.py's fallback CPU baseline and the fp32 comparison arm of tests/test_gpu_baseline_shapes.py) keeps the
    reference's own op sequence: on the golden inputs its field and trajectories equal the reference's outputs BIT FOR BIT."""
    from oracle import port_fp32 as PORT
    g = load_golden(name)
    m = g["meta"]
    c = oracle_cache(g, torch.float32)
    c = {k: (v.detach() if torch.is_tensor(v) else v) for k, v in c.items()}
    with torch.no_grad():
        f = PORT.field(t(g["x"]), c)
        assert np.array_equal(f.numpy(), g["field_f"]), float(np.abs(f.numpy() - g["field_f"]).max())
        for method in ("euler", "rk4"):
            traj = PORT.rollout(t(g["z0"]), t(g["ts"]), c, m["order"], method)
            assert np.array_equal(traj.numpy(), g["traj_" + method]), (method, float(np.abs(traj.numpy() - g["traj_" + method]).max()))
