"""GPU parity of the RBF vector-field kernels through the C ABI (ctypes -> libgpode.so).

Tolerances (north_star, fp32): per-evaluation field rel 1e-5, T-step trajectories rel 1e-4, gradients
rel 1e-4 -- all norm-wise |a-b|_F/|b|_F with nu shared with the reference.
Kernel-level gradients are checked against autograd through the fp64 oracle with Z, ell, var, nu as
independent leaves (exactly the tensors the C ABI differentiates)."""
import numpy as np
import pytest
import torch

from oracle import field as OF
from helpers import RBF_CASES, gpu_sample, load_golden, oracle_cache, rel, t

pytestmark = pytest.mark.gpu

FIELD_TOL = 1e-5
TRAJ_TOL = 1e-4
GRAD_TOL = 1e-4


def _gp():
    import gpode_b200
    return gpode_b200


def _field(s, x, variant):
    f, fp = _gp().gp_field(x, s["Z"], s["nu"], s["eps"], s["phase"], s["w"], s["ell"], s["var"], variant, s["B"])
    return f, fp


@pytest.mark.parametrize("name", RBF_CASES)
def test_field_forward(name):
    g = load_golden(name)
    c64 = oracle_cache(g)
    s = gpu_sample(c64)
    x = t(g["x"], device="cuda")[None]
    f, fp = _field(s, x, g["meta"]["variant"])
    truth = OF.field(t(g["x"], torch.float64), c64)
    e_ref, e_truth, e_floor = rel(f[0], g["field_f"]), rel(f[0], truth), rel(g["field_f"], truth)
    print("%s field: new-vs-ref %.2e new-vs-fp64 %.2e ref-vs-fp64 %.2e" % (name, e_ref, e_truth, e_floor))
    assert e_truth < FIELD_TOL, (e_ref, e_truth, e_floor)
    assert e_ref < FIELD_TOL + e_floor, (e_ref, e_truth, e_floor)
    assert rel(fp[0], OF.prior(t(g["x"], torch.float64), c64)) < FIELD_TOL


@pytest.mark.parametrize("name", RBF_CASES)
def test_field_backward_kernel_level(name):
    g = load_golden(name)
    variant = g["meta"]["variant"]
    c = oracle_cache(g, leaves=True)
    x64 = t(g["x"], torch.float64).requires_grad_(True)
    gout = t(g["g"], torch.float64)
    loss = (OF.field(x64, c) * gout).sum()
    want = torch.autograd.grad(loss, [x64, c["Z"], c["nu"], c["ell"], c["var"]])
    s = gpu_sample(c)
    for k in ("Z", "nu", "ell", "var"):
        s[k].requires_grad_(True)
    x = t(g["x"], device="cuda")[None].requires_grad_(True)
    f, _ = _field(s, x, variant)
    (f[0] * t(g["g"], device="cuda")).sum().backward()
    got = [x.grad[0], s["Z"].grad, s["nu"].grad[0], s["ell"].grad, s["var"].grad]
    for nm, a, b in zip(("dx", "dZ", "dnu", "dell", "dvar"), got, want):
        e = rel(a, b)
        print("%s field-bwd %s: %.2e" % (name, nm, e))
        assert e < GRAD_TOL, (nm, e)
    assert rel(x.grad[0], g["field_dx"]) < 5 * GRAD_TOL  # the reference's own fp32 dx


@pytest.mark.parametrize("name", RBF_CASES)
@pytest.mark.parametrize("method", ["euler", "midpoint", "rk4"])
def test_rollout_forward(name, method):
    g = load_golden(name)
    m = g["meta"]
    c64 = oracle_cache(g)
    s = gpu_sample(c64)
    z0, ts = t(g["z0"], device="cuda"), t(g["ts"], device="cuda")
    traj = _gp().gp_rollout(z0, ts, s["Z"], s["nu"], s["eps"], s["phase"], s["w"], s["ell"], s["var"], m["variant"], m["order"],
                            method, s["B"])
    assert traj.shape == (1, m["N"], m["T"], m["D_in"])
    truth = OF.rollout(t(g["z0"], torch.float64), t(g["ts"], torch.float64), c64, m["order"], method)
    e_truth = rel(traj[0], truth)
    print("%s rollout %s: new-vs-fp64 %.2e" % (name, method, e_truth))
    assert e_truth < TRAJ_TOL
    assert torch.equal(traj[0, :, 0].cpu(), t(g["z0"]))
    if method != "midpoint":
        assert rel(traj[0], g["traj_" + method]) < TRAJ_TOL


@pytest.mark.parametrize("name", RBF_CASES)
@pytest.mark.parametrize("method", ["euler", "midpoint", "rk4"])
def test_rollout_backward_kernel_level(name, method):
    g = load_golden(name)
    m = g["meta"]
    c = oracle_cache(g, leaves=True)
    z64 = t(g["z0"], torch.float64).requires_grad_(True)
    G = t(g["G"], torch.float64)
    loss = (OF.rollout(z64, t(g["ts"], torch.float64), c, m["order"], method) * G).sum()
    want = torch.autograd.grad(loss, [z64, c["Z"], c["nu"], c["ell"], c["var"]])
    s = gpu_sample(c)
    for k in ("Z", "nu", "ell", "var"):
        s[k].requires_grad_(True)
    z0 = t(g["z0"], device="cuda").requires_grad_(True)
    traj = _gp().gp_rollout(z0, t(g["ts"], device="cuda"), s["Z"], s["nu"], s["eps"], s["phase"], s["w"], s["ell"], s["var"],
                            m["variant"], m["order"], method, s["B"])
    (traj[0] * t(g["G"], device="cuda")).sum().backward()
    got = [z0.grad, s["Z"].grad, s["nu"].grad[0], s["ell"].grad, s["var"].grad]
    for nm, a, b in zip(("dz0", "dZ", "dnu", "dell", "dvar"), got, want):
        e = rel(a, b)
        print("%s rollout-bwd %s %s: %.2e" % (name, method, nm, e))
        assert e < GRAD_TOL, (nm, e)


def test_batched_samples_and_per_sample_z0():
    """L function samples in one launch == L separate launches; shared z0 gradient is the sum over samples."""
    g = load_golden("rbf_dimwise_o1_pert")
    m = g["meta"]
    c = oracle_cache(g)
    s1 = gpu_sample(c)
    L = 3
    rs = np.random.RandomState(0)
    s = dict(s1)
    for k in ("eps", "phase", "w", "nu"):
        s[k] = torch.cat([s1[k]] + [s1[k] * float(1.0 + 0.1 * rs.normal()) for _ in range(L - 1)], 0).contiguous()
    gp = _gp()
    z0 = t(g["z0"], device="cuda").requires_grad_(True)
    ts = t(g["ts"], device="cuda")
    G = torch.randn(L, m["N"], m["T"], m["D_in"], device="cuda")
    traj = gp.gp_rollout(z0, ts, s["Z"], s["nu"], s["eps"], s["phase"], s["w"], s["ell"], s["var"], m["variant"], 1, "rk4")
    (traj * G).sum().backward()
    dz_sum = torch.zeros_like(z0)
    for l in range(L):
        zl = t(g["z0"], device="cuda").requires_grad_(True)
        tl = gp.gp_rollout(zl, ts, s["Z"], s["nu"][l:l + 1], s["eps"][l:l + 1], s["phase"][l:l + 1], s["w"][l:l + 1], s["ell"],
                           s["var"], m["variant"], 1, "rk4")
        assert torch.equal(tl[0], traj[l])
        (tl[0] * G[l]).sum().backward()
        dz_sum += zl.grad
    assert rel(z0.grad, dz_sum) < 1e-6
    # per-sample initial states
    z0L = torch.stack([t(g["z0"], device="cuda") * (1 + 0.1 * l) for l in range(L)]).requires_grad_(True)
    trajL = gp.gp_rollout(z0L, ts, s["Z"], s["nu"], s["eps"], s["phase"], s["w"], s["ell"], s["var"], m["variant"], 1, "euler")
    assert torch.equal(trajL[0], gp.gp_rollout(z0L[0].detach(), ts, s["Z"], s["nu"][:1], s["eps"][:1], s["phase"][:1], s["w"][:1],
                                               s["ell"], s["var"], m["variant"], 1, "euler")[0])
    trajL.sum().backward()
    assert z0L.grad.shape == z0L.shape


@pytest.mark.parametrize("N", [1, 31, 257, 1000, 70001])
def test_ragged_and_large_state_counts(N):
    """state counts that do not fill a CTA / warp, and one large enough to use the 2-states-per-thread shape:
    a random subset is checked against the fp64 oracle."""
    g = load_golden("rbf_dimwise_o1")
    c = oracle_cache(g)
    s = gpu_sample(c)
    rs = np.random.RandomState(N)
    x = torch.tensor(1.5 * rs.normal(size=(1, N, 6)), dtype=torch.float32, device="cuda")
    f, _ = _field(s, x, "rbf_dimwise")
    idx = np.unique(np.concatenate([[0, N - 1], rs.randint(0, N, size=min(N, 200))]))
    truth = OF.field(x[0, idx].double().cpu(), c)
    assert rel(f[0, idx], truth) < FIELD_TOL
    ts = 0.1 * torch.arange(5, dtype=torch.float, device="cuda")
    traj = _gp().gp_rollout(x[0], ts, s["Z"], s["nu"], s["eps"], s["phase"], s["w"], s["ell"], s["var"], "rbf_dimwise", 1, "rk4")
    truth = OF.rollout(x[0, idx].double().cpu(), ts.double().cpu(), c, 1, "rk4")
    assert rel(traj[0, idx], truth) < TRAJ_TOL


def test_zero_field_properties():
    """size-independent properties: w = 0 and nu = 0 give f = 0, so the trajectory stays at z0 and dz0 = sum_t G_t."""
    g = load_golden("rbf_dimwise_o1")
    c = oracle_cache(g)
    s = gpu_sample(c)
    s["w"] = torch.zeros_like(s["w"])
    s["nu"] = torch.zeros_like(s["nu"])
    N, T = 4099, 7
    z0 = torch.randn(N, 6, device="cuda").requires_grad_(True)
    ts = 0.1 * torch.arange(T, dtype=torch.float, device="cuda")
    traj = _gp().gp_rollout(z0, ts, s["Z"], s["nu"], s["eps"], s["phase"], s["w"], s["ell"], s["var"], "rbf_dimwise", 1, "rk4")
    assert torch.equal(traj[0], z0.detach()[:, None, :].expand(N, T, 6))
    G = torch.randn(1, N, T, 6, device="cuda")
    (traj * G).sum().backward()
    assert rel(z0.grad, G[0].sum(1)) < 1e-6


def test_errors_are_loud():
    gp = _gp()
    g = load_golden("rbf_dimwise_o1")
    c = oracle_cache(g)
    s = gpu_sample(c)
    x = t(g["x"])[None]  # CPU tensor: no CPU path exists
    with pytest.raises(RuntimeError):
        _field(s, x, "rbf_dimwise")
    with pytest.raises(RuntimeError):  # order 2 needs D_in == 2 D_out
        gp.gp_rollout(t(g["z0"], device="cuda"), t(g["ts"], device="cuda"), s["Z"], s["nu"], s["eps"], s["phase"], s["w"], s["ell"],
                      s["var"], "rbf_dimwise", 2, "rk4")


@pytest.mark.parametrize("D_in,D_out,M,S,N", [(12, 5, 101, 33, 300), (9, 9, 40, 7, 70), (16, 3, 512, 16, 130), (14, 14, 129, 20, 1000)])
def test_wide_inputs_odd_sizes(D_in, D_out, M, S, N):
    """D > 8 (tensor-path parameter gradients, padded input dims, odd M / S, ragged N): field + VJP + parameter gradients
    against the fp64 oracle."""
    rs = np.random.RandomState(D_in * 100 + M)
    f64 = lambda a: torch.tensor(a, dtype=torch.float64)
    c = dict(variant="rbf_dimwise", Z=f64(rs.normal(size=(M, D_in))), ell=f64(1.5 + rs.uniform(size=(D_out, D_in))),
             var=f64(0.5 + rs.uniform(size=D_out)), nu=f64(rs.normal(size=(D_out, M, 1))), eps=f64(rs.normal(size=(D_in, S, D_out))),
             phase=f64(rs.uniform(size=(1, S, D_out)) * 2 * np.pi), w=f64(rs.normal(size=(S, D_out))))
    for k in ("Z", "ell", "var", "nu"):
        c[k].requires_grad_(True)
    c["omega"] = OF.make_omega(c["eps"], c["ell"], "rbf_dimwise")
    x64 = f64(1.5 * rs.normal(size=(N, D_in))).requires_grad_(True)
    gout = f64(rs.normal(size=(N, D_out)))
    f64v = OF.field(x64, c)
    want = torch.autograd.grad((f64v * gout).sum(), [x64, c["Z"], c["nu"], c["ell"], c["var"]])
    s = gpu_sample(c)
    for k in ("Z", "nu", "ell", "var"):
        s[k].requires_grad_(True)
    x = x64.detach().float().cuda()[None].requires_grad_(True)
    f, _ = _field(s, x, "rbf_dimwise")
    assert rel(f[0], f64v) < FIELD_TOL
    (f[0] * gout.float().cuda()).sum().backward()
    got = [x.grad[0], s["Z"].grad, s["nu"].grad[0], s["ell"].grad, s["var"].grad]
    for nm, a, b in zip(("dx", "dZ", "dnu", "dell", "dvar"), got, want):
        e = rel(a, b)
        assert e < GRAD_TOL, (nm, e)


@pytest.fixture(params=["mma", "tc"])
def fwd_kernel(request):
    """the two forward kernels of the D > 8 tensor path: mma.sync (RbfMmaFwdPolicy) and tcgen05 / tensor memory (RbfTcFwdPolicy,
    the default when its 256-unit operand tiles are < 15 % padding), selected through GpodeProblem.flags"""
    with _gp().kernel_flags(_gp().FLAG_FWD_MMA if request.param == "mma" else _gp().FLAG_FWD_TCGEN05):
        yield request.param


@pytest.fixture(params=["mma", "tc"])
def bwd_kernel(request):
    """the two reverse-sweep / parameter-gradient kernel families of the D > 8 tensor path: mma.sync (RbfMmaBwdPolicy, k_rbf_pgrad_mma) and
    tcgen05 / tensor memory (RbfTcBwdPolicy, rbf_bwd_tc.cuh), selected through GpodeProblem.flags"""
    with _gp().kernel_flags(_gp().FLAG_BWD_MMA if request.param == "mma" else _gp().FLAG_BWD_TCGEN05):
        yield request.param


@pytest.mark.parametrize("order,D_in,D_out,M,S,L", [(1, 16, 16, 300, 260, 1), (2, 16, 8, 512, 256, 2), (1, 16, 16, 97, 513, 1), (1, 11, 11, 260, 130, 1),
                                                    (2, 14, 7, 130, 250, 1)])
def test_forward_kernels_ragged_tiles(fwd_kernel, order, D_in, D_out, M, S, L):
    """both forward kernels on shapes whose feature / inducing sections do not fill the 256-unit operand tiles (and one that does),
    first and second order, odd and even input dimensions 11..16, several samples per launch, N not a multiple of the CTA: field, prior
    part and an RK4 rollout against
    the fp64 oracle on a random subset of a chip-filling batch"""
    N = 33001
    rs = np.random.RandomState(M + S)
    f64 = lambda a: torch.tensor(a, dtype=torch.float64)
    Z, ell, var = f64(rs.normal(size=(M, D_in))), f64(1.5 + rs.uniform(size=(D_out, D_in))), f64(0.5 + rs.uniform(size=D_out))
    caches = []
    for l in range(L):
        c = dict(variant="rbf_dimwise", Z=Z, ell=ell, var=var, nu=f64(rs.normal(size=(D_out, M, 1))), eps=f64(rs.normal(size=(D_in, S, D_out))),
                 phase=f64(rs.uniform(size=(1, S, D_out)) * 2 * np.pi), w=f64(rs.normal(size=(S, D_out))))
        c["omega"] = OF.make_omega(c["eps"], c["ell"], "rbf_dimwise")
        caches.append(c)
    ss = [gpu_sample(c) for c in caches]
    s = {k: (ss[0][k] if k in ("Z", "ell", "var", "B") else torch.cat([q[k] for q in ss], 0)) for k in ss[0]}
    x = torch.tensor(1.2 * rs.normal(size=(L, N, D_in)), dtype=torch.float32, device="cuda")
    f, fp = _field(s, x, "rbf_dimwise")
    idx = rs.choice(N, size=200, replace=False)
    ts = 0.1 * torch.arange(4, dtype=torch.float, device="cuda")
    traj = _gp().gp_rollout(x, ts, s["Z"], s["nu"], s["eps"], s["phase"], s["w"], s["ell"], s["var"], "rbf_dimwise", order, "rk4")
    for l, c in enumerate(caches):
        xs = x[l, idx].double().cpu()
        assert rel(f[l, idx], OF.field(xs, c)) < FIELD_TOL, (fwd_kernel, l)
        assert rel(fp[l, idx], OF.prior(xs, c)) < FIELD_TOL, (fwd_kernel, l)
        assert rel(traj[l, idx], OF.rollout(xs, ts.double().cpu(), c, order, "rk4")) < TRAJ_TOL, (fwd_kernel, l)


def test_tensor_path_forward_large_batch(fwd_kernel):
    """D > 8 and a chip-filling batch select the tensor-path forward (RbfMmaFwdPolicy / RbfTcFwdPolicy): field and RK4
    trajectories of 40,000 states against the fp64 oracle on a random subset, plus equality of overlapping small/large launches
    within the field bar (the small launch runs the FFMA kernel)."""
    g = load_golden("rbf_dimwise_d16")
    c = oracle_cache(g)
    s = gpu_sample(c)
    N, D = 40000, 16
    rs = np.random.RandomState(5)
    x = torch.tensor(1.5 * rs.normal(size=(1, N, D)), dtype=torch.float32, device="cuda")
    f, fp = _field(s, x, "rbf_dimwise")
    idx = rs.randint(0, N, size=300)
    truth = OF.field(x[0, idx].double().cpu(), c)
    assert rel(f[0, idx], truth) < FIELD_TOL
    assert rel(fp[0, idx], OF.prior(x[0, idx].double().cpu(), c)) < FIELD_TOL
    f_small, _ = _field(s, x[:, idx].contiguous(), "rbf_dimwise")
    assert rel(f_small[0], f[0, idx]) < FIELD_TOL
    ts = 0.1 * torch.arange(6, dtype=torch.float, device="cuda")
    traj = _gp().gp_rollout(x[0], ts, s["Z"], s["nu"], s["eps"], s["phase"], s["w"], s["ell"], s["var"], "rbf_dimwise", 1, "rk4")
    truth = OF.rollout(x[0, idx].double().cpu(), ts.double().cpu(), c, 1, "rk4")
    assert rel(traj[0, idx], truth) < TRAJ_TOL


@pytest.mark.parametrize("D_in,D_out,M,S,N", [(12, 5, 101, 33, 40000), (16, 4, 200, 31, 33000)])
def test_tensor_path_backward_large_batch(bwd_kernel, D_in, D_out, M, S, N):
    """chip-filling batches at D > 8 run the tensor-path forward AND reverse sweeps (RbfMmaFwdPolicy / RbfMmaBwdPolicy):
    field VJP, parameter gradients and a short RK4 rollout backward against autograd through the fp64 oracle."""
    rs = np.random.RandomState(D_in * 7 + M)
    f64 = lambda a: torch.tensor(a, dtype=torch.float64)
    c = dict(variant="rbf_dimwise", Z=f64(rs.normal(size=(M, D_in))), ell=f64(1.5 + rs.uniform(size=(D_out, D_in))),
             var=f64(0.5 + rs.uniform(size=D_out)), nu=f64(rs.normal(size=(D_out, M, 1))), eps=f64(rs.normal(size=(D_in, S, D_out))),
             phase=f64(rs.uniform(size=(1, S, D_out)) * 2 * np.pi), w=f64(rs.normal(size=(S, D_out))))
    for k in ("Z", "ell", "var", "nu"):
        c[k].requires_grad_(True)
    c["omega"] = OF.make_omega(c["eps"], c["ell"], "rbf_dimwise")
    x64 = f64(1.5 * rs.normal(size=(N, D_in))).requires_grad_(True)
    gout = f64(rs.normal(size=(N, D_out)))
    want = torch.autograd.grad((OF.field(x64, c) * gout).sum(), [x64, c["Z"], c["nu"], c["ell"], c["var"]])
    s = gpu_sample(c)
    for k in ("Z", "nu", "ell", "var"):
        s[k].requires_grad_(True)
    x = x64.detach().float().cuda()[None].requires_grad_(True)
    f, _ = _field(s, x, "rbf_dimwise")
    (f[0] * gout.float().cuda()).sum().backward()
    got = [x.grad[0], s["Z"].grad, s["nu"].grad[0], s["ell"].grad, s["var"].grad]
    for nm, a, b in zip(("dx", "dZ", "dnu", "dell", "dvar"), got, want):
        e = rel(a, b)
        print("tensor-path [%s] field-bwd D=%d %s: %.2e" % (bwd_kernel, D_in, nm, e))
        assert e < GRAD_TOL, (nm, e)


def test_tensor_path_rollout_backward_large_batch(bwd_kernel):
    D, M, S, N, T = 12, 64, 17, 33000, 3
    rs = np.random.RandomState(11)
    f64 = lambda a: torch.tensor(a, dtype=torch.float64)
    c = dict(variant="rbf_dimwise", Z=f64(rs.normal(size=(M, D))), ell=f64(1.5 + rs.uniform(size=(D, D))), var=f64(0.5 + rs.uniform(size=D)),
             nu=f64(0.3 * rs.normal(size=(D, M, 1))), eps=f64(rs.normal(size=(D, S, D))), phase=f64(rs.uniform(size=(1, S, D)) * 2 * np.pi),
             w=f64(rs.normal(size=(S, D))))
    for k in ("Z", "ell", "var", "nu"):
        c[k].requires_grad_(True)
    c["omega"] = OF.make_omega(c["eps"], c["ell"], "rbf_dimwise")
    z64 = f64(rs.normal(size=(N, D))).requires_grad_(True)
    ts64 = 0.1 * torch.arange(T, dtype=torch.float64)
    G = f64(rs.normal(size=(N, T, D)))
    want = torch.autograd.grad((OF.rollout(z64, ts64, c, 1, "rk4") * G).sum(), [z64, c["Z"], c["nu"], c["ell"], c["var"]])
    s = gpu_sample(c)
    for k in ("Z", "nu", "ell", "var"):
        s[k].requires_grad_(True)
    z0 = z64.detach().float().cuda().requires_grad_(True)
    traj = _gp().gp_rollout(z0, ts64.float().cuda(), s["Z"], s["nu"], s["eps"], s["phase"], s["w"], s["ell"], s["var"], "rbf_dimwise", 1, "rk4")
    (traj[0] * G.float().cuda()).sum().backward()
    got = [z0.grad, s["Z"].grad, s["nu"].grad[0], s["ell"].grad, s["var"].grad]
    for nm, a, b in zip(("dz0", "dZ", "dnu", "dell", "dvar"), got, want):
        e = rel(a, b)
        print("tensor-path [%s] rollout-bwd %s: %.2e" % (bwd_kernel, nm, e))
        assert e < GRAD_TOL, (nm, e)


@pytest.mark.parametrize("scale_x,scale_ell", [(300.0, 60.0), (0.002, 0.02), (30.0, 0.3), (1.0, 0.05)])
def test_tensor_path_extreme_magnitudes(fwd_kernel, bwd_kernel, scale_x, scale_ell):
    """the fp16 tensor-path dot products are scaled by exact powers of two per state and per output dimension: states and
    lengthscales far from O(1) (row coefficients from 1e-5 to 1e3, |x| up to ~1e3) must neither overflow nor lose the field bar
    where the field is well conditioned; checked against the fp64 oracle on a chip-filling batch (forward and VJP)."""
    D, M, S, N = 16, 64, 32, 36000
    rs = np.random.RandomState(3)
    f64 = lambda a: torch.tensor(a, dtype=torch.float64)
    c = dict(variant="rbf_dimwise", Z=f64(scale_x * rs.normal(size=(M, D))), ell=f64(scale_x * scale_ell * (1.0 + rs.uniform(size=(D, D)))),
             var=f64(0.5 + rs.uniform(size=D)), nu=f64(rs.normal(size=(D, M, 1))), eps=f64(rs.normal(size=(D, S, D))),
             phase=f64(rs.uniform(size=(1, S, D)) * 2 * np.pi), w=f64(rs.normal(size=(S, D))))
    c["omega"] = OF.make_omega(c["eps"], c["ell"], "rbf_dimwise")
    x64 = f64(scale_x * rs.normal(size=(N, D)))
    for k in ("Z", "nu", "ell", "var"):
        c[k].requires_grad_(True)
    c["omega"] = OF.make_omega(c["eps"], c["ell"], "rbf_dimwise")
    s = gpu_sample(c)
    for k in ("Z", "nu", "ell", "var"):
        s[k].requires_grad_(True)
    x = x64.float().cuda()[None].requires_grad_(True)
    f, _ = _field(s, x, "rbf_dimwise")
    assert torch.isfinite(f).all()
    xs = x64.clone().requires_grad_(True)
    truth = OF.field(xs, c)
    # the phase of the features is |x . omega| ~ 4 / scale_ell: fp32 itself resolves cos only to ~1e-7 x that
    tol = max(FIELD_TOL, 3e-6 / scale_ell)
    assert rel(f[0], truth) < tol, (rel(f[0], truth), tol)
    gsel = torch.tensor(rs.normal(size=(N, D)), dtype=torch.float32)
    f.backward(gsel.cuda()[None])
    want = torch.autograd.grad((truth * gsel.double()).sum(), [xs, c["Z"], c["nu"], c["ell"], c["var"]])
    got = [x.grad[0], s["Z"].grad, s["nu"].grad[0], s["ell"].grad, s["var"].grad]
    gtol = max(GRAD_TOL, 3e-5 / scale_ell)
    for nm, a, b in zip(("dx", "dZ", "dnu", "dell", "dvar"), got, want):
        assert torch.isfinite(a).all(), nm
        if b.norm() < 1e-30:      # K(x, Z) underflows fp32 at this lengthscale: the gradient through it is zero to fp32
            assert a.norm() < 1e-30, nm
            continue
        e = rel(a, b)
        print("extreme (%g, %g) %s: %.2e (tol %.1e)" % (scale_x, scale_ell, nm, e, gtol))
        assert e < gtol, (nm, e, gtol)


@pytest.mark.parametrize("shape", ["small", "tensor"])
def test_run_to_run_spread(shape):
    """The parameter gradients are accumulated with floating-point atomics (include/gpode.h): two runs of the same backward differ
    in their last bits.  This bounds the spread -- and pins what IS bit-reproducible (trajectories, dL/dz0: no atomics)."""
    rs = np.random.RandomState(5)
    if shape == "small":
        D, M, S, N, T = 6, 100, 256, 2000, 6
    else:
        D, M, S, N, T = 16, 512, 256, 33000, 3          # the fused tcgen05 reverse sweep (red.global of the PG tiles)
    f32 = lambda a: torch.tensor(a, dtype=torch.float32, device="cuda")
    Z, ell, var = f32(rs.normal(size=(M, D))), f32(1.5 + rs.uniform(size=(D, D))), f32(0.5 + rs.uniform(size=D))
    nu, eps = f32(0.3 * rs.normal(size=(1, D, M, 1))), f32(rs.normal(size=(1, D, S, D)))
    phase, w = f32(rs.uniform(size=(1, 1, S, D)) * 2 * np.pi), f32(rs.normal(size=(1, S, D)))
    z0v, G = f32(rs.normal(size=(N, D))), f32(rs.normal(size=(1, N, T, D)))
    ts = 0.1 * torch.arange(T, dtype=torch.float32, device="cuda")
    runs = []
    for _ in range(3):
        leaves = [t_.clone().requires_grad_(True) for t_ in (z0v, Z, nu, ell, var)]
        traj = _gp().gp_rollout(leaves[0], ts, leaves[1], leaves[2], eps, phase, w, leaves[3], leaves[4], "rbf_dimwise", 1, "rk4")
        (traj * G).sum().backward()
        runs.append([traj.detach()] + [t_.grad for t_ in leaves])
    for r in runs[1:]:
        assert torch.equal(r[0], runs[0][0]) and torch.equal(r[1], runs[0][1])           # trajectories and dz0: bit-identical
        for nm, a, b in zip(("dZ", "dnu", "dell", "dvar"), r[2:], runs[0][2:]):
            e = rel(a, b)
            print("%s run-to-run %s: %.2e" % (shape, nm, e))
            assert e < 2e-6, (nm, e)
