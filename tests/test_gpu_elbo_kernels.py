"""GPU parity of the kernels either side of the flow (csrc/elbo_kernels.cu) through the C ABI:
device-side Philox draws (bit-exact generator, draw mapping, the "device" draw mode of SVGP_Layer as a distribution) and the fused,
reduced Bernoulli log-likelihood (reference core/vae.py:136-153 + create_model.py:51-53) forward and backward."""
import ctypes

import numpy as np
import pytest
import torch
from scipy import stats

from oracle import philox as P
from helpers import rel

pytestmark = pytest.mark.gpu


def _gp():
    import gpode_b200
    return gpode_b200


def test_philox_generator_is_bit_exact():
    from gpode_b200 import _lib
    lib = _lib.load()
    rs = np.random.RandomState(0)
    n = 4096
    ctr = rs.randint(0, 2 ** 32, size=(n, 4), dtype=np.uint64).astype(np.uint32)
    key = rs.randint(0, 2 ** 32, size=(n, 2), dtype=np.uint64).astype(np.uint32)
    for i, (c, k, _) in enumerate(P.KAT):
        ctr[i], key[i] = c, k
    dc, dk = torch.tensor(ctr.astype(np.int64), device="cuda").to(torch.int32), torch.tensor(key.astype(np.int64), device="cuda").to(torch.int32)
    out = torch.empty((n, 4), dtype=torch.int32, device="cuda")
    rc = lib.gpode_philox_raw(_lib.ptr(dc), _lib.ptr(dk), _lib.ptr(out), n, _lib.stream_handle(out.device))
    assert rc == 0
    got = out.cpu().numpy().view(np.uint32)
    assert np.array_equal(got, P.philox4x32_10(ctr, key))
    for i, (_, _, want) in enumerate(P.KAT):
        assert tuple(int(v) for v in got[i]) == want


@pytest.mark.parametrize("n", [1, 7, 4096, 100003])
def test_philox_fill_matches_oracle(n):
    st = _gp().PhiloxStream(seed=0x1234567887654321)
    st.offset = 2 ** 33 + 5          # a 64-bit counter
    outs = [torch.empty(n, device="cuda"), torch.empty(max(1, n // 3), device="cuda"), torch.empty(n + 2, device="cuda")[1:-1]]   # the last one unaligned
    kinds = [0, 1, 0]
    off = st.offset
    st.fill(outs, kinds)
    assert st.offset == off + (n + 3) // 4
    for s, (o, k) in enumerate(zip(outs, kinds)):
        want = P.fill(o.numel(), k, st.seed, off, s)
        got = o.cpu().numpy()
        if k == 1:
            assert np.array_equal(got, want)                # 24-bit uniforms: exact
        else:
            assert np.abs(got - want).max() < 2e-6          # logf / sincospif vs numpy: a few ulp of a |value| <= 6
    with pytest.raises(RuntimeError):
        st.fill([torch.empty(4)], [0])                      # CPU tensor: no host path


def test_device_draw_mode_of_the_layer():
    """set_rng("device"): build_cache_batched takes w, eps, phase, eps_u of all L samples from ONE Philox launch.  Same distribution
    as the host helpers (KS tests; prior-field moments at fixed states against the host mode), reproducible from the seed."""
    gp = _gp()
    from gpode_b200.core import svpy as SV
    from gpode_b200.core.svpy import SVGP_Layer
    np.random.seed(0)
    L, N = 8, 64
    x = torch.tensor(np.random.normal(size=(L, N, 6)), dtype=torch.float32, device="cuda")
    try:
        for kernel in ("RBF", "DF"):
            layer = SVGP_Layer(6, 6, 100, 256, dimwise=True, device="cuda", kernel=kernel)
            SV.set_rng("device", seed=11)
            a = layer.build_cache_batched(L)
            SV.set_rng("device", seed=11)
            b = layer.build_cache_batched(L)
            assert torch.equal(a.eps, b.eps) and torch.equal(a.w, b.w) and torch.equal(a.phase, b.phase) and torch.equal(a.nu, b.nu)
            c = layer.build_cache_batched(L)                                  # the stream moves on
            assert not torch.equal(a.eps, c.eps)
            for t_, dist in ((a.eps, "norm"), (a.w, "norm")):
                assert stats.kstest(t_.flatten().cpu().numpy(), dist).pvalue > 1e-3
            ph = (a.phase / (2 * np.pi)).flatten().cpu().numpy()
            assert ph.min() >= 0 and ph.max() < 1 and stats.kstest(ph, "uniform").pvalue > 1e-3
            # statistical parity of the function prior: mean ~ 0 and variance ~ var at every state, device vs host draws (L x rep samples)
            def moments(mode):
                SV.set_rng(mode, seed=5)
                fs = []
                for _ in range(24):
                    s = layer.build_cache_batched(L)
                    nu0 = torch.zeros_like(s.nu)
                    _, fp = gp.gp_field(x, s.Z, nu0, s.eps, s.phase, s.w, s.ell, s.var, s.variant, s.B)
                    fs.append(fp)
                f = torch.cat(fs, 0)                                         # (24 L, N, D)
                return f.mean().item(), f.var(0).mean().item(), f.shape[0] * N * 6
            with torch.no_grad():
                m_d, v_d, cnt = moments("device")
                m_h, v_h, _ = moments("host")
            print("%s prior field: device mean %.4f var %.4f | host mean %.4f var %.4f" % (kernel, m_d, v_d, m_h, v_h))
            assert abs(m_d) < 0.05 and abs(m_h) < 0.05
            assert abs(v_d - v_h) < 0.15 * v_h
    finally:
        SV.set_rng("host")


@pytest.mark.parametrize("L,N,T,pix", [(1, 25, 16, 784), (4, 256, 16, 784), (3, 7, 5, 49), (2, 1, 1, 3)])
def test_bernoulli_lhood_forward_backward(L, N, T, pix):
    rs = np.random.RandomState(L + N)
    x = torch.tensor(rs.normal(size=(N, T, 1, pix)), dtype=torch.float64)                   # normalised pixels, not in [0, 1]
    z = torch.tensor(rs.uniform(0.02, 0.98, size=(L, N, T, 1, pix)), dtype=torch.float64).requires_grad_(True)
    XL = x.repeat([L, 1, 1, 1, 1])                                                           # vae.py:142-146 as written
    log_p = torch.log(z) * XL.view_as(z) + torch.log(1 - z) * (1 - XL.view_as(z))
    want = log_p.sum([2, 3, 4]).mean(0)
    G = torch.tensor(rs.normal(size=N), dtype=torch.float64)
    gz, = torch.autograd.grad((want * G).sum(), [z])
    zc = z.detach().float().cuda().requires_grad_(True)
    got = _gp().bernoulli_lhood(x.float().cuda(), zc)
    assert got.shape == (N,)
    e = rel(got, want)
    (got * G.float().cuda()).sum().backward()
    e_g = rel(zc.grad, gz)
    print("bernoulli L=%d N=%d P=%d: lhood %.2e  dz %.2e" % (L, N, T * pix, e, e_g))
    assert e < 1e-6 and e_g < 1e-6


def test_fused_elbo_matches_reference_formula():
    """gpode_b200.core.odegpvae.elbo against create_model.py:37-58 restated with a stub model (prior / q_dist as in vae.py)"""
    from gpode_b200.core import odegpvae as GO
    rs = np.random.RandomState(3)
    L, N, T, q = 3, 10, 4, 6

    class Enc:
        def q_dist(self, mu_s, logvar_s, mu_v=None, logvar_v=None):
            return torch.distributions.Normal(mu_s, torch.exp(0.5 * logvar_s))

    class Dec:
        distribution = "bernoulli"

    class Vae:
        encoder, decoder = Enc(), Dec()
        prior = torch.distributions.Normal(torch.zeros(q, device="cuda"), torch.ones(q, device="cuda"))

    class Flow:
        def kl(self):
            return torch.tensor(1.25, device="cuda")

    class Model:
        vae, flow = Vae(), Flow()
    f32 = lambda a: torch.tensor(a, dtype=torch.float32, device="cuda")
    X, Xrec = f32(rs.normal(size=(N, T, 1, 28, 28))), f32(rs.uniform(0.05, 0.95, size=(L, N, T, 1, 28, 28)))
    mu, logv = f32(rs.normal(size=(N, q))), f32(0.1 * rs.normal(size=(N, q)))
    lh, klr, klu = GO.elbo(Model(), X, Xrec, mu, logv, None, None, L)
    XL = X.double().repeat([L, 1, 1, 1, 1, 1])
    want = (torch.log(Xrec.double()) * XL + torch.log(1 - Xrec.double()) * (1 - XL)).sum([2, 3, 4, 5]).mean(0).mean()
    assert abs(lh.item() - want.item()) < 1e-6 * abs(want.item())
    assert klu.item() == 1.25 and klr.shape == ()
