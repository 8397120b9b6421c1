"""CPU: the rearranged DF algebra the CUDA kernels implement (csrc/df.h, csrc/df_kernels.cuh) -- merged
cos/sin feature (R cos(theta - phi)), exponent form 2^(r2 k_ij + lc_ij), the W/V form of the VJP and the
dc_ij accumulator of the lengthscale gradient -- restated in fp64 torch and checked against autograd
through the oracle (oracle/field.py, which is pinned to the reference by tests/golden)."""
import math

import pytest
import torch

from oracle import field as OF

dt = torch.float64


@pytest.mark.parametrize("D,M,S,N", [(5, 7, 9, 11), (2, 3, 4, 5), (6, 10, 8, 4)])
def test_df_kernel_algebra_matches_autograd(D, M, S, N):
    torch.manual_seed(D * 100 + M)
    x = torch.randn(N, D, dtype=dt)
    Z = torch.randn(M, D, dtype=dt).requires_grad_(True)
    ell = (1.5 + torch.rand(D, D, dtype=dt)).requires_grad_(True)
    var = (0.5 + torch.rand(D, dtype=dt)).requires_grad_(True)
    nu = torch.randn(M * D, 1, dtype=dt).requires_grad_(True)
    eps = torch.randn(D, S, D, dtype=dt)
    phase = torch.rand(1, S, D, dtype=dt) * 2 * math.pi
    w = torch.randn(2 * S, D, dtype=dt)
    g = torch.randn(N, D, dtype=dt)
    xg = x.clone().requires_grad_(True)
    omega = OF.make_omega(eps, ell, "df")
    B = OF.df_B(omega).detach().requires_grad_(True)   # B is an independent input of the C ABI
    c = dict(variant="df", Z=Z, ell=ell, var=var, nu=nu, omega=omega, phase=phase, w=w, B=B)
    f = OF.field(xg, c)
    fp_o = OF.prior(xg, c)
    want = torch.autograd.grad((f * g).sum(), [xg, Z, nu, ell, var, B])

    with torch.no_grad():
        cc = 1 / ell ** 2
        kk = -0.5 * math.log2(math.e) * cc
        lc = torch.log2(var[None, :] * cc ** 2)
        h = (D - 1) / torch.diagonal(cc)
        nuM = nu.view(M, D)
        Om = eps / ell.t()[:, None, :]
        w1, w2 = w[:S], w[S:]
        R, phi = torch.hypot(w1, w2), torch.atan2(w2, w1)
        bp = phase[0] - phi
        amp = torch.sqrt(var / S)
        Bp = amp[None, None, :] * R[:, :, None] * B
        th = torch.einsum("nd,dsa->nsa", x, Om) + bp
        fp = torch.einsum("nsa,sac->nc", torch.cos(th), Bp)
        d = x[:, None, :] - Z[None, :, :]
        r2 = (d ** 2).sum(-1)
        e = torch.exp2(r2[:, :, None, None] * kk + lc)
        p = nuM[None] * d
        s = torch.einsum("nmi,nmij->nmj", p, e)
        ediag = torch.diagonal(e, dim1=2, dim2=3)
        hr = h - r2[:, :, None]
        fu = (d * s + nuM[None] * ediag * hr).sum(1)
        assert (fp - fp_o).abs().max() < 1e-12 and (fp + fu - f).abs().max() < 1e-12
        # VJP
        Gc = torch.einsum("nc,sac->nsa", g, Bp)
        t = torch.cos(th + math.pi / 2) * Gc
        q = torch.einsum("nsa,dsa->nad", t, Om)
        dell_theta = (x[:, None, :] * q).sum(0)
        dBp = torch.einsum("nc,nsa->sac", g, torch.cos(th))
        dB = amp[None, None, :] * R[:, :, None] * dBp
        u = g[:, None, :] * d
        tt = torch.einsum("nmj,nmkj->nmk", u, e)
        pe = p[:, :, :, None] * e
        wj = (kk * pe).sum(2)
        kjj = torch.diagonal(kk)
        gne = g[:, None, :] * nuM[None] * ediag
        WV = (u * wj).sum(-1) + (gne * (kjj * hr - math.log2(math.e))).sum(-1)
        wv = 2 * math.log(2) * WV
        dd = g[:, None, :] * s + nuM[None] * tt + d * wv[:, :, None]
        dx = q.sum(1) + dd.sum(1)
        dZ = -dd.sum(0)
        dnu = (d * tt + g[:, None, :] * ediag * hr).sum(0)
        a2 = r2[:, :, None, None] * kk + 2 * math.log2(math.e)
        dc = (u[:, :, None, :] * pe * a2).sum((0, 1))
        a2jj = torch.diagonal(a2, dim1=2, dim2=3)
        dc = dc + torch.diag((gne * (a2jj * hr - h * math.log2(math.e))).sum((0, 1)))
        dell = -(dell_theta + 2 * math.log(2) * dc) / ell
        dvar = (g * (f - 0.5 * fp_o)).sum(0) / var
    for nm, a, b in zip(["dx", "dZ", "dnu", "dell", "dvar", "dB"], [dx, dZ, dnu.reshape(-1, 1), dell, dvar, dB], want):
        err = ((a - b).norm() / b.norm()).item()
        assert err < 1e-11, (nm, err)
