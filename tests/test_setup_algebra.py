"""CPU: the closed-form backward of nu = Lc^-T (u - Lc^-1 p), Lc = chol(A) that csrc/setup_kernels.cu implements
(no generic Cholesky backward: A_bar = Lc^-T S Lc^-1 + 1/2 sum (r q^T + q r^T), S_ij = -1/2 sum u[max(i,j)] bb[min(i,j)])
checked in fp64 against autograd through torch.linalg.cholesky / solve_triangular (the reference's op sequence,
experiments/model/core/kernels.py:163-171)."""
import pytest
import torch

dt = torch.float64


@pytest.mark.parametrize("M,R", [(9, 3), (33, 1), (64, 5)])
def test_nu_backward_closed_form(M, R):
    torch.manual_seed(M)
    B = torch.randn(M, M, dtype=dt)
    A = (B @ B.T + M * torch.eye(M, dtype=dt)).requires_grad_(True)
    u = torch.randn(M, R, dtype=dt, requires_grad=True)
    p = torch.randn(M, R, dtype=dt, requires_grad=True)
    Lc = torch.linalg.cholesky(A)
    a = torch.linalg.solve_triangular(Lc, p, upper=False)
    nu = torch.linalg.solve_triangular(Lc.T, u - a, upper=True)
    nub = torch.randn(M, R, dtype=dt)
    gA, gu, gp = torch.autograd.grad((nu * nub).sum(), [A, u, p])
    with torch.no_grad():
        bb = torch.linalg.solve_triangular(Lc, nub, upper=False)
        r = torch.linalg.solve_triangular(Lc.T, bb, upper=True)
        q = torch.linalg.solve_triangular(Lc.T, a, upper=True)
        i, j = torch.meshgrid(torch.arange(M), torch.arange(M), indexing="ij")
        S = -0.5 * (u[torch.maximum(i, j)] * bb[torch.minimum(i, j)]).sum(-1)
        Y = torch.linalg.solve_triangular(Lc.T, S, upper=True)
        X = torch.linalg.solve_triangular(Lc.T, Y.T, upper=True)      # = X^T = X
        Abar = X + 0.5 * (r @ q.T + q @ r.T)
    assert (bb - gu).abs().max() < 1e-12
    assert (-r - gp).abs().max() < 1e-12
    assert (Abar - 0.5 * (gA + gA.T)).abs().max() < 1e-12 * max(1.0, gA.abs().max().item())
    assert (X - X.T).abs().max() < 1e-12
