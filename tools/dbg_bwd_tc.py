"""debug driver: one field backward / short rollout backward on the tcgen05 reverse sweep, against the mma.sync kernels"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "vae-gp-ode_b200"))
import numpy as np, torch
import gpode_b200 as gp

def run(D_in, D_out, M, S, N, flags, T=0, seed=0):
    rs = np.random.RandomState(seed)
    f32 = lambda a: torch.tensor(a, dtype=torch.float32, device="cuda")
    Z, ell, var = f32(rs.normal(size=(M, D_in))), f32(1.5 + rs.uniform(size=(D_out, D_in))), f32(0.5 + rs.uniform(size=D_out))
    nu, eps = f32(0.3 * rs.normal(size=(1, D_out, M, 1))), f32(rs.normal(size=(1, D_in, S, D_out)))
    phase, w = f32(rs.uniform(size=(1, 1, S, D_out)) * 2 * np.pi), f32(rs.normal(size=(1, S, D_out)))
    x = f32(1.5 * rs.normal(size=(1, N, D_in))).requires_grad_(True)
    for t in (Z, ell, var, nu):
        t.requires_grad_(True)
    with gp.kernel_flags(flags):
        if T:
            ts = 0.1 * torch.arange(T, dtype=torch.float32, device="cuda")
            out = gp.gp_rollout(x[0], ts, Z, nu, eps, phase, w, ell, var, "rbf_dimwise", 1, "rk4")
        else:
            out, _ = gp.gp_field(x, Z, nu, eps, phase, w, ell, var, "rbf_dimwise")
        g = f32(np.random.RandomState(seed + 1).normal(size=tuple(out.shape)))
        (out * g).sum().backward()
    torch.cuda.synchronize()
    return [t.grad.detach().clone() for t in (x, Z, nu, ell, var)]

if __name__ == "__main__":
    shapes = [(12, 5, 101, 33, 40000, 0), (16, 16, 512, 256, 33000, 0), (16, 16, 512, 256, 33000, 3)]
    which = int(sys.argv[1]) if len(sys.argv) > 1 else -1
    for i, (D_in, D_out, M, S, N, T) in enumerate(shapes):
        if which >= 0 and i != which:
            continue
        a = run(D_in, D_out, M, S, N, gp.FLAG_BWD_MMA, T)
        b = run(D_in, D_out, M, S, N, gp.FLAG_BWD_TCGEN05, T)
        rel = lambda u, v: ((u - v).double().norm() / v.double().norm().clamp_min(1e-300)).item()
        print("shape", (D_in, D_out, M, S, N, T), "tc-vs-mma:", " ".join("%s %.2e" % (n, rel(u, v)) for n, u, v in zip(("dx", "dZ", "dnu", "dell", "dvar"), b, a)), flush=True)
