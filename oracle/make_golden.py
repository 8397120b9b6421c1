"""Generate tests/golden/*.npz from the LIVE reference modules.  TEST INFRASTRUCTURE ONLY.

Run in the build container (needs /root/reference):   python -m oracle.make_golden
Every vector is produced by the unmodified reference classes (SVGP_Layer, RBF, DivergenceFreeKernel,
ODEfunc, Flow -- experiments/model/core/{svpy,kernels,flow}.py) on CPU in fp32, with seeded draws
(oracle/reference_harness.py) and with ``torchdiffeq`` replaced by oracle/solvers.py (absent
third-party dependency; parity unpinned at that boundary, see oracle/__init__.py).

Per case the file holds: the leaf parameters (state_dict layout), the host draws in draw order, a
batch of states ``x`` with the reference field f(x) and its VJP, an initial state ``z0`` with the
reference trajectories for euler and rk4 and the reference gradients of sum(traj * G) + kl w.r.t.
z0 and every GP leaf parameter.
"""
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
from oracle import reference_harness as rh  # noqa: E402

OUT = os.path.join(os.path.dirname(HERE), "tests", "golden")

CASES = {
    # name: kernel, dimwise, D_in, D_out, order, M, S, N, T, ell, var, perturb
    "rbf_dimwise_o1": dict(kernel="RBF", dimwise=True, D_in=6, D_out=6, order=1, M=100, S=256, N=25, T=16, ell=2.0, var=1.0, perturb=0.0),
    "rbf_dimwise_o1_pert": dict(kernel="RBF", dimwise=True, D_in=6, D_out=6, order=1, M=64, S=128, N=33, T=9, ell=1.0, var=0.7, perturb=1.0),
    "rbf_shared_o1": dict(kernel="RBF", dimwise=False, D_in=6, D_out=6, order=1, M=64, S=128, N=25, T=16, ell=1.5, var=0.7, perturb=0.0),
    "rbf_dimwise_o2": dict(kernel="RBF", dimwise=True, D_in=6, D_out=3, order=2, M=100, S=256, N=25, T=16, ell=2.0, var=1.0, perturb=0.0),
    "rbf_dimwise_d16": dict(kernel="RBF", dimwise=True, D_in=16, D_out=16, order=1, M=96, S=64, N=17, T=8, ell=3.0, var=1.0, perturb=0.5),
    "rbf_dimwise_d3": dict(kernel="RBF", dimwise=True, D_in=3, D_out=3, order=1, M=40, S=50, N=19, T=6, ell=1.0, var=0.5, perturb=0.3),
    "df_o1": dict(kernel="DF", dimwise=True, D_in=6, D_out=6, order=1, M=100, S=256, N=32, T=16, ell=2.0, var=1.0, perturb=0.0),
    "df_o1_pert": dict(kernel="DF", dimwise=True, D_in=6, D_out=6, order=1, M=48, S=96, N=21, T=9, ell=1.5, var=0.8, perturb=0.02),
    "df_d4": dict(kernel="DF", dimwise=True, D_in=4, D_out=4, order=1, M=40, S=64, N=20, T=6, ell=1.2, var=0.6, perturb=0.01),
}


def variant_of(c):
    return "df" if c["kernel"] == "DF" else ("rbf_dimwise" if c["dimwise"] else "rbf_shared")


def build_layer(c, seed):
    """reference SVGP_Layer with seeded numpy init (svpy.py:76-86) and the case's ell / var."""
    ref = rh.load_reference()
    from model.misc.constraint_utils import invsoftplus
    np.random.seed(seed)
    gp = ref["svpy"].SVGP_Layer(D_in=c["D_in"], D_out=c["D_out"], M=c["M"], S=c["S"], q_diag=False,
                                dimwise=c["dimwise"], device="cpu", kernel=c["kernel"])
    rs = np.random.RandomState(seed + 1000)
    k = gp.kern
    lval = c["ell"] + c["perturb"] * rs.uniform(size=tuple(k.unconstrained_lengthscales.shape))
    vval = c["var"] + c["perturb"] * rs.uniform(size=tuple(k.unconstrained_variance.shape))
    k.unconstrained_lengthscales.data = invsoftplus(torch.tensor(lval.astype(np.float32)))
    k.unconstrained_variance.data = invsoftplus(torch.tensor(vval.astype(np.float32)))
    with torch.no_grad():
        gp.Us_sqrt.optvar.add_(0.02 * torch.tensor(rs.normal(size=tuple(gp.Us_sqrt.optvar.shape)).astype(np.float32)))
    return gp


def leaf_params(gp):
    return dict(raw_ell=gp.kern.unconstrained_lengthscales, raw_var=gp.kern.unconstrained_variance,
                Z=gp.inducing_loc.optvar, Um=gp.Um.optvar, Us_sqrt=gp.Us_sqrt.optvar)


def make_case(name, c, seed):
    ref = rh.load_reference()
    out = {}
    gp = build_layer(c, seed)
    leaves = leaf_params(gp)
    for k, v in leaves.items():
        out["p_" + k] = v.detach().numpy().copy()
    rs = np.random.RandomState(seed + 2000)
    D_s = c["D_in"]
    x = torch.tensor((1.5 * rs.normal(size=(64, c["D_in"]))).astype(np.float32), requires_grad=True)
    g = torch.tensor(rs.normal(size=(64, c["D_out"])).astype(np.float32))
    z0 = torch.tensor(rs.normal(size=(c["N"], D_s)).astype(np.float32), requires_grad=True)
    ts = 0.1 * torch.arange(c["T"], dtype=torch.float)          # odegpvae.py:39
    G = torch.tensor(rs.normal(size=(c["N"], c["T"], D_s)).astype(np.float32))
    out.update(x=x.detach().numpy(), g=g.numpy(), z0=z0.detach().numpy(), ts=ts.numpy(), G=G.numpy())

    # ---- field level: one build_cache, f(x), VJP (grads flow through nu/omega to the leaves) ----
    rec = rh.seed_draws(seed + 3000)
    gp.build_cache()
    names = ["w", "eps", "phase01", "eps_u"]
    for nme, (_, v) in zip(names, rec.log):
        out["draw_" + nme] = v
    f = gp(x)
    out["field_f"] = f.detach().numpy()
    out["field_nu"] = gp.kern.nu.detach().numpy()
    out["field_fprior"] = gp.kern.rff_forward(x, gp.S).detach().numpy()
    loss = (f * g).sum()
    grads = torch.autograd.grad(loss, [x] + list(leaves.values()), allow_unused=True)
    out["field_dx"] = grads[0].numpy()
    for k, gv in zip(leaves.keys(), grads[1:]):
        out["field_d" + k] = np.zeros_like(out["p_" + k]) if gv is None else gv.numpy()

    # ---- rollout level: Flow.forward with the same draws (flow.py:68-86), euler and rk4 ----
    for method in ("euler", "rk4"):
        flow = ref["flow"].Flow(diffeq=gp, order=c["order"], solver=method, use_adjoint=False)
        rh.seed_draws(seed + 3000)
        traj = flow(z0, ts)                                     # (N,T,D_s)
        out["traj_" + method] = traj.detach().numpy()
        out["nevals_" + method] = np.float32(flow.num_evals())
        kl = flow.kl()
        out["kl"] = np.float32(kl.item())
        loss = (traj * G).sum() + kl
        grads = torch.autograd.grad(loss, [z0] + list(leaves.values()), allow_unused=True)
        out["roll_%s_dz0" % method] = grads[0].numpy()
        for k, gv in zip(leaves.keys(), grads[1:]):
            out["roll_%s_d%s" % (method, k)] = np.zeros_like(out["p_" + k]) if gv is None else gv.numpy()
    meta = dict(c)
    meta["variant"] = variant_of(c)
    out["meta"] = np.array(repr(meta))
    os.makedirs(OUT, exist_ok=True)
    np.savez_compressed(os.path.join(OUT, name + ".npz"), **out)
    return out


def main():
    torch.manual_seed(0)
    torch.set_num_threads(8)
    for i, (name, c) in enumerate(CASES.items()):
        o = make_case(name, c, seed=100 + 17 * i)
        print(name, "f", o["field_f"].shape, "traj", o["traj_rk4"].shape, "|traj|max", float(np.abs(o["traj_rk4"]).max()),
              "nevals", float(o["nevals_rk4"]), "kl", float(o["kl"]))


if __name__ == "__main__":
    main()
