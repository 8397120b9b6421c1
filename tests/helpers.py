"""Shared helpers for the parity tests: golden fixtures (tests/golden, produced by oracle/make_golden.py
from the live reference) and oracle caches built from them."""
import ast
import os

import numpy as np
import torch

from oracle import field as OF

GOLDEN_DIR = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
RBF_CASES = ["rbf_dimwise_o1", "rbf_dimwise_o1_pert", "rbf_shared_o1", "rbf_dimwise_o2", "rbf_dimwise_d16", "rbf_dimwise_d3"]
DF_CASES = ["df_o1", "df_o1_pert", "df_d4"]
ALL_CASES = RBF_CASES + DF_CASES


def load_golden(name):
    z = np.load(os.path.join(GOLDEN_DIR, name + ".npz"))
    g = {k: z[k] for k in z.files}
    g["meta"] = ast.literal_eval(str(g["meta"]))
    return g


def t(a, dtype=torch.float32, device="cpu"):
    return torch.tensor(np.asarray(a), dtype=dtype, device=device)


def rel(a, b):
    """norm-wise relative error |a-b|_F / |b|_F in float64 (SURVEY.md §8d parity metric)."""
    a = torch.as_tensor(a).detach().double().cpu()
    b = torch.as_tensor(b).detach().double().cpu()
    return ((a - b).norm() / b.norm().clamp_min(1e-300)).item()


def oracle_cache(g, dtype=torch.float64, shared_nu=True, leaves=False):
    """Oracle cache from a golden case; shared_nu=True pins nu to the reference's own fp32 nu (the 1e-5
    field parity is only meaningful with a shared nu)."""
    m = g["meta"]
    M = m["M"]
    draws = dict(w=t(g["draw_w"], dtype), eps=t(g["draw_eps"], dtype), phase01=t(g["draw_phase01"], dtype),
                 eps_u=t(g["draw_eps_u"], dtype))
    Lq = OF.tril_from_packed(t(g["p_Us_sqrt"], dtype), M)
    nu = t(g["field_nu"], dtype) if shared_nu else None
    c = OF.build_cache(m["variant"], t(g["p_Z"], dtype), t(g["p_raw_ell"], dtype), t(g["p_raw_var"], dtype), t(g["p_Um"], dtype),
                       Lq, draws, nu_override=nu)
    c["eps"] = draws["eps"]
    if leaves:
        # make Z, ell, var, nu independent leaves (kernel-level gradient oracle); omega/B stay functions of ell
        for k in ("Z", "ell", "var", "nu"):
            c[k] = c[k].detach().clone().requires_grad_(True)
        c["omega"] = OF.make_omega(c["eps"], c["ell"], m["variant"])
        if m["variant"] == "df":
            c["B"] = OF.df_B(c["omega"])
    return c


def gpu_sample(c, device="cuda"):
    """Tensors for the CUDA entry points (leading sample axis L=1) from an oracle cache."""
    f32 = lambda v: v.detach().to(torch.float32).to(device).contiguous()
    out = dict(Z=f32(c["Z"]), ell=f32(c["ell"]), var=f32(c["var"]), eps=f32(c["eps"])[None], phase=f32(c["phase"])[None],
               w=f32(c["w"])[None], nu=f32(c["nu"])[None], B=None)
    if c["variant"] == "df":
        out["B"] = f32(c["B"])[None]
    return out
