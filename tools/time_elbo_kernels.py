"""Timing of the kernels either side of the flow (csrc/elbo_kernels.cu) against their HBM roofline, next to the stock PyTorch ops they
replace.  Usage (GPU box): python tools/time_elbo_kernels.py  -> one JSON line"""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "vae-gp-ode_b200")):
    sys.path.insert(0, p)
import numpy as np  # noqa: E402
import torch  # noqa: E402

import gpode_b200 as gp  # noqa: E402


def timed(fn, reps=20, warm=3):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(reps):
        fn()
    b.record()
    torch.cuda.synchronize()
    return a.elapsed_time(b) / reps


def main():
    peak = 6549.8
    try:
        peak = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"]
    except Exception:
        pass
    out = {"hbm_peak_gbs": peak}
    L, N, T, pix = 4, 2048, 16, 784                       # config-2 shapes x 8 trajectories: z is 411 MB (> the 126 MB L2)
    x = torch.randn(N, T, 1, 28, 28, device="cuda")
    z = torch.rand(L, N, T, 1, 28, 28, device="cuda") * 0.9 + 0.05
    g = torch.randn(N, device="cuda")
    elems = z.numel()
    zr = z.clone().requires_grad_(True)
    ms_f = timed(lambda: gp.bernoulli_lhood(x, zr))
    lh = gp.bernoulli_lhood(x, zr)
    ms_b = timed(lambda: torch.autograd.grad(lh, zr, g, retain_graph=True))

    def ref_fwd():
        XL = x.repeat([L, 1, 1, 1, 1, 1])
        return (torch.log(zr) * XL + torch.log(1 - zr) * (1 - XL)).sum([2, 3, 4, 5]).mean(0)
    ms_rf = timed(ref_fwd, reps=5)
    lr = ref_fwd()
    ms_rb = timed(lambda: torch.autograd.grad(lr, zr, g, retain_graph=True), reps=5)
    out["bernoulli"] = {"elements": elems, "fwd_ms": round(ms_f, 4), "bwd_ms": round(ms_b, 4),
                        "fwd_gbs": round(elems * 4 * (1 + 1.0 / L) / ms_f / 1e6, 1), "bwd_gbs": round(elems * 4 * (2 + 1.0 / L) / ms_b / 1e6, 1),
                        "fwd_frac_hbm": round(elems * 4 * (1 + 1.0 / L) / ms_f / 1e6 / peak, 3), "bwd_frac_hbm": round(elems * 4 * (2 + 1.0 / L) / ms_b / 1e6 / peak, 3),
                        "torch_fwd_ms": round(ms_rf, 3), "torch_bwd_ms": round(ms_rb, 3),
                        "bytes_model": "fwd: z once + x once per sample group (4 + 4/L B per element); bwd: z + dz + x (8 + 4/L B)"}
    # draws: config-5 function samples (L = 8: w, eps, phase, eps_u) device vs host numpy + H2D
    st = gp.PhiloxStream(1)
    L5, S, D, M = 8, 256, 16, 512
    outs = [torch.empty(L5, S, D, device="cuda"), torch.empty(L5, D, S, D, device="cuda"), torch.empty(L5, 1, S, D, device="cuda"), torch.empty(L5, M, D, device="cuda")]
    ms_d = timed(lambda: st.fill(outs, [0, 0, 1, 0]))
    rs = np.random.RandomState(0)
    import time
    t0 = time.perf_counter()
    for _ in range(5):
        hs = [torch.tensor(rs.normal(size=tuple(o.shape)).astype(np.float32)).cuda() for o in outs]
    torch.cuda.synchronize()
    ms_h = (time.perf_counter() - t0) / 5 * 1e3
    n = sum(o.numel() for o in outs)
    out["draws"] = {"numbers": n, "device_ms": round(ms_d, 4), "host_numpy_plus_h2d_ms": round(ms_h, 3), "device_gnumbers_per_s": round(n / ms_d / 1e6, 2)}
    big = torch.empty(1 << 28, device="cuda")
    ms_big = timed(lambda: st.fill([big], [0]), reps=5)
    out["draws"]["fill_1GiB_normal_ms"] = round(ms_big, 3)
    out["draws"]["fill_1GiB_write_gbs"] = round(big.numel() * 4 / ms_big / 1e6, 1)
    print(json.dumps(out))


if __name__ == "__main__":
    main()
