"""Where does the end-to-end step (bench.py's e2e_step) of a small workload spend its time?  torch.profiler table of the CUDA
kernels and the wall-clock per phase.  Usage (GPU box): python tools/profile_e2e.py [workload] [steps]"""
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "vae-gp-ode_b200")):
    sys.path.insert(0, p)
import numpy as np  # noqa: E402
import torch  # noqa: E402

import bench  # noqa: E402


def main():
    name = sys.argv[1] if len(sys.argv) > 1 else "cfg2_df_d6_m100_t16_rk4"
    steps = int(sys.argv[2]) if len(sys.argv) > 2 else 10
    w = dict(bench.WORKLOADS[name])
    dev = torch.device("cuda", 0)
    gp, flow = bench.build_model(w, dev, seed=1)
    N, L, T, D = w["N"], w["L"], w["T"], w["D_in"]
    z0_host = torch.randn(N, D).pin_memory()
    dtraj = torch.randn(L, N, T, D, device=dev)
    ts = 0.1 * torch.arange(T, dtype=torch.float, device=dev)
    params = [gp.kern.unconstrained_lengthscales, gp.kern.unconstrained_variance, gp.inducing_loc.optvar, gp.Um.optvar, gp.Us_sqrt.optvar]
    bench.seeded_draw_patch(77)
    phases = {}

    def tick(k, t0):
        torch.cuda.synchronize()
        phases[k] = phases.get(k, 0.0) + time.perf_counter() - t0
        return time.perf_counter()

    def step(sync):
        for p_ in params:
            p_.grad = None
        t0 = time.perf_counter()
        z = z0_host.to(dev, non_blocking=True).requires_grad_(True)
        fs = gp.build_cache_batched(L)
        if sync:
            t0 = tick("setup (draws + H2D + inducing sample + prior at Z + compute_nu)", t0)
        traj = flow._rollout(z, ts, fs)
        if sync:
            t0 = tick("rollout forward", t0)
        loss = (traj * dtraj).sum() + flow.kl()
        if sync:
            t0 = tick("loss", t0)
        loss.backward()
        if sync:
            t0 = tick("backward (rollout + setup)", t0)
        host = [loss.detach().cpu()] + [p_.grad.cpu() for p_ in params] + [z.grad.cpu()]
        if sync:
            tick("D2H", t0)
        return host

    for _ in range(3):
        step(False)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(steps):
        step(False)
    torch.cuda.synchronize()
    print("%s: %.3f ms per e2e step (no phase syncs)" % (name, (time.perf_counter() - t0) / steps * 1e3))
    for _ in range(steps):
        step(True)
    for k, v in phases.items():
        print("  %-70s %.3f ms" % (k, v / steps * 1e3))
    from torch.profiler import ProfilerActivity, profile
    with profile(activities=[ProfilerActivity.CPU, ProfilerActivity.CUDA]) as prof:
        for _ in range(steps):
            step(False)
        torch.cuda.synchronize()
    print(prof.key_averages().table(sort_by="cuda_time_total", row_limit=25, max_name_column_width=60))
    print(prof.key_averages().table(sort_by="self_cpu_time_total", row_limit=15, max_name_column_width=60))


if __name__ == "__main__":
    main()
