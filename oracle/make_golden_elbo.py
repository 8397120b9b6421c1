"""Generate tests/golden/elbo_*.npz from the LIVE reference's full model.  TEST INFRASTRUCTURE ONLY.

Run in the build container (needs /root/reference):   python -m oracle.make_golden_elbo
For each BASELINE config 1-3 (oracle/elbo_harness.CONFIGS) and solver the unmodified reference -- build_model, ODEGPVAE, VAE,
SVGP_Layer, Flow, compute_loss (experiments/model/create_model.py:9-35,61-73) with torchdiffeq replaced by oracle/solvers.py --
runs one training-step loss + backward on CPU on the synthetic rotating-digit batch (oracle/glyph.py, seed 121), twice:
  ref32: exactly as the reference runs (fp32 parameters, fp32 draws) -> the like-for-like target;
  ref64: the same code, parameters and draws in float64 (model.double(), settings.torch_float patched) -> the truth both fp32
         implementations are measured against (three-number parity report).
``--flow`` / ``--noise`` write the two companion files described at flow_boundary() / noise_envelope() and leave the main files alone.
Stored: the model's state_dict after main.py's initialisation, every draw in draw order, the four loss terms and the gradient of every
named parameter for both runs; config 3 additionally the T = 64 forecast (ODEGPVAE.forward(X, T_custom=64)): latent trajectories
and per-frame reconstruction sums.  The input batch is regenerated from its seed on the GPU box (checksum stored).
"""
import os
import sys
import time

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
from oracle import elbo_harness as EH  # noqa: E402
from oracle import glyph  # noqa: E402
from oracle import reference_harness as rh  # noqa: E402

OUT = os.path.join(os.path.dirname(HERE), "tests", "golden")
CASES = [("cfg1", "euler"), ("cfg1", "rk4"), ("cfg2", "rk4"), ("cfg3", "euler"), ("cfg3", "rk4")]
MODEL_SEED, DRAW_SEED, NOISE_SEED = 121, 4242, 777


class NoiseRecorder:
    def __init__(self, seed):
        self.rs, self.log = np.random.RandomState(seed), []

    def __call__(self, shape):
        v = self.rs.normal(size=shape).astype(np.float32)
        self.log.append(v)
        return torch.tensor(v)


class FlowSpy:
    """records what crosses the flow boundary of one compute_loss: z0 (the encoder sample fed to every MC sample's rollout), the latent
    trajectories ztL (L,N,T,D_s), and -- after backward -- dL/dztL and the part of dL/dz0 that arrives through the rollouts"""

    def __init__(self, flow_cls):
        self.cls, self.orig, self.z0, self.zt, self.G, self.dz0 = flow_cls, flow_cls.forward, None, [], [], None

    def __enter__(self):
        spy = self

        def forward(flow, z0, ts):
            if spy.z0 is None:
                spy.z0 = z0.detach().clone()
            zin = z0.detach().clone().requires_grad_(True)     # a branch point of our own: its gradient is the flow's share of dL/dz0
            zt = spy.orig(flow, zin + (z0 - z0.detach()), ts)
            i = len(spy.zt)
            spy.zt.append(zt.detach().clone())
            spy.G.append(None)
            zt.register_hook(lambda g, i=i: spy.G.__setitem__(i, g.detach().clone()))
            zin.register_hook(lambda g: setattr(spy, "dz0", g.detach().clone() if spy.dz0 is None else spy.dz0 + g.detach()))
            return zt
        self.cls.forward = forward
        return self

    def __exit__(self, *a):
        self.cls.forward = self.orig


def run_reference(ref, cfg, solver, dtype, gp_draws=None, enc_noise=None, forecast_T=None, noise_seed=None, spy=None):
    """one compute_loss + backward of the reference on CPU; records the draws when none are given, replays them otherwise"""
    c = EH.CONFIGS[cfg]
    args = EH.make_args(cfg, solver, "cpu")
    from model.misc.settings import Settings
    saved = Settings.torch_float
    torch.set_default_dtype(dtype)
    Settings.torch_float = property(lambda self: dtype)
    try:
        torch.set_default_dtype(torch.float32)        # initial values are always drawn / rounded in fp32, like main.py
        Settings.torch_float = saved
        model = EH.init_model(ref["create_model"], ref["initialization"], args, MODEL_SEED)
        sd = {k: v.detach().clone() for k, v in model.state_dict().items()}
        if dtype == torch.float64:
            torch.set_default_dtype(dtype)
            Settings.torch_float = property(lambda self: dtype)
            model.double()
        X = torch.tensor(EH.inputs(cfg), dtype=dtype)
        if gp_draws is None:
            rec = rh.seed_draws(DRAW_SEED)
            noise = NoiseRecorder(NOISE_SEED if noise_seed is None else noise_seed)
        else:
            rec = EH.Replay(gp_draws, dtype)
            for mod in ("kernels", "svpy"):
                ref[mod].sample_normal = rec
                ref[mod].sample_uniform = rec
            noise = EH.Replay(enc_noise, dtype)
        EH.patch_encoder_noise(ref["vae"], noise)
        if spy is not None:
            with spy:
                scal, grads = EH.run_loss(ref["create_model"], model, X, c["L"])
        else:
            scal, grads = EH.run_loss(ref["create_model"], model, X, c["L"])
        out = dict(sd=sd, scal=scal, grads=grads)
        if gp_draws is None:
            out["gp_draws"] = [v for _, v in rec.log]
            out["enc_noise"] = list(noise.log)
        if forecast_T:
            # forward-only long rollout (plots / evaluation path, main.py:233-244): fresh draws continue the same streams
            ztl = {}
            orig = model.sample_trajectories

            def spy(z0, T, L=1):
                ztl["z"] = orig(z0, T, L)
                return ztl["z"]
            model.sample_trajectories = spy
            n_gp, n_noise = (len(rec.log), len(noise.log)) if gp_draws is None else (rec.i, noise.i)
            with torch.no_grad():
                Xrec, _, _ = model(X, 1, T_custom=forecast_T)
            out["fc_ztL"] = ztl["z"].detach().double().numpy()
            out["fc_xrec_sums"] = Xrec.detach().double().sum((3, 4, 5)).numpy()
            if gp_draws is None:
                out["fc_gp_draws"] = [v for _, v in rec.log[n_gp:]]
                out["fc_enc_noise"] = list(noise.log[n_noise:])
        return out
    finally:
        torch.set_default_dtype(torch.float32)
        Settings.torch_float = saved


N_NOISE = {"cfg1": 16, "cfg2": 6, "cfg3": 16}


def flow_boundary(ref):
    """tests/golden/elbo_flow_<cfg>_<solver>.npz: the fp64 truth AT THE FLOW BOUNDARY of the same compute_loss the main goldens hold --
    z0, ztL, dL/dztL and the rollouts' share of dL/dz0.  The conv decoder is piecewise linear (ReLU): a 1e-6 change of ztL that flips one
    unit moves every gradient of the model by ~1e-3 (measured: config 1 / rk4, one unit of 2,163,200), so the full-model gradient is only
    comparable at 1e-4 where no unit flips; with dL/dztL taken from the truth the flow's own forward and backward are comparable flat."""
    rel = lambda a, b: float(np.linalg.norm(a - b) / max(np.linalg.norm(b), 1e-300))
    for cfg, solver in CASES:
        g = np.load(os.path.join(OUT, "elbo_%s_%s.npz" % (cfg, solver)))
        n_gp, n_noise = int(g["n_gp_draws"]), int(g["n_enc_noise"])
        draws = [g["gp_draw/%03d" % i] for i in range(n_gp)]
        noise = [g["enc_noise/%03d" % i] for i in range(n_noise)]
        spy = FlowSpy(ref["flow"].Flow)
        r64 = run_reference(ref, cfg, solver, torch.float64, gp_draws=draws, enc_noise=noise, spy=spy)
        total = np.sqrt(sum(float(np.sum(g["ref64/grad/" + k].astype(np.float64) ** 2)) for k in r64["grads"]))
        worst = max(rel(r64["grads"][k].astype(np.float32), g["ref64/grad/" + k]) for k in r64["grads"]
                    if np.linalg.norm(g["ref64/grad/" + k]) > 1e-9 * total)     # (conv biases in front of a BatchNorm: exact zeros + noise)
        assert worst < 1e-6, (cfg, solver, worst)      # the same run as the stored truth
        out = dict(z0=spy.z0.numpy(), ztL=torch.stack(spy.zt).numpy(), G=torch.stack(spy.G).numpy(), dz0=spy.dz0.numpy(),
                   kl_gp=np.float64(r64["scal"]["kl_gp"]))
        np.savez_compressed(os.path.join(OUT, "elbo_flow_%s_%s.npz" % (cfg, solver)), **out)
        print("elbo_flow_%s_%s: ztL %s |G| %.3e |dz0| %.3e (stored truth reproduced to %.1e)" % (cfg, solver, out["ztL"].shape, np.linalg.norm(out["G"]),
                                                                                             np.linalg.norm(out["dz0"]), worst))


def noise_envelope(ref):
    """tests/golden/elbo_noise_<cfg>.npz: the reference's OWN fp32 error (ref32 vs ref64, same draws) on every parameter gradient over
    N_NOISE fresh draws of the encoders' reparameterisation noise x the solver variants -- samples of one error distribution (rounding
    noise amplified by the whitening solves, and unit flips of the ReLU decoder), of which the main goldens hold a single sample each."""
    rel = lambda a, b: float(np.linalg.norm(a - b) / max(np.linalg.norm(b), 1e-300))
    for cfg in sorted(set(c for c, _ in CASES)):
        t0 = time.time()
        errs = {}
        for solver in [s for c, s in CASES if c == cfg]:
            for j in range(N_NOISE[cfg]):
                r32 = run_reference(ref, cfg, solver, torch.float32, noise_seed=NOISE_SEED + 1 + j)
                r64 = run_reference(ref, cfg, solver, torch.float64, gp_draws=r32["gp_draws"], enc_noise=r32["enc_noise"])
                total = np.sqrt(sum(float(np.sum(v ** 2)) for v in r64["grads"].values()))
                for k in r32["grads"]:
                    errs.setdefault(k, []).append(rel(r32["grads"][k], r64["grads"][k]) if np.linalg.norm(r64["grads"][k]) > 1e-9 * total else 0.0)
        np.savez_compressed(os.path.join(OUT, "elbo_noise_%s.npz" % cfg), **{k: np.array(v) for k, v in errs.items()})
        med = np.median([np.median(v) for v in errs.values()])
        print("elbo_noise_%s: %d samples per parameter, median error %.1e, max %.1e (%.0fs)" % (cfg, len(next(iter(errs.values()))), med,
                                                                                               max(max(v) for v in errs.values()), time.time() - t0))


def main():
    torch.set_num_threads(8)
    ref = rh.load_reference()
    os.makedirs(OUT, exist_ok=True)
    if "--flow" in sys.argv[1:] or "--noise" in sys.argv[1:]:
        if "--flow" in sys.argv[1:]:
            flow_boundary(ref)
        if "--noise" in sys.argv[1:]:
            noise_envelope(ref)
        return
    for cfg, solver in CASES:
        t0 = time.time()
        fT = 64 if cfg == "cfg3" else None
        r32 = run_reference(ref, cfg, solver, torch.float32, forecast_T=fT)
        draws = r32["gp_draws"] + r32.get("fc_gp_draws", [])
        noise = r32["enc_noise"] + r32.get("fc_enc_noise", [])
        r64 = run_reference(ref, cfg, solver, torch.float64, gp_draws=draws, enc_noise=noise, forecast_T=fT)
        out = {"meta": np.array(repr(dict(cfg=cfg, solver=solver, model_seed=MODEL_SEED, x_seed=121, **EH.CONFIGS[cfg]))),
               "x_checksum": np.float64(glyph.checksum(EH.inputs(cfg)))}
        for k, v in r32["sd"].items():
            out["sd/" + k] = v.numpy()
        out["n_gp_draws"], out["n_enc_noise"] = np.int64(len(r32["gp_draws"])), np.int64(len(r32["enc_noise"]))
        for i, v in enumerate(draws):
            out["gp_draw/%03d" % i] = v
        for i, v in enumerate(noise):
            out["enc_noise/%03d" % i] = v
        for tag, r in (("ref32", r32), ("ref64", r64)):
            for k, v in r["scal"].items():
                out["%s/%s" % (tag, k)] = np.float64(v)
            for k, v in r["grads"].items():
                out["%s/grad/%s" % (tag, k)] = v.astype(np.float32)      # fp32 storage: 6e-8 relative, far below the 1e-4 bars
            if fT:
                out["%s/fc_ztL" % tag] = r["fc_ztL"].astype(np.float32)
                out["%s/fc_xrec_sums" % tag] = r["fc_xrec_sums"]
        name = "elbo_%s_%s" % (cfg, solver)
        np.savez_compressed(os.path.join(OUT, name + ".npz"), **out)
        rel = lambda a, b: float(np.linalg.norm(a - b) / max(np.linalg.norm(b), 1e-300))
        worst = max((rel(r32["grads"][k], r64["grads"][k]), k) for k in r32["grads"])
        print("%s: %.1fs loss %.6f (fp64 %.6f) nlhood %.4f kl_reg %.5f kl_gp %.5f | ref32-vs-ref64 loss %.1e, worst grad %.1e (%s) | %d KB" % (
            name, time.time() - t0, r32["scal"]["loss"], r64["scal"]["loss"], r32["scal"]["nlhood"], r32["scal"]["kl_reg"], r32["scal"]["kl_gp"],
            abs(r32["scal"]["loss"] - r64["scal"]["loss"]) / abs(r64["scal"]["loss"]), worst[0], worst[1],
            os.path.getsize(os.path.join(OUT, name + ".npz")) // 1024))


if __name__ == "__main__":
    main()
