"""Generate tests/golden/elbo_*.npz from the LIVE reference's full model.  TEST INFRASTRUCTURE ONLY.

Run in the build container (needs /root/reference):   python -m oracle.make_golden_elbo
For each BASELINE config 1-3 (oracle/elbo_harness.CONFIGS) and solver the unmodified reference -- build_model, ODEGPVAE, VAE,
SVGP_Layer, Flow, compute_loss (experiments/model/create_model.py:9-35,61-73) with torchdiffeq replaced by oracle/solvers.py --
runs one training-step loss + backward on CPU on the synthetic rotating-digit batch (oracle/glyph.py, seed 121), twice:
  ref32: exactly as the reference runs (fp32 parameters, fp32 draws) -> the like-for-like target;
  ref64: the same code, parameters and draws in float64 (model.double(), settings.torch_float patched) -> the truth both fp32
         implementations are measured against (three-number parity report).
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


def run_reference(ref, cfg, solver, dtype, gp_draws=None, enc_noise=None, forecast_T=None):
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
            noise = NoiseRecorder(NOISE_SEED)
        else:
            rec = EH.Replay(gp_draws, dtype)
            for mod in ("kernels", "svpy"):
                ref[mod].sample_normal = rec
                ref[mod].sample_uniform = rec
            noise = EH.Replay(enc_noise, dtype)
        EH.patch_encoder_noise(ref["vae"], noise)
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


def main():
    torch.set_num_threads(8)
    ref = rh.load_reference()
    os.makedirs(OUT, exist_ok=True)
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
