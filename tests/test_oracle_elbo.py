"""CPU: the full-model golden vectors (tests/golden/elbo_*.npz) against the reference itself -- the LIVE tree under /root/reference
in the build container, its verbatim copy under oracle/_ref elsewhere -- and the synthetic rotating-digit generator."""
import os
import sys

import numpy as np
import pytest
import torch

from oracle import elbo_harness as EH
from oracle import glyph
from oracle import reference_harness as rh

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
from test_gpu_elbo import load, rel  # noqa: E402


def test_rotating_glyph_generator_is_deterministic_and_normalised():
    x = glyph.rotating_sequences(5, 16, seed=121)
    assert x.shape == (5, 16, 1, 28, 28) and x.dtype == np.float32
    assert np.array_equal(x, glyph.rotating_sequences(5, 16, seed=121))
    raw = x * np.float32(glyph.MNIST_STD) + np.float32(glyph.MNIST_MEAN)          # data/utils.py:13-14 undone
    assert raw.min() > -1e-6 and raw.max() < 1 + 1e-6 and raw.max() > 0.9
    # frame t is frame 0 rotated by t * 360 / 15 degrees (data/mnist.py:174-175): a full turn brings the digit back
    assert rel(x[:, 15], x[:, 0]) < 0.05 and rel(x[:, 7], x[:, 0]) > 0.5
    for cfg in ("cfg1", "cfg3"):
        g = load(cfg, "rk4")
        assert abs(glyph.checksum(EH.inputs(cfg)) - float(g["x_checksum"])) <= 1e-9 * abs(float(g["x_checksum"]))


@pytest.mark.skipif(not rh.reference_available(), reason="reference tree not present")
@pytest.mark.parametrize("cfg,solver", [("cfg1", "rk4"), ("cfg3", "euler")])
def test_golden_elbo_reproduces_from_the_reference(cfg, solver):
    """replaying the stored draws through the reference's own classes gives the stored ELBO terms and gradients back"""
    ref = rh.load_reference()
    g = load(cfg, solver)
    args = EH.make_args(cfg, solver, "cpu")
    model = EH.init_model(ref["create_model"], ref["initialization"], args, g["meta"]["model_seed"])
    model.load_state_dict({k[3:]: torch.tensor(v) for k, v in g.items() if k.startswith("sd/")}, strict=True)
    draws = EH.Replay([g["gp_draw/%03d" % i] for i in range(int(g["n_gp_draws"]))])
    noise = EH.Replay([g["enc_noise/%03d" % i] for i in range(int(g["n_enc_noise"]))])
    saved = (ref["kernels"].sample_normal, ref["kernels"].sample_uniform, ref["svpy"].sample_normal, ref["vae"].Encoder.sample)
    try:
        ref["kernels"].sample_normal = ref["kernels"].sample_uniform = ref["svpy"].sample_normal = draws
        EH.patch_encoder_noise(ref["vae"], noise)
        scal, grads = EH.run_loss(ref["create_model"], model, torch.tensor(EH.inputs(cfg)), g["meta"]["L"])
    finally:
        ref["kernels"].sample_normal, ref["kernels"].sample_uniform, ref["svpy"].sample_normal, ref["vae"].Encoder.sample = saved
    for k in ("loss", "nlhood", "kl_reg", "kl_gp"):
        assert abs(scal[k] - float(g["ref32/" + k])) <= 2e-6 * abs(float(g["ref32/" + k])), k
    total = np.sqrt(sum(float(np.sum(v.astype(np.float64) ** 2)) for k, v in g.items() if k.startswith("ref32/grad/")))
    for k, v in grads.items():
        want = g["ref32/grad/" + k]
        if np.linalg.norm(want) > 1e-9 * total:
            assert rel(v, want) < 1e-4, k          # thread-count dependent fp32 summation order only
