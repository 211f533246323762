"""Committed golden fixtures (tests/golden/*.npz, made by tests/golden/make_golden.py from the fp64 oracle).
CPU: the fp32 oracle still reproduces them (guards against oracle drift).  GPU: the CUDA path reproduces them
through the C ABI without the oracle in the loop."""
import glob
import os

import numpy as np
import pytest
import torch

from helpers import flat_from_named, small_params, torch_batch, torch_params
from oracle import philox
from oracle import visdial_oracle as O

GOLD = sorted(g for g in glob.glob(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "*.npz"))
              if not os.path.basename(g).startswith("c4_b32"))      # the benched-size fixture has its own test file


def _load(path):
    z = np.load(path)
    enc, dec = os.path.basename(path)[:-4].split("__")
    p = small_params(enc, dec)
    batch = {k[6:]: z[k] for k in z.files if k.startswith("batch_")}
    return p, z, batch


def _close(a, b, rtol, atol, what):
    a = np.asarray(a, np.float64).ravel()
    b = np.asarray(b, np.float64).ravel()
    err = np.abs(a - b).max()
    assert err <= atol + rtol * np.abs(b).max(), "%s: err %.3e vs scale %.3e" % (what, err, np.abs(b).max())


def test_fixtures_exist():
    assert len(GOLD) == 11          # one per encoder of encoders/*.lua (tests/golden/make_golden.py)


@pytest.mark.parametrize("path", GOLD, ids=[os.path.basename(g)[:-4] for g in GOLD])
def test_oracle_fp32_reproduces_golden(path):
    p, z, batch = _load(path)
    P = torch_params(p, z["flat"])
    tb = torch_batch(batch)
    ev = O.forward_backward(O.Ctx(train=False), p, P, tb, only_forward=True)
    _close(ev["encOut"].numpy(), z["eval_encOut"], 1e-4, 1e-5, "encOut")
    _close(ev["decOut"].numpy(), z["eval_decOut"], 1e-4, 1e-5, "decOut")
    assert ev["loss"] == pytest.approx(float(z["eval_loss"]), rel=1e-4)
    psite = {O.SITE_FUSION: p["dropout"]}
    tr = O.forward_backward(O.Ctx(train=True, mask_fn=philox.make_mask_fn(5, 2, psite)), p, P, tb)
    assert tr["loss"] == pytest.approx(float(z["train_loss"]), rel=1e-4)
    _close(flat_from_named(p, tr["grads"]), z["train_grad"], 2e-3, 1e-6, "grads")
    if "ranks" in z.files:
        assert np.array_equal(O.compute_ranks(ev["decOut"]).numpy(), z["ranks"])


@pytest.mark.gpu
@pytest.mark.parametrize("path", GOLD, ids=[os.path.basename(g)[:-4] for g in GOLD])
def test_gpu_reproduces_golden(path):
    from visdial_b200 import VD_MATH_FP32, Batch, Engine
    p, z, batch = _load(path)
    eng = Engine(p)
    eng.set_math_mode(VD_MATH_FP32)
    eng.set_parameters(z["flat"])
    eng.set_training(0)
    b = Batch(batch)
    enc = eng.encoder_forward(b).numpy()
    eng.forward_connect()
    dec = eng.decoder_forward(b).numpy()
    loss = eng.criterion_forward(b)
    _close(enc, z["eval_encOut"], 1e-4, 2e-5, "encOut")
    _close(dec, z["eval_decOut"], 1e-4, 5e-5, "decOut")
    assert loss == pytest.approx(float(z["eval_loss"]), rel=1e-4)
    if "ranks" in z.files:
        assert np.array_equal(eng.retrieve(b, use_gt=False), z["ranks"])
    eng.set_training(1)
    eng.set_dropout_seed(5, 2)
    eng.zero_grad()
    tl = eng.forward_backward(b)
    assert tl == pytest.approx(float(z["train_loss"]), rel=1e-4)
    _close(eng.get_gradients(), z["train_grad"], 2e-3, 1e-5, "grads")
    eng.close()
