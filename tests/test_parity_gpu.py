"""GPU parity tests proper: the CUDA path (through the C ABI) against the CPU oracle on the same
seeded inputs.  Tolerances: fp32 CUDA-core mode — loss/outputs 2e-5 abs (fp32 re-association only);
TF32 tensor-core mode — stated per test (TF32 operands have a 10-bit mantissa).  Integer results
(ranks) are compared exactly."""
import numpy as np
import pytest
import torch

from helpers import (CONFIGS, flat_from_named, full_params, seg_slices, small_batch, small_params, torch_batch,
                     torch_params)
from oracle import philox
from oracle import visdial_oracle as O
from visdial_b200 import VD_MATH_FP32, VD_MATH_TF32, Batch, Engine, Model, init_parameters
from visdial_b200.synthetic import make_batch

pytestmark = pytest.mark.gpu

SEED, ITER = 11, 3


def _assert_close(a, b, atol, rtol, what):
    a = np.asarray(a, dtype=np.float64).reshape(-1)
    b = np.asarray(b, dtype=np.float64).reshape(-1)
    assert a.shape == b.shape, what
    scale = max(float(np.abs(b).max()), 1e-30)
    err = float(np.abs(a - b).max())
    assert err <= atol + rtol * scale, "%s: max err %.3e (scale %.3e, atol %.1e rtol %.1e)" % (what, err, scale, atol, rtol)


def _compare_grads(p, g_eng, g_ref, atol, rtol):
    sl = seg_slices(p)
    for name, s in sl.items():
        _assert_close(g_eng[s], g_ref[name].numpy(), atol, rtol, "grad " + name)


def _run_engine(p, flat, nb, mode, training, fused=True):
    eng = Engine(p)
    eng.set_math_mode(mode)
    eng.set_parameters(flat)
    eng.set_training(training)
    eng.set_dropout_seed(SEED, ITER)
    eng.zero_grad()
    return eng


@pytest.mark.parametrize("enc,dec", CONFIGS)
def test_eval_forward_matches_oracle_fp32(enc, dec):
    p = small_params(enc, dec)
    flat = init_parameters(p, seed=3)
    nb = small_batch(p, B=3)
    eng = _run_engine(p, flat, nb, VD_MATH_FP32, 0)
    b = Batch(nb)
    encOut = eng.encoder_forward(b).numpy()
    eng.forward_connect()
    decOut = eng.decoder_forward(b).numpy()
    loss = eng.criterion_forward(b)
    ref = O.forward_backward(O.Ctx(train=False), p, torch_params(p, flat), torch_batch(nb), only_forward=True)
    _assert_close(encOut, ref["encOut"].numpy(), 2e-5, 1e-5, "encOut")
    _assert_close(decOut, ref["decOut"].numpy(), 5e-5, 1e-5, "decOut")
    assert abs(loss - ref["loss"]) <= 1e-4 * max(1.0, abs(ref["loss"]))
    eng.close()


@pytest.mark.parametrize("enc,dec", CONFIGS)
def test_train_forward_backward_matches_oracle_fp32(enc, dec):
    """Training mode with dropout ON: the oracle is given the engine's Philox masks."""
    p = small_params(enc, dec)
    flat = init_parameters(p, seed=3)
    nb = small_batch(p, B=3)
    eng = _run_engine(p, flat, nb, VD_MATH_FP32, 1)
    loss = eng.forward_backward(Batch(nb))
    g = eng.get_gradients()
    psite = {O.SITE_FUSION: p["dropout"]}
    ref = O.forward_backward(O.Ctx(train=True, mask_fn=philox.make_mask_fn(SEED, ITER, psite)), p, torch_params(p, flat),
                             torch_batch(nb))
    assert abs(loss - ref["loss"]) <= 1e-4 * max(1.0, abs(ref["loss"])), (loss, ref["loss"])
    _compare_grads(p, g, ref["grads"], 2e-5, 2e-4)
    eng.close()


@pytest.mark.parametrize("enc,dec", [("mn-att-ques-im-hist", "disc"), ("hrea-ques-im-hist", "gen")])
def test_model_protocol_equals_fused_path(enc, dec):
    """Model:forwardBackward driven call-by-call (model.lua:297-337) == vd_forward_backward."""
    p = small_params(enc, dec)
    nb = small_batch(p, B=2)
    m = Model(p, seed=5)
    m.engine.set_math_mode(VD_MATH_FP32)
    m.engine.set_dropout_seed(SEED, ITER)
    m.wrapper.zeroGradParameters()
    l1 = m.forwardBackward(nb)
    g1 = m.engine.get_gradients()
    m.wrapper.zeroGradParameters()
    l2 = m.engine.forward_backward(Batch(nb))
    g2 = m.engine.get_gradients()
    assert l1 == pytest.approx(l2, rel=1e-6)
    _assert_close(g1, g2, 1e-6, 1e-4, "fused vs protocol grads")   # atomics re-order fp32 sums
    m.engine.close()


@pytest.mark.parametrize("enc,dec", [("mn-att-ques-im-hist", "disc"), ("lf-ques", "gen")])
def test_train_iteration_adam_matches_oracle(enc, dec):
    """Model:trainIteration = zeroGrad, fwd/bwd, clamp(-5,5), adam, lr decay (model.lua:66-106)."""
    p = small_params(enc, dec)
    nb = small_batch(p, B=2)

    class DL:
        def getTrainBatch(self, params):
            return nb
    m = Model(p, seed=5)
    m.engine.set_math_mode(VD_MATH_FP32)
    flat0 = m.engine.get_parameters()
    W = torch.from_numpy(flat0.copy())
    state = {}
    lr = p["learningRate"]
    for it in range(1, 4):
        lr_used = m.optims["learningRate"]
        m.trainIteration(DL())
        psite = {O.SITE_FUSION: p["dropout"]}
        ref = O.forward_backward(O.Ctx(train=True, mask_fn=philox.make_mask_fn(1234, it, psite)), p,
                                 torch_params(p, W.numpy()), torch_batch(nb))
        dW = torch.from_numpy(flat_from_named(p, ref["grads"]))
        O.clamp_adam(W, dW, state, lr)
        W[:p["embedSize"]] = 0         # pad row is re-zeroed at the next forward (LookupTableMaskZero)
        lr = O.decay_lr(lr, p)
        got = m.engine.get_parameters()
        got[:p["embedSize"]] = 0
        # Adam normalises the step to ~lr: compare in units of lr
        assert float(np.abs(got - W.numpy()).max()) < 0.05 * p["learningRate"], it
        # The optimiser formula itself, isolated from gradient noise: replay optim_updates.lua:62-91 in fp64 on the gradient
        # the engine consumed (get_gradients after the step = clamped dW) and compare m, v, t and the weights directly —
        # an eps-placement or bias-correction slip moves small-gradient weights by O(lr), far outside these bounds.
        g = m.engine.get_gradients().astype(np.float64)
        if it == 1:
            m_ref, v_ref, w_ref = np.zeros_like(g), np.zeros_like(g), flat0.astype(np.float64)
        w_ref[:p["embedSize"]] = 0                  # LookupTableMaskZero re-zeroes the pad row at every forward
        m_ref = 0.9 * m_ref + 0.1 * g
        v_ref = 0.999 * v_ref + 0.001 * g * g
        step = lr_used * np.sqrt(1 - 0.999 ** it) / (1 - 0.9 ** it)
        w_ref = w_ref - step * m_ref / (np.sqrt(v_ref) + 1e-8)
        em, ev, et = m.engine.optim_state()
        assert et == it
        assert float(np.abs(em - m_ref).max()) <= 1e-6 * float(np.abs(m_ref).max()) + 1e-12, it
        assert float(np.abs(ev - v_ref).max()) <= 3e-5 * float(np.abs(v_ref).max()) + 1e-20, (it, float(np.abs(ev - v_ref).max()), float(np.abs(v_ref).max()))
        raw = m.engine.get_parameters().astype(np.float64)
        raw[:p["embedSize"]] = 0
        w_chk = w_ref.copy(); w_chk[:p["embedSize"]] = 0
        assert float(np.abs(raw - w_chk).max()) < 2e-3 * lr_used + 3e-7, (it, float(np.abs(raw - w_chk).max()))
    assert m.optims["learningRate"] == pytest.approx(lr)
    m.engine.close()


@pytest.mark.parametrize("enc,dec", [("mn-att-ques-im-hist", "disc"), ("lf-ques-im-hist", "disc"),
                                     ("lf-ques", "gen"), ("hrea-ques-im-hist", "gen")])
def test_retrieve_ranks_bit_exact(enc, dec):
    p = small_params(enc, dec)
    flat = init_parameters(p, seed=3)
    nb = small_batch(p, B=3, gen_eval=(dec == "gen"))
    eng = _run_engine(p, flat, nb, VD_MATH_FP32, 1)
    r_gt = eng.retrieve(Batch(nb), use_gt=True)
    r_all = eng.retrieve(Batch(nb), use_gt=False)
    P = torch_params(p, flat)
    tb = torch_batch(nb)
    ref_gt = O.retrieve_batch(O.Ctx(), p, P, tb, use_gt=True).numpy()
    ref_all = O.retrieve_batch(O.Ctx(), p, P, tb, use_gt=False).numpy()
    assert np.array_equal(r_gt, ref_gt)
    assert np.array_equal(r_all, ref_all)
    eng.close()


def test_rank_kernel_ties_and_permutation():
    p = small_params("lf-ques", "disc", numOptions=100)
    eng = Engine(p)
    import ctypes as C
    rng = np.random.default_rng(0)
    s = rng.standard_normal((64, 100)).astype(np.float32)
    s[:, 10] = s[:, 3]            # exact ties: lower index wins
    s[5] = 0.0                    # a fully tied row ranks 1..100 in index order
    from visdial_b200 import engine as E
    from visdial_b200._lib import check
    dev_s, dev_r = C.c_void_p(), C.c_void_p()
    lib = eng.lib
    # use the parameter gradient buffer as scratch device memory
    w, dw = eng.param_buffers()
    check(lib.vd_memcpy_h2d(eng.h, dw, s.ctypes.data, s.nbytes))
    ranks_dev = dw + s.nbytes
    check(lib.vd_compute_ranks(eng.h, dw, 64, None, ranks_dev))
    out = np.empty((64, 100), dtype=np.int32)
    check(lib.vd_memcpy_d2h(eng.h, out.ctypes.data, ranks_dev, out.nbytes))
    ref = O.compute_ranks(torch.from_numpy(s)).numpy()
    assert np.array_equal(out, ref)
    assert out[5].tolist() == list(range(1, 101))
    assert all(sorted(r.tolist()) == list(range(1, 101)) for r in out)
    eng.close()


def test_headline_shapes_against_oracle_fp32():
    """mn-att-ques-im-hist + disc at the reference's real layer sizes (E=300, H=512, pool5 14x14x512,
    100 options x 20 tokens, V=10000), B=2 so the oracle finishes in seconds."""
    p = full_params("mn-att-ques-im-hist", "disc")
    flat = init_parameters(p, seed=3)
    nb = make_batch(p, 2, seed=5)
    eng = _run_engine(p, flat, nb, VD_MATH_FP32, 1)
    loss = eng.forward_backward(Batch(nb))
    g = eng.get_gradients()
    ref = O.forward_backward(O.Ctx(train=True, mask_fn=philox.make_mask_fn(SEED, ITER), structure="batched"), p,
                             torch_params(p, flat), torch_batch(nb))
    assert abs(loss - ref["loss"]) <= 2e-4 * max(1.0, abs(ref["loss"])), (loss, ref["loss"])
    _compare_grads(p, g, ref["grads"], 1e-5, 2e-3)
    ranks = eng.retrieve(Batch(nb), use_gt=False)
    ref_r = O.retrieve_batch(O.Ctx(structure="batched"), p, torch_params(p, flat), torch_batch(nb), use_gt=False).numpy()
    # ranks are exact wherever the oracle's score gap to the neighbours exceeds the fp32 noise floor
    sc = O.forward_backward(O.Ctx(structure="batched"), p, torch_params(p, flat), torch_batch(nb), only_forward=True)["decOut"].numpy()
    gap = np.abs(sc[:, :, None] - sc[:, None, :]) + np.eye(100)[None] * 1e9
    safe = gap.min(2) > 2e-5                      # options whose score is isolated from every other option
    assert safe.mean() > 0.9
    assert np.array_equal(ranks[safe], ref_r[safe])
    eng.close()


def test_full_size_properties_baseline_batch():
    """BASELINE config 4 at B=32 (N=320 rounds): size-independent properties."""
    p = full_params("mn-att-ques-im-hist", "disc")
    flat = init_parameters(p, seed=3)
    nb = make_batch(p, 32, seed=5)
    eng = Engine(p)
    eng.set_parameters(flat)
    eng.set_training(1)
    eng.set_dropout_seed(SEED, ITER)
    eng.zero_grad()
    l1 = eng.forward_backward(Batch(nb))
    g1 = eng.get_gradients()
    assert np.isfinite(l1) and abs(l1 - np.log(100)) < 1.0          # random init: loss ~ ln(100)
    assert np.isfinite(g1).all() and np.abs(g1).max() > 0
    # linearity of gradient accumulation: a second backward without zeroGrad doubles dW
    l2 = eng.forward_backward(Batch(nb))
    g2 = eng.get_gradients()
    assert l2 == pytest.approx(l1, rel=1e-5)
    _assert_close(g2, 2 * g1, 1e-6, 2e-3, "accumulated grads")
    # ranks: every row is a permutation of 1..100; the rank of the gt equals the full-rank entry
    r_all = eng.retrieve(Batch(nb), use_gt=False)
    r_gt = eng.retrieve(Batch(nb), use_gt=True)
    assert (np.sort(r_all, 1) == np.arange(1, 101)[None, :]).all()
    assert np.array_equal(r_gt, r_all[np.arange(320), nb["answer_ind"] - 1])
    # eval is idempotent and independent of the dropout seed
    eng.set_dropout_seed(99, 99)
    assert np.array_equal(eng.retrieve(Batch(nb), use_gt=False), r_all)
    eng.close()


def test_errors_are_loud():
    from visdial_b200 import VdError
    p = small_params("mn-att-ques-im-hist", "disc")
    eng = Engine(p)
    nb = small_batch(p, B=2)
    with pytest.raises(VdError):                       # backward before forward
        eng.decoder_backward(Batch(nb))
    bad = dict(nb)
    bad.pop("hist")
    with pytest.raises(VdError):                       # encoder needs history
        eng.encoder_forward(Batch(bad))
    eng.set_training(0)
    eng.forward_backward(Batch(nb), only_forward=True)
    with pytest.raises(VdError):                       # no backward from an eval-mode forward
        eng.criterion_backward(Batch(nb)) or eng.decoder_backward(Batch(nb))
    eng.close()


def test_checkpoint_round_trip_resumes_training(tmp_path):
    """train.lua:99-102 / :33-34,78-80: a model restored from the .t7 checkpoint (weights, learning rate, Adam state)
    continues exactly like the one that wrote it; the model_final.t7 form carries float weights only."""
    from visdial_b200 import t7
    p = small_params("mn-att-ques-im-hist", "disc", batchSize=3)

    class DL:
        def __init__(self): self.i = 0
        def getTrainBatch(self, params):
            self.i += 1
            return small_batch(params, B=3, seed=40 + self.i)

    a = Model(p, seed=3)
    a.engine.set_math_mode(VD_MATH_FP32)
    dl = DL()
    for _ in range(2):
        a.trainIteration(dl)
    path = str(tmp_path / "model_epoch_1.t7")
    a.save(path)
    a.save(str(tmp_path / "model_final.t7"), final=True)

    raw = t7.load(path)
    assert set(raw) == {"modelW", "optims", "modelParams", "layout"} and raw["optims"]["t"] == 2
    assert raw["modelParams"]["encoder"] == "mn-att-ques-im-hist"
    final = t7.load(str(tmp_path / "model_final.t7"))
    assert "optims" not in final and np.array_equal(final["modelW"], raw["modelW"])

    b = Model(raw["modelParams"] | {"gpuid": 0}, seed=99)                 # evaluate.lua:61-91: params come from the file
    b.engine.set_math_mode(VD_MATH_FP32)
    b.load(path, restore_adam_state=True)
    b.iteration = a.iteration
    assert b.optims["learningRate"] == a.optims["learningRate"]
    np.testing.assert_array_equal(b.engine.get_parameters(), a.engine.get_parameters())
    dl_b = DL(); dl_b.i = dl.i
    la, lb = a.trainIteration(dl), b.trainIteration(dl_b)
    assert abs(la - lb) <= 1e-6 * max(1.0, abs(la))
    np.testing.assert_allclose(b.engine.get_parameters(), a.engine.get_parameters(), rtol=0, atol=2.1e-3)
    assert float(np.abs(b.engine.get_parameters() - a.engine.get_parameters()).mean()) < 1e-6
    with pytest.raises(Exception):
        Model(small_params("lf-ques", "gen"), seed=1).load(path)           # parameter count mismatch is an error
    a.engine.close(); b.engine.close()
