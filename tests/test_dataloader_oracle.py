"""Known-answer checks of oracle/dataloader_oracle.py against hand-worked examples of the reference's loops
(/root/reference/utils.lua:6-45, dataloader.lua:159-321,378-478).  CPU only."""
import numpy as np
import pytest

from oracle import dataloader_oracle as D
from visdial_b200.synthetic import make_corpus

S, E = 98, 99   # <START>, <END>


def test_right_align_3d_breaks_at_first_empty_round():
    seq = np.array([[[1, 2, 0], [3, 0, 0], [4, 5, 6]],
                    [[7, 0, 0], [0, 0, 0], [8, 9, 0]]])
    lens = np.array([[2, 1, 3], [1, 0, 2]])
    out = D.right_align(seq, lens)
    # dialog 0: every round copied; dialog 1: the loop breaks at round 2, so round 3 stays zero (utils.lua:20-22)
    assert out.tolist() == [[[0, 1, 2], [0, 0, 3], [4, 5, 6]],
                            [[0, 0, 7], [0, 0, 0], [0, 0, 0]]]


def test_right_align_2d_skips_empty_rows():
    out = D.right_align(np.array([[1, 2, 0], [0, 0, 0], [3, 0, 0]]), np.array([2, 0, 1]))
    assert out.tolist() == [[0, 1, 2], [0, 0, 0], [0, 0, 3]]


def test_process_answers_start_end_and_empty():
    ans = np.array([[[5, 6, 0], [0, 0, 0]]])
    a_in, a_out, ln = D.process_answers(ans, np.array([[2, 0]]), S, E)
    assert a_in.tolist() == [[[S, 5, 6, 0], [S, 0, 0, 0]]]
    assert a_out.tolist() == [[[5, 6, E, 0], [E, 0, 0, 0]]]       # empty answer still gets <END> (dataloader.lua:193)
    assert ln.tolist() == [[3, 1]]


def test_process_options_empty_option_has_no_end():
    o_in, o_out, ln = D.process_options(np.array([[5, 6, 0], [0, 0, 0]]), np.array([2, 0]), S, E)
    assert o_in.tolist() == [[S, 5, 6, 0], [S, 0, 0, 0]]
    assert o_out.tolist() == [[5, 6, E, 0], [0, 0, 0, 0]]         # dataloader.lua:306-310
    assert ln.tolist() == [3, 1]


def _tiny():
    # Lq = 2, La = 2 -> per-round history width 4; caption tensor width 4
    ques = np.array([[[1, 2], [3, 0], [4, 0]]])
    ql = np.array([[2, 1, 1]])
    ans = np.array([[[5, 0], [6, 7], [8, 0]]])
    al = np.array([[1, 2, 1]])
    cap = np.array([[9, 10, 11, 0]])
    return cap, np.array([3]), ques, ql, ans, al


def test_history_per_round():
    cap, cl, ques, ql, ans, al = _tiny()
    h, hl, mx = D.process_history(cap, cl, ques, ql, ans, al, False, E, 60)
    assert hl.tolist() == [[3, 3, 3]]
    assert h.tolist() == [[[0, 9, 10, 11], [0, 1, 2, 5], [0, 3, 6, 7]]]
    assert mx == 60


def test_history_concat():
    cap, cl, ques, ql, ans, al = _tiny()
    h, hl, mx = D.process_history(cap, cl, ques, ql, ans, al, True, E, 60)
    assert mx == 12 and h.shape == (1, 3, 12)                     # min(3 * (2 + 2), 300)
    assert hl.tolist() == [[3, 7, 11]]
    assert h[0, 0].tolist() == [0] * 9 + [9, 10, 11]
    assert h[0, 1].tolist() == [0] * 5 + [9, 10, 11, E, 1, 2, 5]
    assert h[0, 2].tolist() == [0] + [9, 10, 11, E, 1, 2, 5, E, 3, 6, 7]


def test_history_concat_overflow_raises():
    n, R, L = 1, 10, 20
    ques = np.ones((n, R, L), dtype=np.int64); ans = np.ones((n, R, L), dtype=np.int64)
    full = np.full((n, R), L)
    with pytest.raises(IndexError):
        D.process_history(np.ones((n, 40), dtype=np.int64), np.array([40]), ques, full, ans, full, True, E, 60)


def test_empty_caption_zeroes_the_whole_history():
    cap, cl, ques, ql, ans, al = _tiny()
    h, hl, _ = D.process_history(cap, np.array([0]), ques, ql, ans, al, False, E, 60)
    assert hl.tolist() == [[0, 3, 3]] and not h.any()             # rightAlign breaks at round 1


def test_batch_trimming_and_option_lookup():
    params = {"vocabSize": 100, "encoder": "lf-ques-im-hist", "imgFeatureSize": 8, "maxQuesCount": 10, "numOptions": 100}
    raw = make_corpus(params, 16, 40, seed=3)
    o = D.DataloaderOracle(raw, use_history=True, concat_history=False, use_im=True, start=S, end=E, img_norm=True)
    inds = np.array([6, 7, 9])
    b = o.get_batch(inds, "disc", test_batch=False)
    Tq = raw["ques_length"][inds].max()
    assert b["ques_fwd"].shape == (3, 10, Tq)
    assert (b["ques_fwd"][:, :, -1] != 0).all()                   # right-aligned: last column always a token
    assert b["answer_in"].shape[2] == raw["ans_length"][inds].max() + 1
    assert b["options"].shape == (30, 100, 20) and b["answer_ind"].shape == (30,)
    i, r, k = 1, 4, 17
    assert (b["options"][i * 10 + r, k] == raw["opt_list"][raw["opt"][inds[i], r, k] - 1]).all()
    assert np.allclose(np.linalg.norm(b["img_feat"], axis=1), 1, atol=1e-5)
    g = o.get_batch(inds, "gen", test_batch=True)
    To = (raw["opt_length"][raw["opt"][inds].reshape(-1) - 1]).max() + 1
    assert g["option_in"].shape == (3, 10, 100, To) and (g["option_in"][..., 0] == S).all()
    assert "options" not in g and g["num_rounds"].shape == (3,)


def test_attention_features_are_permuted_after_normalising():
    params = {"vocabSize": 100, "encoder": "mn-att-ques-im-hist", "imgFeatureSize": 8, "imgSpatialSize": 3}
    raw = make_corpus(params, 8, 20, seed=5)
    x = D.prepare_images(raw["images"], True, True)
    assert x.shape == (8, 3, 3, 8)
    assert np.allclose(np.linalg.norm(x, axis=3), 1, atol=1e-5)   # sum over dim 2 of N x C x S x S = channel norm
    assert np.allclose(x[2, 1, 2] * np.linalg.norm(raw["images"][2, :, 1, 2]), raw["images"][2, :, 1, 2], atol=1e-5)


# ---- committed fixtures (tests/golden/dataloader/*.npz, made by tests/golden/make_golden_dataloader.py) ----------
import glob
import os
import sys

GOLD = sorted(glob.glob(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "dataloader", "*.npz")))


def test_dataloader_fixtures_exist():
    assert len(GOLD) == 4


@pytest.mark.parametrize("path", GOLD, ids=[os.path.basename(g)[:-4] for g in GOLD])
def test_oracle_reproduces_dataloader_golden(path):
    sys.path.insert(0, os.path.join(os.path.dirname(path), ".."))
    from make_golden_dataloader import build
    enc, dec = os.path.basename(path)[:-4].split("__")
    _, out = build(enc, dec)
    z = np.load(path)
    assert sorted(z.files) == sorted(out)
    for k in z.files:
        if z[k].dtype.kind == "f":
            np.testing.assert_allclose(out[k], z[k], rtol=1e-6, atol=1e-7, err_msg=k)
        else:
            assert np.array_equal(out[k], z[k]), k
