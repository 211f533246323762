"""Generates tests/golden/dataloader/*.npz from oracle/dataloader_oracle.py (no runnable reference, no reference
fixtures: DESIGN.md §11).  Each file holds a small raw corpus in the prepro.py layout, the prepared tensors and two
assembled batches (train-form and test-form).  They pin the oracle against drift and give the GPU tests committed
vectors that travel to the GPU box.  Re-run:  python tests/golden/make_golden_dataloader.py"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(HERE, "..", ".."))
sys.path.insert(0, os.path.join(HERE, ".."))

from helpers import small_params  # noqa: E402
from oracle.dataloader_oracle import DataloaderOracle  # noqa: E402
from visdial_b200.synthetic import make_corpus  # noqa: E402

CASES = [("mn-att-ques-im-hist", "disc"), ("lf-ques-im-hist", "gen"), ("hrea-ques-im-hist", "gen"), ("lf-ques", "disc")]
INDS = np.array([0, 1, 3, 4, 5, 5, 11])


def build(enc, dec):
    p = small_params(enc, dec)
    concat = "lf" in enc and "hist" in enc
    raw = make_corpus(p, 12, 40, seed=77, max_ques_len=8, max_ans_len=6, max_cap_len=14,
                      ques_len_cap=5 if concat else None, ans_len_cap=4 if concat else None)
    V = p["vocabSize"]
    orc = DataloaderOracle(raw, use_history="hist" in enc, concat_history=concat, use_im="im" in enc, start=V - 1, end=V,
                           img_norm=True, att="att" in enc)
    out = {"raw_" + k: v for k, v in raw.items()}
    for k in ("ques_fwd", "ans_in", "ans_out", "opt_in", "opt_out"):
        out["prep_" + k] = getattr(orc, k).astype(np.int32)
    if "hist" in enc:
        out["prep_hist"] = orc.hist.astype(np.int32)
        out["prep_hist_len"] = orc.hist_len.astype(np.int32)
    if "im" in enc:
        out["prep_img_fv"] = orc.img_fv
    for tag, test in (("train", False), ("test", True)):
        b = orc.get_batch(INDS, dec, test_batch=test)
        for k, v in b.items():
            out["%s_%s" % (tag, k)] = v.astype(np.float32 if k == "img_feat" else np.int32)
    return p, out


def main():
    for enc, dec in CASES:
        _, out = build(enc, dec)
        name = os.path.join(HERE, "dataloader", "%s__%s.npz" % (enc, dec))
        np.savez_compressed(name, **out)
        print(name, os.path.getsize(name), "bytes")


if __name__ == "__main__":
    main()
