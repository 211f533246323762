"""Where do Model.generateAnswers and oracle.generate_answers part ways?  usage: python tools/debug_genans.py <encoder>"""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
from helpers import small_params, torch_batch, torch_params  # noqa: E402
from oracle import dataloader_oracle as D, visdial_oracle as O  # noqa: E402
from visdial_b200 import Model, init_parameters  # noqa: E402
from visdial_b200.dataloader import Dataloader  # noqa: E402
from visdial_b200.synthetic import make_corpus  # noqa: E402

enc = sys.argv[1] if len(sys.argv) > 1 else "hrea-ques-im-hist"
params = small_params(enc, "gen", vocabSize=9)
concat = "lf" in enc and "hist" in enc
raw = make_corpus(params, 12, 40, seed=77, max_ques_len=8, max_ans_len=6, max_cap_len=14,
                  ques_len_cap=5 if concat else None, ans_len_cap=4 if concat else None)
V = params["vocabSize"]
orc = D.DataloaderOracle(raw, use_history="hist" in enc, concat_history=concat, use_im="im" in enc, start=V - 1, end=V,
                         img_norm=True, att="att" in enc)
model = Model(dict(params, batchSize=1), seed=3)
model.engine.set_math_mode(1)
flat = init_parameters(params, seed=3)
model.engine.set_parameters(flat)
opt = dict(params, useHistory="hist" in enc, concatHistory="lf" in enc, useIm="im" in enc, maxHistoryLen=60, imgNorm=1)
dl = Dataloader(model.engine).initialize(opt, ["val"], {"val": raw})
P = torch_params(params, flat)
for conv in (0, 3):
    model.wrapper.evaluate()
    batch = dl.getIndexData(np.array([conv]), model.params, "val")
    got_b = batch.numpy() if hasattr(batch, "numpy") else None
    ref_b = orc.get_index_data(np.array([conv]))
    if got_b is not None:
        for k in ref_b:
            if k in got_b:
                a, b = np.asarray(got_b[k], np.float64), np.asarray(ref_b[k], np.float64)
                print(conv, "batch", k, a.shape, b.shape, "maxdiff", float(np.abs(a - b).max()) if a.shape == b.shape else "SHAPE")
    encOut = model.forwardBackward(batch, True, True).numpy()
    tb = torch_batch(ref_b)
    with torch.no_grad():
        inputs = O.prepare_inputs(params, tb)
        eo, state = O.ENCODERS[enc](O.Ctx(), params, P, inputs)
    print(conv, "encOut maxdiff", float(np.abs(encOut - eo.numpy()).max()), "scale", float(eo.abs().max()))
    rl = state.get("rnnLayers")
    for l in range(2):
        h, c = model.engine.encoder_rnn_state(l, encOut.shape[0])
        if h is not None and rl is not None:
            print(conv, "layer", l, "h diff", float(np.abs(h.numpy() - rl[l][0][-1].numpy()).max()), "c diff",
                  float(np.abs(c.numpy() - rl[l][1][-1].numpy()).max()))
model.engine.close()
