"""Model:generateAnswers (model.lua:432-613) wall time per dialog: beam search (beamSize 5, beamLen 20) over the 10 rounds of a
dialog, decoder stepped on the device through vd_gen_decoder_step, hypothesis bookkeeping on the host like the reference.
usage: python tools/bench_generate.py [encoder] [dialogs]"""
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from visdial_b200 import VD_MATH_F16, Model  # noqa: E402
from visdial_b200.dataloader import Dataloader  # noqa: E402
from visdial_b200.engine import DEFAULT_PARAMS, derive_flags  # noqa: E402
from visdial_b200.synthetic import make_corpus  # noqa: E402

enc = sys.argv[1] if len(sys.argv) > 1 else "hrea-ques-im-hist"
nd = int(sys.argv[2]) if len(sys.argv) > 2 else 8
p = dict(DEFAULT_PARAMS)
p.update(encoder=enc, decoder="gen", vocabSize=10000, imgFeatureSize=512 if "att" in enc else 4096, batchSize=1)
p = derive_flags(p)
raw = make_corpus(p, nd, 2000, seed=5)
m = Model(p, seed=3)
m.engine.set_math_mode(VD_MATH_F16)
dl = Dataloader(m.engine).initialize(dict(p, maxHistoryLen=60), ["val"], {"val": raw})
m.generateAnswers(dl, "val", {"beamSize": 5, "beamLen": 20, "maxThreads": 1}, strict=False)
t0 = time.perf_counter()
out = m.generateAnswers(dl, "val", {"beamSize": 5, "beamLen": 20, "maxThreads": nd}, strict=False)
dt = time.perf_counter() - t0
t0 = time.perf_counter()
m.generateAnswers(dl, "val", {"beamSize": 5, "beamLen": 20, "maxThreads": max(1, nd // 2), "hostBeam": 1}, strict=False)
dt_host = (time.perf_counter() - t0) / max(1, nd // 2)
done = sum(1 for d in out for r in d["dialog"] if r is not None)
print(json.dumps({"encoder": enc, "dialogs": nd, "ms_per_dialog": dt / nd * 1e3, "ms_per_round": dt / nd / 10 * 1e3,
                  "rounds_with_a_finished_beam": done,
                  "ms_per_dialog_reference_structure": dt_host * 1e3, "beamSize": 5, "beamLen": 20, "vocabSize": 10000}))
dl.close(); m.engine.close()
