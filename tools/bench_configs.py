"""GPU timing of the other BASELINE configs (train step and eval/retrieve step), device-timed, batch resident in
HBM.  Not the graded bench line (bench.py is); fills BASELINE.md §4 and the C5 ranker sweep."""
import argparse, json, os, sys, time
sys.path.insert(0, os.path.join(os.path.dirname(__file__), ".."))
sys.path.insert(0, os.path.join(os.path.dirname(__file__), "..", "tests"))
import numpy as np
from helpers import full_params
from visdial_b200 import Batch, Model
from visdial_b200.synthetic import make_batch

ap = argparse.ArgumentParser()
ap.add_argument("--sweep", action="store_true", help="C5: disc ranker forward-only sweep over B")
args = ap.parse_args()


def time_steps(eng, fn, steps=5, warm=2):
    for _ in range(warm):
        fn()
    eng.synchronize()
    eng.timer_start()
    for _ in range(steps):
        fn()
    return eng.timer_stop() / steps


def run(enc, dec, B, V, train=True, hist_concat=300):
    p = full_params(enc, dec, vocabSize=V, batchSize=B)
    m = Model(p, seed=1)
    eng = m.engine
    nb = make_batch(p, B, seed=3, max_hist_concat=hist_concat, gen_eval=False)
    b = Batch(nb).to_device(eng)
    out = {"config": "%s + %s" % (enc, dec), "B": B, "V": V, "Th": int(nb["hist"].shape[2]) if "hist" in nb else 0}

    class DL:
        def getTrainBatch(self, params):
            return b
    if train:
        ms = time_steps(eng, lambda: m.trainIteration(DL()))
        out.update(train_ms=round(ms, 3), train_rounds_per_s=round(B * 10 / (ms * 1e-3), 1))
    if dec == "disc":
        ms = time_steps(eng, lambda: eng.retrieve(b, use_gt=True))
        out.update(eval_ms=round(ms, 3), eval_rounds_per_s=round(B * 10 / (ms * 1e-3), 1))
    eng.close()
    return out


if args.sweep:
    for B in (32, 64, 128, 256, 512, 1024):
        print(json.dumps(run("mn-att-ques-im-hist", "disc", B, 10000, train=False)), flush=True)
else:
    print(json.dumps(run("lf-ques", "gen", 4, 1000)), flush=True)
    print(json.dumps(run("lf-ques-im-hist", "disc", 32, 10000)), flush=True)
    print(json.dumps(run("hrea-ques-im-hist", "gen", 32, 10000)), flush=True)
    print(json.dumps(run("mn-att-ques-im-hist", "disc", 32, 10000)), flush=True)
    print(json.dumps(run("mn-att-ques-im-hist", "gen", 32, 10000)), flush=True)
