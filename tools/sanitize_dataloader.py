"""Small driver for compute-sanitizer: device dataloader (corpus prepare + batch gather) feeding two training
iterations of mn-att-ques-im-hist + disc with the option stream overlapped (mid-size layers, TF32 path)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import numpy as np
from helpers import small_params
from visdial_b200 import Model
from visdial_b200.dataloader import Dataloader
from visdial_b200.synthetic import make_corpus

enc = sys.argv[1] if len(sys.argv) > 1 else "mn-att-ques-im-hist"
dec = sys.argv[2] if len(sys.argv) > 2 else "disc"
p = small_params(enc, dec, vocabSize=60, embedSize=32, rnnHiddenSize=128, imgEmbedSize=32, commonEmbeddingSize=64,
                 numOptions=20, batchSize=6)
concat = "lf" in enc and "hist" in enc
raw = make_corpus(p, 24, 120, seed=1, ques_len_cap=14 if concat else None, ans_len_cap=13 if concat else None)
m = Model(p, seed=2)
opt = dict(p, useHistory="hist" in enc, concatHistory=concat, useIm="im" in enc, maxHistoryLen=60, imgNorm=1)
dl = Dataloader(m.engine, seed=3).initialize(opt, ["train", "val"], {"train": raw, "val": raw})
train = "--no-train" not in sys.argv
if train:
    for _ in range(2):
        print("loss", m.trainIteration(dl))
b, nxt = dl.getTestBatch(0, p, "val")
if train:
    print("ranks", m.retrieveBatch(b)[:10])
else:                                   # dataloader kernels only: every batch form, read back
    for mode in (0, 1, 2):
        d = dl.corpus["train"].get_batch(np.array([0, 5, 5, 23, 7]), mode)
        print(mode, {k: v.shape for k, v in d.numpy().items()})
    print({k: dl.corpus["val"].read(k).shape for k in ("ques_fwd", "hist", "ans_in", "opt_out", "img_fv")})
m.engine.synchronize()
dl.close()
m.engine.close()
print("done")
