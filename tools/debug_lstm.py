"""GPU debug helper: one TF32 forward of the mid-size config (run under compute-sanitizer)."""
import sys, os
sys.path.insert(0, os.path.join(os.path.dirname(__file__), "..")); sys.path.insert(0, os.path.join(os.path.dirname(__file__), "..", "tests"))
from helpers import small_params
from visdial_b200 import Engine, Batch, VD_MATH_TF32, init_parameters
from visdial_b200.synthetic import make_batch
enc, dec = sys.argv[1], sys.argv[2]
p = small_params(enc, dec, rnnHiddenSize=128, embedSize=64, vocabSize=200, numOptions=10, commonEmbeddingSize=64,
                 imgFeatureSize=64 if "att" in enc else 256, imgSpatialSize=4, imgEmbedSize=32)
nb = make_batch(p, 13, seed=7, max_ques_len=9, max_ans_len=6, max_cap_len=12, max_hist_len=14, max_hist_concat=40, empty_round_every=4)
eng = Engine(p); eng.set_math_mode(VD_MATH_TF32); eng.set_parameters(init_parameters(p, seed=3)); eng.set_training(1); eng.zero_grad()
only_fwd = len(sys.argv) > 3
print("loss", eng.forward_backward(Batch(nb), only_forward=only_fwd))
