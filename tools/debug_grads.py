"""GPU debug helper: per-segment gradient error vs the oracle at the headline layer sizes (B=2)."""
import sys, os
import numpy as np
sys.path.insert(0, os.path.join(os.path.dirname(__file__), "..")); sys.path.insert(0, os.path.join(os.path.dirname(__file__), "..", "tests"))
from helpers import full_params, seg_slices, torch_batch, torch_params
from oracle import philox, visdial_oracle as O
from visdial_b200 import Engine, Batch, VD_MATH_TF32, VD_MATH_FP32, init_parameters
from visdial_b200.synthetic import make_batch
enc, dec = (sys.argv[1], sys.argv[2]) if len(sys.argv) > 2 else ("mn-att-ques-im-hist", "disc")
p = full_params(enc, dec)
flat = init_parameters(p, seed=3)
nb = make_batch(p, 2, seed=5)
ref = O.forward_backward(O.Ctx(train=True, mask_fn=philox.make_mask_fn(11, 3), structure="batched"), p, torch_params(p, flat), torch_batch(nb))
for mode, name in ((VD_MATH_FP32, "fp32"), (VD_MATH_TF32, "tf32")):
    eng = Engine(p); eng.set_math_mode(mode); eng.set_parameters(flat); eng.set_training(1); eng.set_dropout_seed(11, 3); eng.zero_grad()
    loss = eng.forward_backward(Batch(nb)); g = eng.get_gradients()
    print(name, "loss", loss, "ref", ref["loss"])
    for n, s in seg_slices(p).items():
        r = ref["grads"][n].numpy().ravel().astype(np.float64); a = g[s].astype(np.float64)
        print("  %-28s max|ref| %.3e  max err %.3e  rel %.3e  rms rel %.3e" % (n, np.abs(r).max(), np.abs(a - r).max(), np.abs(a - r).max() / max(np.abs(r).max(), 1e-30), np.sqrt(((a - r) ** 2).mean()) / max(np.sqrt((r ** 2).mean()), 1e-30)))
    eng.close()
