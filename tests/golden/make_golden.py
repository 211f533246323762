"""Generates tests/golden/*.npz from the ORACLE (there is no runnable reference: the Lua/Torch7 stack is absent and
the reference ships no vectors — DESIGN.md §2).  The fixtures pin the oracle against drift between rounds and give the
GPU tests committed inputs/outputs that travel to the GPU box.  Re-run:  python tests/golden/make_golden.py"""
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(HERE, "..", ".."))
sys.path.insert(0, os.path.join(HERE, ".."))

from helpers import small_batch, small_params, torch_batch, torch_params, flat_from_named  # noqa: E402
from oracle import philox, visdial_oracle as O  # noqa: E402
from visdial_b200 import init_parameters  # noqa: E402

CASES = [("mn-att-ques-im-hist", "disc"), ("lf-ques", "gen"), ("hrea-ques-im-hist", "gen"), ("lf-ques-im-hist", "disc"),
         # the seven sub-graph encoders, each with one decoder
         ("lf-ques-im", "disc"), ("lf-ques-hist", "gen"), ("hre-ques-hist", "disc"), ("hre-ques-im-hist", "gen"),
         ("mn-ques-hist", "gen"), ("mn-ques-im-hist", "disc"), ("lf-att-ques-im-hist", "disc")]


def main():
    torch.set_num_threads(1)
    for enc, dec in CASES:
        name = os.path.join(HERE, "%s__%s.npz" % (enc, dec))
        if os.path.exists(name) and "--all" not in sys.argv:      # committed fixtures stay byte-identical; --all regenerates every one
            continue
        p = small_params(enc, dec)
        flat = init_parameters(p, seed=21)
        nb = small_batch(p, B=2, seed=13, gen_eval=False)
        P64 = torch_params(p, flat, dtype=torch.float64)
        tb = torch_batch(nb)
        if "img_feat" in tb:
            tb["img_feat"] = tb["img_feat"].double()
        psite = {O.SITE_FUSION: p["dropout"]}
        ev = O.forward_backward(O.Ctx(train=False), p, P64, tb, only_forward=True)
        tr = O.forward_backward(O.Ctx(train=True, mask_fn=philox.make_mask_fn(5, 2, psite)), p, P64, tb)
        out = {"flat": flat, "eval_loss": np.float64(ev["loss"]), "eval_encOut": ev["encOut"].numpy(),
               "eval_decOut": ev["decOut"].numpy(), "train_loss": np.float64(tr["loss"]),
               "train_grad": flat_from_named(p, tr["grads"]).astype(np.float64)}
        for k, v in nb.items():
            out["batch_" + k] = v
        if dec == "disc":
            out["ranks"] = O.compute_ranks(ev["decOut"]).numpy().astype(np.int32)
        np.savez_compressed(name, **out)
        print(name, os.path.getsize(name), "bytes")


if __name__ == "__main__":
    main()
