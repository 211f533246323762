"""Oracle fixture of the BENCHED workload (BASELINE config 4 at B = 32 dialogs: mn-att-ques-im-hist + disc, pool5
14x14x512, 10 rounds, 100 options x 20 tokens, V = 10 000) — VERDICT r01 "weak" items 1-2: the tensor-core modes are
compared with the oracle at the size that is benched, not at B = 2.

The oracle runs in fp64 ("batched" structure: same maths as the reference structure, the 100 option passes stacked)
on the exact inputs bench.py uses for rank 0 / batch 0 (init seed 1234, batch seed 1234); dropout masks come from the
Philox twin (seed 11, iteration 1).  The full gradient is 13.8 M values (110 MB in fp64), so the fixture keeps

  * train: loss, per-segment {l2 norm, abs-max, sum} of the gradient and a strided sample (every 997th element)
  * eval : loss, the (320,100) score matrix, encOut, the (320,100) rank matrix, processRanks of the gt ranks

Re-run (about 10 minutes, ~40 GB of host RAM):  python tests/golden/make_golden_c4_b32.py
"""
import os
import sys
import time

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(HERE, "..", ".."))
sys.path.insert(0, os.path.join(HERE, ".."))

from helpers import full_params, torch_batch, torch_params, flat_from_named, seg_slices  # noqa: E402
from oracle import philox, visdial_oracle as O  # noqa: E402
from visdial_b200 import init_parameters  # noqa: E402
from visdial_b200.synthetic import make_batch  # noqa: E402

B, INIT_SEED, BATCH_SEED, DROP_SEED, DROP_ITER, STRIDE = 32, 1234, 1234, 11, 1, 997


def main():
    dtype = torch.float64 if os.environ.get("GOLDEN_F32") != "1" else torch.float32
    torch.set_num_threads(int(os.environ.get("GOLDEN_THREADS", "8")))
    p = full_params("mn-att-ques-im-hist", "disc")
    flat = init_parameters(p, seed=INIT_SEED)
    nb = make_batch(p, B, seed=BATCH_SEED)
    P = torch_params(p, flat, dtype=dtype)
    tb = torch_batch(nb)
    tb["img_feat"] = tb["img_feat"].to(dtype)
    t0 = time.time()
    ev = O.forward_backward(O.Ctx(train=False, structure="batched"), p, P, tb, only_forward=True)
    print("eval forward %.1f s, loss %.9f" % (time.time() - t0, ev["loss"]), flush=True)
    scores = ev["decOut"]
    ranks = O.compute_ranks(scores).numpy().astype(np.int32)
    gt = nb["answer_ind"].reshape(-1).astype(np.int64) - 1
    gt_ranks = ranks[np.arange(ranks.shape[0]), gt]
    metrics = O.process_ranks(torch.from_numpy(gt_ranks))
    t0 = time.time()
    tr = O.forward_backward(O.Ctx(train=True, mask_fn=philox.make_mask_fn(DROP_SEED, DROP_ITER), structure="batched"), p, P, tb)
    print("train forward+backward %.1f s, loss %.9f" % (time.time() - t0, tr["loss"]), flush=True)
    g = flat_from_named_f64(p, tr["grads"])
    names, norms, amax, sums = [], [], [], []
    for name, sl in seg_slices(p).items():
        names.append(name)
        norms.append(float(np.sqrt((g[sl] ** 2).sum())))
        amax.append(float(np.abs(g[sl]).max()))
        sums.append(float(g[sl].sum()))
    out = dict(B=np.int32(B), init_seed=np.int32(INIT_SEED), batch_seed=np.int32(BATCH_SEED), drop_seed=np.int32(DROP_SEED),
               drop_iter=np.int32(DROP_ITER), stride=np.int32(STRIDE),
               eval_loss=np.float64(ev["loss"]), eval_scores=scores.numpy().astype(np.float64),
               eval_encOut=ev["encOut"].numpy().astype(np.float32), eval_ranks=ranks, gt_ranks=gt_ranks.astype(np.int32),
               metrics=np.array([metrics[k] for k in ("r1", "r5", "r10", "medianR", "meanR", "meanRR")], dtype=np.float64),
               train_loss=np.float64(tr["loss"]), train_scores=tr["decOut"].numpy().astype(np.float64),
               grad_sample=g[::STRIDE].copy(), seg_names=np.array(names), seg_norm=np.array(norms), seg_absmax=np.array(amax),
               seg_sum=np.array(sums))
    name = os.path.join(HERE, "c4_b32__mn-att-ques-im-hist__disc.npz")
    np.savez_compressed(name, **out)
    print(name, os.path.getsize(name), "bytes")


def flat_from_named_f64(p, named):
    from visdial_b200 import engine as E
    segs, n = E.layout(p)
    flat = np.zeros(n, dtype=np.float64)
    for s in segs:
        flat[s.offset:s.offset + s.size] = named[s.name].detach().numpy().astype(np.float64).reshape(-1)
    return flat


if __name__ == "__main__":
    main()
