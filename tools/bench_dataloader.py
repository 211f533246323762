"""Device batch assembly (visdial_b200.dataloader) alone: device time per batch and achieved HBM bandwidth over a
range of batch sizes, next to the oracle's numpy indexing on the host.  Diagnostic; bench.py carries the B=32 figure."""
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from bench import headline_params, measured_peaks  # noqa: E402
from visdial_b200.dataloader import Dataloader  # noqa: E402
from visdial_b200.engine import Engine  # noqa: E402
from visdial_b200.synthetic import make_corpus  # noqa: E402

p = headline_params(0)
eng = Engine(p)
raw = make_corpus(p, num_threads=512, num_opt_list=8000, seed=3)
dl = Dataloader(eng, seed=1).initialize(dict(p, imgNorm=0, maxHistoryLen=60), ["train"], {"train": raw})
peak = measured_peaks()["hbm"]
for B in (32, 64, 128, 256, 512):
    for _ in range(4):
        dl.getTrainBatch(p, B)
    eng.synchronize()
    eng.profile_reset()
    eng.profile(1)
    nb, by = 30, 0
    eng.timer_start()
    for _ in range(nb):
        dl.getTrainBatch(p, B)
        by += dl.corpus["train"].batch_bytes()[0]
    ms_stream = eng.timer_stop()
    st = eng.kernel_stats("corpus_gather")
    eng.profile(False)
    gbs = by / (st["ms"] * 1e-3) / 1e9
    print("B=%4d  %.1f MB/batch  device %.1f us/batch (events around the 2 launches)  stream %.1f us/batch  %.0f GB/s = %.2f of HBM peak"
          % (B, by / nb / 1e6, st["ms"] / nb * 1e3, ms_stream / nb * 1e3, gbs, gbs / peak), flush=True)
dl.close()
eng.close()
