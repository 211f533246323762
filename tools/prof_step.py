"""A few C4 training steps and nothing else (for `ncu -k regex:... -s N -c M python tools/prof_step.py`).
usage: python tools/prof_step.py [--math f16|tf32] [--steps 2] [--batch 32] [--config C4]"""
import argparse
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import bench  # noqa: E402
from visdial_b200 import VD_MATH_F16, VD_MATH_FP32, VD_MATH_TF32, Batch, Model  # noqa: E402
from visdial_b200.synthetic import make_batch  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--math", default="f16")
ap.add_argument("--steps", type=int, default=2)
ap.add_argument("--batch", type=int, default=32)
ap.add_argument("--config", default="C4")
ap.add_argument("--overlap", type=int, default=1)
ap.add_argument("--warmup", type=int, default=0)
a = ap.parse_args()
p = bench.config_params(a.config)
p["batchSize"] = a.batch
m = Model(p, seed=1234)
m.engine.set_math_mode({"f16": VD_MATH_F16, "tf32": VD_MATH_TF32, "fp32": VD_MATH_FP32}[a.math])
m.engine.set_option_overlap(bool(a.overlap))
b = Batch(make_batch(p, a.batch, seed=1234)).to_device(m.engine)


class L:
    def getTrainBatch(self, params):
        return b


for _ in range(a.warmup):
    m.trainIteration(L())
m.engine.synchronize()
m.engine.timer_start()
for _ in range(a.steps):
    m.trainIteration(L())
print("ms/step", m.engine.timer_stop() / a.steps)
