"""Data-parallel equivalence: the same global batch on 1 GPU and sharded by dialog over 2 GPUs
(one process per GPU, one NCCL all-reduce of the flat gradient inside vd_clamp_adam_step) must give the same
gradient and the same weights after the optimiser step, to fp32 reduction-order tolerance."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.multiprocessing as mp

from helpers import small_batch, small_params

pytestmark = pytest.mark.gpu


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    return port


def _worker(rank, world, port, enc, dec, out):
    import torch.distributed as dist
    from visdial_b200 import VD_MATH_FP32, Batch, Engine, init_parameters
    from visdial_b200 import dist as vdist
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    p = small_params(enc, dec, gpuid=rank)
    full = small_batch(p, B=4, seed=3)
    mine = vdist.shard_batch(full, rank, world, p["maxQuesCount"])
    eng = Engine(p)
    eng.set_math_mode(VD_MATH_FP32)
    eng.set_parameters(init_parameters(p, seed=3))
    eng.set_training(2)                      # training graph, dropout off: masks are indexed per local shard
    vdist.attach_engine(eng, rank, world)
    eng.zero_grad()
    loss = eng.forward_backward(Batch(mine))
    eng.clamp_adam_step(1e-3)
    w = eng.get_parameters()
    g = eng.get_gradients()                  # all-reduced, scaled, clamped
    gathered = vdist.gather_ranks(np.asarray([loss], dtype=np.float64), rank, world)
    if rank == 0:
        out.put((w, g, gathered))
    eng.close()
    dist.destroy_process_group()


@pytest.mark.parametrize("enc,dec", [("mn-att-ques-im-hist", "disc"), ("lf-ques", "gen")])
def test_two_gpus_match_one(enc, dec):
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs (run with gpurun --gpus 2)")
    from visdial_b200 import VD_MATH_FP32, Batch, Engine, init_parameters
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, enc, dec, q)) for r in range(2)]
    for pr in procs:
        pr.start()
    w2, g2, losses = q.get(timeout=300)
    for pr in procs:
        pr.join(timeout=60)
        assert pr.exitcode == 0
    p = small_params(enc, dec, gpuid=0)
    full = small_batch(p, B=4, seed=3)
    eng = Engine(p)
    eng.set_math_mode(VD_MATH_FP32)
    eng.set_parameters(init_parameters(p, seed=3))
    eng.set_training(2)
    eng.zero_grad()
    loss1 = eng.forward_backward(Batch(full))
    eng.clamp_adam_step(1e-3)
    w1, g1 = eng.get_parameters(), eng.get_gradients()
    eng.close()
    if dec == "disc":                        # mean criterion: global loss = mean of the equal-size shard losses
        assert float(np.mean(losses)) == pytest.approx(loss1, rel=1e-5)
    else:                                    # sum criterion: global loss = sum of the shard losses
        assert float(np.sum(losses)) == pytest.approx(loss1, rel=1e-5)
    scale = max(float(np.abs(g1).max()), 1e-30)
    assert float(np.abs(g2 - g1).max()) < 1e-5 * scale + 1e-7
    assert float(np.abs(w2 - w1).max()) < 0.02 * 1e-3            # Adam steps are ~lr: compare in units of lr
