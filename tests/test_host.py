"""Host-side logic that needs no GPU: the synthetic dataloader contract, the plugin loaders, and the
data-parallel plumbing (world_size-2 gloo: dialog sharding + unique-id broadcast + result gather)."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from helpers import ALL_ENCODERS, full_params, small_batch, small_params
from visdial_b200 import decoders, encoders
from visdial_b200 import dist as vdist
from visdial_b200.synthetic import make_batch
from visdial_b200.utils import processRanks


def test_batch_contract_small():
    p = small_params("mn-att-ques-im-hist", "disc")
    b = small_batch(p, B=4)
    B, R, V = 4, 10, p["vocabSize"]
    q = b["ques_fwd"]
    assert q.shape[:2] == (B, R) and q.dtype == np.int32
    # right aligned: once a token appears, no pad follows
    for row in q.reshape(-1, q.shape[2]):
        nz = np.nonzero(row)[0]
        assert len(nz) == 0 or (nz[-1] == len(row) - 1 and np.all(row[nz[0]:] != 0))
    assert (q.reshape(B * R, -1) != 0).sum(1).min() == 0               # an all-pad round exists
    assert q.max() <= V - 2
    h = b["hist"]
    assert h.shape[:2] == (B, R) and h.shape[2] <= 9
    assert np.all(h[:, :, -1] != 0)                                     # every round has some history
    ai, ao = b["answer_in"].reshape(B * R, -1), b["answer_out"].reshape(B * R, -1)
    assert np.all(ai[:, 0] == V - 1)                                    # <START>
    assert np.array_equal(ai != 0, ao != 0)                             # aligned lengths
    L = (ao != 0).sum(1)
    assert np.all(ao[np.arange(B * R), L - 1] == V)                     # <END> closes every answer
    o = b["options"]
    assert o.shape == (B * R, p["numOptions"], 20)
    for row in o.reshape(-1, 20)[:200]:                                  # left aligned
        nz = np.nonzero(row)[0]
        assert len(nz) > 0 and nz[0] == 0 and np.all(row[:nz[-1] + 1] != 0)
    gt = b["answer_ind"]
    assert gt.min() >= 1 and gt.max() <= p["numOptions"]
    assert b["img_feat"].shape == (B, 3, 3, 8)


def test_batch_contract_headline_shapes():
    p = full_params("mn-att-ques-im-hist", "disc")
    b = make_batch(p, 2, seed=1)
    assert b["ques_fwd"].shape == (2, 10, 20) and b["hist"].shape == (2, 10, 40)
    assert b["options"].shape == (20, 100, 20) and b["img_feat"].shape == (2, 14, 14, 512)
    assert b["answer_in"].shape == (2, 10, 20)
    p2 = full_params("lf-ques-im-hist", "disc")
    b2 = make_batch(p2, 2, seed=1)
    assert b2["img_feat"].shape == (2, 4096)
    assert np.allclose(np.linalg.norm(b2["img_feat"], axis=1), 1, atol=1e-5)      # L2-normalised fc7
    assert 40 < b2["hist"].shape[2] <= 300                                        # concatenated history
    g = make_batch(full_params("lf-ques", "gen"), 1, seed=1, gen_eval=True)
    assert g["option_in"].shape == g["option_out"].shape and g["option_in"].shape[:3] == (1, 10, 100)
    assert np.all(g["option_in"][..., 0] == 9999)


def test_plugin_loaders():
    for name in ALL_ENCODERS:
        m = encoders.load(name)
        enc = m.model(small_params(name, "disc"))
        # the gModule encoders (mn-*, lf-att-*) do not export rnnLayers (gen.lua:30 then skips decoderConnect's copy)
        assert (enc.rnnLayers is None) == (name.startswith("mn") or name.startswith("lf-att"))
    with pytest.raises(Exception):
        encoders.load("lf-ques-im-hist-nope")
    for name in ("disc", "gen"):
        d = decoders.load(name)
        assert callable(d.model) and callable(d.forwardConnect) and callable(d.backwardConnect)
    with pytest.raises(ValueError):
        decoders.load("nope")


def test_process_ranks():
    m = processRanks(np.array([[1, 2, 10], [100, 5, 1]]), verbose=False)
    assert m["r@1"] == pytest.approx(2 / 6) and m["r@5"] == pytest.approx(4 / 6) and m["r@10"] == pytest.approx(5 / 6)
    assert m["meanR"] == pytest.approx(119 / 6) and m["medianR"] == pytest.approx(3.5)


def test_shard_dialogs():
    assert vdist.shard_range(32, 0, 1) == (0, 32)
    spans = [vdist.shard_range(35, r, 4) for r in range(4)]
    assert spans[0][0] == 0 and spans[-1][1] == 35
    assert all(spans[i][1] == spans[i + 1][0] for i in range(3))       # contiguous, never splits a dialog
    assert max(b - a for a, b in spans) - min(b - a for a, b in spans) <= 1
    p = small_params("mn-att-ques-im-hist", "disc")
    b = small_batch(p, B=5)
    parts = [vdist.shard_batch(b, r, 2, p["maxQuesCount"]) for r in range(2)]
    assert np.array_equal(np.concatenate([x["options"] for x in parts]), b["options"])
    assert np.array_equal(np.concatenate([x["ques_fwd"] for x in parts]), b["ques_fwd"])
    assert np.array_equal(np.concatenate([x["answer_ind"] for x in parts]), b["answer_ind"])


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    return port


def _worker(rank, world, port, out):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    uid = vdist.broadcast_unique_id(lambda: bytes(range(128)) if rank == 0 else None, rank)
    p = small_params("mn-att-ques-im-hist", "disc")
    b = small_batch(p, B=5)
    mine = vdist.shard_batch(b, rank, world, p["maxQuesCount"])
    ranks_local = mine["answer_ind"].astype(np.int32)                   # stand-in for per-rank rank output
    gathered = vdist.gather_ranks(ranks_local, rank, world)
    t = vdist.max_over_ranks(float(rank + 1))
    if rank == 0:
        out.put((uid, gathered.tolist(), b["answer_ind"].tolist(), t))
    dist.destroy_process_group()


def test_gloo_world2_plumbing():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for pr in procs:
        pr.start()
    uid, gathered, expect, t = q.get(timeout=120)
    for pr in procs:
        pr.join(timeout=60)
        assert pr.exitcode == 0
    assert uid == bytes(range(128))          # rank 0's id reached every rank
    assert gathered == expect                # rank-ordered concatenation == unsharded order
    assert t == 2.0                          # timing is the max over ranks


def test_eval_split_partition_walks_every_dialog_once_in_order():
    """Dataloader(rank, world): contiguous per-rank ranges of an evaluation split, walked in batches; the rank-ordered
    concatenation is 0..n-1 (what dist.gather_ranks relies on), for ragged sizes too."""
    from visdial_b200.dataloader import eval_partition, test_batch_indices
    for n, world, bs in ((2064, 8, 32), (17, 4, 5), (3, 4, 2), (40, 1, 16)):
        seen = []
        for rank in range(world):
            lo, hi = eval_partition(n, rank, world)
            start = 0
            while start < hi - lo:
                inds, nxt = test_batch_indices(start, bs, lo, hi)
                assert 0 < len(inds) <= bs and nxt > start
                seen += inds.tolist()
                start = nxt
        assert seen == list(range(n)), (n, world, bs)


def test_bench_reference_arm_prints_the_contract_line():
    """`bench.py --impl reference` needs no GPU: one JSON line with the contract's keys (a bounded CPU sample)."""
    import json
    import subprocess
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    out = subprocess.run([sys.executable, os.path.join(root, "bench.py"), "--impl", "reference", "--steps", "1", "--warmup", "0",
                          "--ref-batch", "1", "--cpu-threads", "4"], capture_output=True, text=True, timeout=300, cwd=root)
    assert out.returncode == 0, out.stderr[-2000:]
    line = json.loads(out.stdout.strip().splitlines()[-1])
    for k in ("impl", "metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling",
              "vs_baseline", "dtype", "data", "config", "cpu_baseline", "e2e"):
        assert k in line, k
    assert line["impl"] == "reference" and line["unit"] == "QA-rounds/s" and line["value"] > 0
    assert line["cpu_baseline"]["kind"] == "port" and line["cpu_baseline"]["cores"] == 4
    assert line["e2e"]["h2d_bytes_per_step"] == 0 and line["e2e"]["value"] == line["value"]
    assert line["native_so_loaded"] == []                 # the reference process never loads the product library
    assert line["cpu_baseline_batched"]["value"] > 0
    assert line["config"]["dialogs_per_gpu"] == 1


def test_bench_reference_arm_picks_threads_and_runs_other_configs():
    import json
    import subprocess
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    out = subprocess.run([sys.executable, os.path.join(root, "bench.py"), "--impl", "reference", "--config", "C1", "--steps", "1",
                          "--warmup", "1", "--no-batched"], capture_output=True, text=True, timeout=600, cwd=root)
    assert out.returncode == 0, out.stderr[-2000:]
    line = json.loads(out.stdout.strip().splitlines()[-1])
    assert line["metric"].startswith("QA-rounds/sec lf-ques+gen") and line["value"] > 0
    assert line["config"]["dialogs_per_gpu"] == 4 and line["warmup"] == 1
    assert len(line["thread_sweep_qa_rounds_per_s"]) >= 1 and line["cpu_baseline"]["cores"] >= 1


def test_oracle_layout_twin_matches_the_engine():
    """oracle/layout.py (numpy-only, used by the reference arm) == vd_layout_* / init_parameters of the product."""
    import bench
    from oracle import layout as OL
    from visdial_b200 import engine as E
    assert {k: bench.DEFAULTS[k] for k in E.DEFAULT_PARAMS} == E.DEFAULT_PARAMS
    for enc, dec in [(e, d) for e in ALL_ENCODERS for d in ("disc", "gen")]:
        p = small_params(enc, dec, numAttentionLayers=2 if "att" in enc else 1)
        segs, n = E.layout(p)
        osegs, on = OL.layout(p)
        assert n == on and len(segs) == len(osegs)
        for a, b in zip(segs, osegs):
            assert (a.name, a.offset, a.rows, a.cols, a.init, a.fan_in) == (b.name, b.offset, b.rows, b.cols, b.init, b.fan_in)
        assert np.array_equal(E.init_parameters(p, seed=5), OL.init_parameters(p, seed=5))
    for name in bench.CONFIGS:
        p = bench.config_params(name)
        q = E.derive_flags(dict(p))
        assert p == q
