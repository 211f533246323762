"""Checkpoints in the reference's schema and container.

    train.lua:99-102   torch.save('model_epoch_%d.t7', {modelW = model.wrapperW, optims = model.optims, modelParams = modelParams})
    train.lua:120-121  torch.save('model_final.t7',    {modelW = model.wrapperW:float(), modelParams = modelParams})
    train.lua:33-34,78-80 / evaluate.lua:58-91 / generate.lua:53-83
                       savedModel = torch.load(path); Model(savedModel.modelParams); wrapperW:copy(savedModel.modelW);
                       optims.learningRate = savedModel.optims.learningRate

`optims` is the table `adam` keeps its state in (model_utils/optim_updates.lua:62-91): learningRate, t, m, v, tmp.

WEIGHT ORDER.  `modelW` is the engine's flat parameter vector in the engine's documented segment order (DESIGN.md §3,
`vd_layout_segment`), and the checkpoint carries that order explicitly as `layout = {{name, offset, rows, cols}, ...}`
(an extra key the reference ignores).  The reference's `wrapper:getParameters()` flattens in nngraph's module
traversal order, which depends on the un-vendored nn/nngraph versions and cannot be derived without running Torch7;
a checkpoint trained by the reference therefore loads here only through a `permutation` (see `load_checkpoint`) that
a maintainer produces once on a Torch7 box by dumping the sizes of `wrapper:parameters()`.  Files written here load
here, and load in the reference as a table of the right shape.
"""
from __future__ import annotations

from typing import Dict, Optional

import numpy as np

from . import t7
from .engine import layout as _layout


def _plain(v):
    if isinstance(v, (np.integer,)):
        return int(v)
    if isinstance(v, (np.floating,)):
        return float(v)
    return v


def checkpoint_table(model, final: bool = False) -> Dict:
    """The Lua table train.lua writes, from a visdial_b200.Model."""
    eng = model.engine
    w = eng.get_parameters()
    segs, n = _layout(model.params)
    out = {
        "modelW": w if final else t7.CudaTensor(w),                                  # train.lua:120 / :100
        "modelParams": {k: _plain(v) for k, v in model.params.items()
                        if isinstance(v, (str, int, float, bool, np.integer, np.floating))},
        "layout": [{"name": s.name, "offset": int(s.offset), "rows": int(s.rows), "cols": int(s.cols)} for s in segs],
    }
    if not final:                                                                    # train.lua:101
        m, v, t = eng.optim_state()
        out["optims"] = {"learningRate": float(model.optims["learningRate"]), "t": int(t),
                         "m": t7.CudaTensor(m), "v": t7.CudaTensor(v)}
    return out


def save_checkpoint(model, path: str, final: bool = False):
    t7.save(path, checkpoint_table(model, final))


def load_checkpoint(path: str, permutation: Optional[np.ndarray] = None) -> Dict:
    """torch.load(path) with modelW as a float32 vector.  `permutation[i]` = index in the FILE's modelW of the engine's
    parameter i (for checkpoints whose flattening order is not the engine's)."""
    ck = t7.load(path)
    if not isinstance(ck, dict) or "modelW" not in ck or "modelParams" not in ck:
        raise t7.T7Error("%s is not a visdial checkpoint (modelW / modelParams missing)" % path)
    w = np.ascontiguousarray(np.asarray(ck["modelW"]).reshape(-1), dtype=np.float32)
    if permutation is not None:
        w = w[np.asarray(permutation, dtype=np.int64)]
    ck["modelW"] = w
    return ck


def restore(model, ck: Dict, restore_adam_state: bool = False):
    """train.lua:78-80: wrapperW:copy(savedModel.modelW); optims.learningRate = savedModel.optims.learningRate.
    `restore_adam_state` additionally restores m, v, t (the reference does not: its Adam restarts from zero moments)."""
    eng = model.engine
    w = ck["modelW"]
    if w.size != eng.num_params:
        raise t7.T7Error("checkpoint has %d parameters, the model %d (different encoder/decoder or vocabulary?)"
                         % (w.size, eng.num_params))
    eng.set_parameters(w)
    opt = ck.get("optims")
    if opt and "learningRate" in opt:
        model.optims["learningRate"] = float(opt["learningRate"])
    if restore_adam_state and opt and "m" in opt and "v" in opt:
        m = np.ascontiguousarray(np.asarray(opt["m"]).reshape(-1), dtype=np.float32)
        v = np.ascontiguousarray(np.asarray(opt["v"]).reshape(-1), dtype=np.float32)
        eng.set_optim_state(m, v, int(opt.get("t", 0)))
