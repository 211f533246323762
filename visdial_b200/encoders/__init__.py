"""Encoder plugins, loaded by name like `dofile('encoders/<name>.lua')` (model.lua:19-20).

One table instead of one file per name: every graph of encoders/*.lua runs in libvisdial_b200.so (csrc/engine.cu composes it from the
encoder name), so the only thing a plugin file would carry is whether the upstream encoder exports `enc.rnnLayers` — the Sequential
encoders do (e.g. lf-ques.lua:17-25, hre-ques-hist.lua:20-29), the nngraph gModule ones do not (mn-*.lua, lf-att-ques-im-hist.lua), and
decoders/gen.lua:30-42 copies the last encoder state into the decoder only when it is there."""
from types import SimpleNamespace

from ..modules import EncoderModule

EXPORTS_RNN_LAYERS = {
    "lf-ques": True, "lf-ques-im": True, "lf-ques-hist": True, "lf-ques-im-hist": True, "lf-att-ques-im-hist": False,
    "hre-ques-hist": True, "hre-ques-im-hist": True, "hrea-ques-im-hist": True,
    "mn-ques-hist": False, "mn-ques-im-hist": False, "mn-att-ques-im-hist": False,
}


def load(name: str):
    """-> an object with `.model(params)`, the shape `dofile('encoders/<name>.lua')` returns."""
    if name not in EXPORTS_RNN_LAYERS:
        raise ValueError("unknown encoder '%s' (no encoders/%s.lua)" % (name, name))

    def model(params):
        if params["encoder"] != name:
            raise ValueError("params.encoder is '%s', plugin is '%s'" % (params["encoder"], name))
        return EncoderModule(params, name, has_rnn_layers=EXPORTS_RNN_LAYERS[name])
    return SimpleNamespace(NAME=name, model=model)
