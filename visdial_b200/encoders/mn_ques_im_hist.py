"""encoders/mn-ques-im-hist.lua plugin: `model(params)` returns the encoder half of the engine.
The graph itself (reference encoders/mn-ques-im-hist.lua) runs in libvisdial_b200.so (csrc/engine.cu)."""
from ..modules import EncoderModule

NAME = "mn-ques-im-hist"


def model(params):
    assert params["encoder"] == NAME
    return EncoderModule(params, NAME, has_rnn_layers=False)
