"""encoders/lf-ques-im.lua plugin: `model(params)` returns the encoder half of the engine.
The graph itself (reference encoders/lf-ques-im.lua) runs in libvisdial_b200.so (csrc/engine.cu)."""
from ..modules import EncoderModule

NAME = "lf-ques-im"


def model(params):
    assert params["encoder"] == NAME
    return EncoderModule(params, NAME, has_rnn_layers=True)
