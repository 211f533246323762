"""encoders/hre-ques-hist.lua plugin: `model(params)` returns the encoder half of the engine.
The graph itself (reference encoders/hre-ques-hist.lua) runs in libvisdial_b200.so (csrc/engine.cu)."""
from ..modules import EncoderModule

NAME = "hre-ques-hist"


def model(params):
    assert params["encoder"] == NAME
    return EncoderModule(params, NAME, has_rnn_layers=True)
