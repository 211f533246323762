"""decoders/disc.lua plugin (reference decoders/disc.lua:3-38)."""
from ..modules import DecoderModule


def model(params, enc):
    return DecoderModule(params, "disc", enc)


def forwardConnect(enc, dec, encOut, seqLen):      # disc.lua:35 — dummy
    pass


def backwardConnect(enc, dec):                     # disc.lua:38 — dummy
    pass
