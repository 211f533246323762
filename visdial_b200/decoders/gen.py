"""decoders/gen.lua plugin (reference decoders/gen.lua:3-68)."""
from ..modules import DecoderModule


def model(params, enc):
    return DecoderModule(params, "gen", enc)


def forwardConnect(enc, dec, encOut, seqLen):      # gen.lua:30-42
    dec._eng().forward_connect()


def backwardConnect(enc, dec):                     # gen.lua:45-60 -> gradient wrt encOut
    return dec._eng().backward_connect(dec._last_batch)


def decoderConnect(dec):                           # gen.lua:63-68 — only used by beam search (out of scope)
    raise NotImplementedError("decoderConnect is used by generate.lua's beam search, outside the hot path")
