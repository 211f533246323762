"""decoders/gen.lua plugin (reference decoders/gen.lua:3-68)."""
from ..modules import DecoderModule


def model(params, enc):
    return DecoderModule(params, "gen", enc)


def forwardConnect(enc, dec, encOut, seqLen):      # gen.lua:30-42
    dec._eng().forward_connect()


def backwardConnect(enc, dec):                     # gen.lua:45-60 -> gradient wrt encOut
    return dec._eng().backward_connect(dec._last_batch)


def decoderConnect(dec):                           # gen.lua:63-68
    """The reference copies every decoder layer's last output / cell into userPrevOutput / userPrevCell so that the next
    single-token forward continues the sequence.  Here the chaining is explicit: `vd_gen_decoder_step` takes the previous
    (h, c) of both layers as arguments and `Model.generateAnswers` hands it the state the previous step returned, so
    there is nothing left to copy."""
    return None
