"""nn.Module-protocol objects returned by the encoder / decoder plugins: they obey
:forward / :backward / :training / :evaluate / :getParameters like the nn graphs the reference's
plugins return, but every call lands in libvisdial_b200.so."""
from __future__ import annotations

from .engine import Batch, DeviceTensor, Engine, init_parameters


class _Half:
    def __init__(self, params, name):
        self.params, self.name = params, name
        self.engine: Engine = None

    def _eng(self) -> Engine:
        if self.engine is None:
            raise RuntimeError("module is not inside a wrapper yet (model.lua:42 nn.Sequential():add(enc):add(dec))")
        return self.engine


class EncoderModule(_Half):
    """Returned by encoders/<name>.model(params).  Fields the decoders read: .wordEmbed (disc.lua:12,
    gen.lua:10) is implicit — one engine holds the shared table; .rnnLayers (gen.lua:30-42)."""

    def __init__(self, params, name, has_rnn_layers):
        super().__init__(params, name)
        self.wordEmbed = "wordEmbed.weight"
        self.rnnLayers = ["ques.lstm1", "ques.lstm2"] if has_rnn_layers else None
        self.output = None

    def forward(self, batch: Batch) -> DeviceTensor:          # encoder:forward(inputs), model.lua:297
        self.output = self._eng().encoder_forward(batch)
        return self.output

    def backward(self, batch: Batch, gradOutput: DeviceTensor):   # encoder:backward, model.lua:323,337
        self._eng().encoder_backward(batch, gradOutput)


class DecoderModule(_Half):
    def __init__(self, params, name, enc: EncoderModule):
        super().__init__(params, name)
        self.enc = enc
        self.rnnLayers = ["dec.lstm1", "dec.lstm2"] if name == "gen" else None
        self.output = None

    def forward(self, batch: Batch) -> DeviceTensor:          # decoder:forward, model.lua:313,329
        self.output = self._eng().decoder_forward(batch)
        return self.output

    def backward(self, batch: Batch, gradOutput=None):        # decoder:backward, model.lua:319,335
        self._eng().decoder_backward(batch)
        # disc returns {gradOptions, gradEncOut}; the caller uses [2] (model.lua:337)
        return [None, self._eng().backward_connect(batch)] if self.name == "disc" else None


class Criterion:
    """nn.CrossEntropyCriterion (disc) / SequencerCriterion(MaskZeroCriterion(ClassNLL sum)) (gen),
    model.lua:32-39."""

    def __init__(self, kind):
        self.kind = kind
        self.engine = None

    def cuda(self):
        return self

    def forward(self, decOut, batch: Batch) -> float:
        return self.engine.criterion_forward(batch)

    def backward(self, decOut, batch: Batch):
        self.engine.criterion_backward(batch)
        return None


class Sequential:
    """nn.Sequential():add(enc):add(dec) (model.lua:42): owns the engine = the flat parameter vector."""

    def __init__(self, enc: EncoderModule, dec: DecoderModule, params, seed=1234):
        self.enc, self.dec, self.params = enc, dec, params
        self.engine = None
        self._seed = seed

    def cuda(self):                                            # model.lua:48-51
        if self.engine is None:
            self.engine = Engine(self.params)
            self.enc.engine = self.dec.engine = self.engine
            self.engine.set_parameters(init_parameters(self.engine.params, self._seed))
        return self

    def get(self, i):                                          # model.lua:53-54
        return self.enc if i == 1 else self.dec

    def getParameters(self):                                   # model.lua:55 -> flat device buffers
        w, dw = self.engine.param_buffers()
        n = self.engine.num_params
        return DeviceTensor(self.engine, w, (n,)), DeviceTensor(self.engine, dw, (n,))

    def training(self):                                        # model.lua:57
        self.engine.set_training(1)

    def evaluate(self):                                        # model.lua:111
        self.engine.set_training(0)

    def zeroGradParameters(self):                              # model.lua:68
        self.engine.zero_grad()
