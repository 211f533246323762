"""Shared test helpers: small configs, batches, oracle <-> engine parameter plumbing."""
import numpy as np
import torch

from visdial_b200 import engine as E
from visdial_b200.synthetic import make_batch

CONFIGS = [("lf-ques", "gen"), ("lf-ques-im-hist", "disc"), ("hrea-ques-im-hist", "gen"),
           ("mn-att-ques-im-hist", "disc"), ("lf-ques", "disc"), ("mn-att-ques-im-hist", "gen"),
           ("hrea-ques-im-hist", "disc"), ("lf-ques-im-hist", "gen"),
           # the seven sub-graph encoders of encoders/*.lua (each once with either decoder; the ones that export
           # rnnLayers to decoderConnect also with gen)
           ("lf-ques-im", "disc"), ("lf-ques-im", "gen"), ("lf-ques-hist", "gen"), ("hre-ques-hist", "disc"),
           ("hre-ques-hist", "gen"), ("hre-ques-im-hist", "gen"), ("mn-ques-hist", "gen"), ("mn-ques-im-hist", "disc"),
           ("lf-att-ques-im-hist", "disc"), ("lf-att-ques-im-hist", "gen")]
ALL_ENCODERS = ("lf-ques", "lf-ques-im", "lf-ques-hist", "lf-ques-im-hist", "lf-att-ques-im-hist", "hre-ques-hist",
                "hre-ques-im-hist", "hrea-ques-im-hist", "mn-ques-hist", "mn-ques-im-hist", "mn-att-ques-im-hist")


def small_params(encoder, decoder, **kw):
    p = dict(E.DEFAULT_PARAMS)
    p.update(encoder=encoder, decoder=decoder, vocabSize=40, embedSize=12, rnnHiddenSize=32, numLayers=2,
             imgFeatureSize=8 if "att" in encoder else 24, imgSpatialSize=3, imgEmbedSize=8,
             commonEmbeddingSize=16, numAttentionLayers=1, maxQuesCount=10, numOptions=7, dropout=0.5, gpuid=0)
    p.update(kw)
    return E.derive_flags(p)


def full_params(encoder, decoder, **kw):
    p = dict(E.DEFAULT_PARAMS)
    p.update(encoder=encoder, decoder=decoder, vocabSize=10000,
             imgFeatureSize=512 if "att" in encoder else 4096)
    p.update(kw)
    return E.derive_flags(p)


def small_batch(params, B=3, seed=7, gen_eval=False):
    return make_batch(params, B, seed=seed, max_ques_len=6, max_ans_len=5, max_cap_len=8, max_hist_len=9,
                      max_hist_concat=30, gen_eval=gen_eval, empty_round_every=2)


def torch_batch(batch):
    out = {}
    for k, v in batch.items():
        t = torch.from_numpy(np.ascontiguousarray(v))
        out[k] = t.long() if v.dtype.kind in "iu" else t
    return out


def torch_params(params, flat, dtype=torch.float32):
    return {k: torch.from_numpy(np.array(v)).to(dtype) for k, v in E.split_parameters(params, flat).items()}


def flat_from_named(params, named):
    segs, n = E.layout(params)
    flat = np.zeros(n, dtype=np.float32)
    for s in segs:
        flat[s.offset:s.offset + s.size] = np.asarray(named[s.name].detach().cpu().numpy(), dtype=np.float32).reshape(-1)
    return flat


def seg_slices(params):
    segs, _ = E.layout(params)
    return {s.name: slice(s.offset, s.offset + s.size) for s in segs}
