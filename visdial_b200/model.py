"""Mirror of the reference's `Model` class (/root/reference/model.lua:8-430) over the C engine.
Same method names, argument meaning and call order as the Lua original, so that the parity tests
read like the reference's own driver code.  `generateAnswers` (beam search / sampling, model.lua:432-613) steps the
decoder on the device through vd_gen_decoder_step and keeps the hypothesis bookkeeping on the host like the Lua."""
from __future__ import annotations

import numpy as np

from . import decoders, encoders
from .engine import Batch, DEFAULT_PARAMS, derive_flags
from .modules import Criterion, Sequential


def _as_batch(b):
    """getTestBatch returns (batch, nextStartId) in the reference (dataloader.lua:375); accept either form."""
    if isinstance(b, tuple):
        b = b[0]
    return b if isinstance(b, Batch) else Batch(b)


def _image_id(x):
    """dataloader.lua:42-44: `tonumber(string.match(name, '000%d+'))` on the json image names; ints pass through."""
    if isinstance(x, (int, np.integer)):
        return int(x)
    import re
    mt = re.search(r"000\d+", str(x))
    return int(mt.group(0)) if mt else x


def _num_tokens(batch) -> int:
    """answerOut:gt(0):sum() (model.lua:78); device batches carry it from the host length table."""
    n = getattr(batch, "num_answer_tokens", None)
    return int(n) if n is not None else int((batch["answer_out"] > 0).sum())


class Model:
    def __init__(self, params: dict, seed: int = 1234):                  # model.lua:11-62
        p = dict(DEFAULT_PARAMS)
        p.update(params)
        self.params = derive_flags(p)
        encoder = encoders.load(p["encoder"])                             # :19-20
        decoder = decoders.load(p["decoder"])                             # :22-23
        enc = encoder.model(p)                                            # :25
        dec = decoder.model(p, enc)                                       # :26
        self.forwardConnect = decoder.forwardConnect                      # :28-29
        self.backwardConnect = decoder.backwardConnect
        self.decoderConnect = getattr(decoder, "decoderConnect", None)
        self.criterion = Criterion(p["decoder"])                          # :32-39
        self.wrapper = Sequential(enc, dec, p, seed)                      # :42 (weight-init.lua is a no-op here)
        if p["gpuid"] < 0:
            raise ValueError("visdial_b200 has no CPU path: gpuid must be >= 0")
        self.wrapper.cuda()                                               # :48-51
        self.criterion.engine = self.wrapper.engine
        self.encoder = self.wrapper.get(1)                                # :53-54
        self.decoder = self.wrapper.get(2)
        self.wrapperW, self.wrapperdW = self.wrapper.getParameters()      # :55
        self.wrapper.training()                                           # :57
        self.optims = {"learningRate": p["learningRate"]}                 # :60-61
        self.runningLoss = 0.0                                            # global `runningLoss`, train.lua:89
        self.engine = self.wrapper.engine
        self.iteration = 0

    # ---------------------------------------------------------------------------------------------
    def trainIteration(self, dataloader):                                 # model.lua:66-106
        self.wrapper.zeroGradParameters()                                 # :68
        batch = dataloader.getTrainBatch(self.params)                     # :71
        if not isinstance(batch, Batch):
            batch = Batch(batch)
        self.iteration += 1
        self.engine.set_dropout_seed(self.params.get("seed", 1234), self.iteration)
        curLoss = self.forwardBackward(batch)                             # :74
        if self.params["decoder"] == "gen":                               # :76-85
            numTokens = _num_tokens(batch)
            cur = curLoss / max(numTokens, 1)
        else:                                                             # :86-93
            cur = curLoss
        self.runningLoss = 0.95 * self.runningLoss + 0.05 * cur if self.runningLoss > 0 else cur
        # :96-99 clamp(-5,5) + adam, fused with the gradient all-reduce when world > 1
        self.engine.clamp_adam_step(self.optims["learningRate"])
        if self.optims["learningRate"] > self.params["minLRate"]:          # :102-105
            self.optims["learningRate"] *= self.params["lrDecayRate"]
        return curLoss

    def forwardBackward(self, batch, onlyForward=False, encOutOnly=False):   # model.lua:249-342
        if not isinstance(batch, Batch):
            batch = Batch(batch)
        # :252-294 (time-major views, image repeat, MN mask) are index arithmetic inside the engine
        encOut = self.encoder.forward(batch)                               # :297
        seqLen = batch.c.Tq
        self.decoder._last_batch = batch
        self.forwardConnect(self.encoder, self.decoder, encOut, seqLen)    # :300
        if encOutOnly:
            return encOut                                                  # :302
        if self.params["decoder"] == "gen":
            # decOut only travels on to the criterion here: let the engine keep the (rows, V) log-probabilities on chip
            self.engine.set_lazy_decout(True)
            try:
                decOut = self.decoder.forward(batch)                       # :313
            finally:
                self.engine.set_lazy_decout(False)
            curLoss = self.criterion.forward(decOut, batch)                # :314
            if not onlyForward:
                gradCriterionOut = self.criterion.backward(decOut, batch)  # :318
                self.decoder.backward(batch, gradCriterionOut)             # :319
                gradDecOut = self.backwardConnect(self.encoder, self.decoder)   # :322
                self.encoder.backward(batch, gradDecOut)                   # :323
        else:
            decOut = self.decoder.forward(batch)                           # :329
            curLoss = self.criterion.forward(decOut, batch)                # :330
            if not onlyForward:
                gradCriterionOut = self.criterion.backward(decOut, batch)  # :334
                t = self.decoder.backward(batch, gradCriterionOut)         # :335
                self.encoder.backward(batch, t[1])                         # :337 (t[2] in Lua)
        return curLoss

    def retrieveBatch(self, batch):                                        # model.lua:344-430
        if not isinstance(batch, Batch):
            batch = Batch(batch)
        return self.engine.retrieve(batch, use_gt=bool(self.params.get("useGt", True)))

    # ---------------------------------------------------------------------------------------------
    def evaluate(self, dataloader, dtype="val"):                           # model.lua:109-139
        self.wrapper.evaluate()
        curLoss, numTokens, n = 0.0, 0, 0
        numThreads = dataloader.numThreads[dtype]
        for startId in range(0, numThreads, self.params["batchSize"]):
            batch = _as_batch(dataloader.getTestBatch(startId, self.params, dtype))
            curLoss += self.forwardBackward(batch, True)
            if self.params["decoder"] == "gen":
                numTokens += _num_tokens(batch)
            n += 1
        self.wrapper.training()
        # :133 divides the summed loss by the answer-token count; for disc (a mean criterion per batch, no token count in
        # the disc batch table) the mean over batches is the comparable figure
        return curLoss / max(numTokens, 1) if self.params["decoder"] == "gen" else curLoss / max(n, 1)

    def _rank_walk(self, dataloader, dtype, use_gt):
        """The getTestBatch loop shared by retrieve / predict (model.lua:153-166, :208-221)."""
        self.wrapper.evaluate()
        out = []
        numThreads = dataloader.numThreads[dtype]
        for startId in range(0, numThreads, self.params["batchSize"]):
            batch = _as_batch(dataloader.getTestBatch(startId, self.params, dtype))
            out.append(self.engine.retrieve(batch, use_gt=use_gt))
        self.wrapper.training()
        return out

    def _round_table(self, dataloader, dtype, ranks, last_round_only):
        """model.lua:175-185 / :227-243: one {image_id, round_id, ranks} entry per (dialog, round)."""
        img = getattr(dataloader, "unique_img_" + dtype, None)
        nr = getattr(dataloader, dtype + "_num_rounds", None)
        first = dataloader.part[dtype][0] if dtype in getattr(dataloader, "part", {}) else 0
        table = []
        for i in range(ranks.shape[0]):
            n_i = int(nr[first + i]) if nr is not None else ranks.shape[1]
            iid = _image_id(img[first + i]) if img is not None else first + i
            rounds = [n_i] if last_round_only else range(1, n_i + 1)
            for j in rounds:
                r = ranks[i, j - 1]
                table.append({"image_id": iid, "round_id": int(j), "ranks": r.tolist() if np.ndim(r) else int(r)})
        return table

    def retrieve(self, dataloader, dtype="val", as_table=False, verbose=False):   # model.lua:142-189
        """Ground-truth ranks of a split, (numThreads, maxQuesCount).  `as_table=True` returns what the reference returns:
        the {image_id, round_id, ranks} list, after printing utils.processRanks of the matrix."""
        use_gt = bool(self.params.get("useGt", True))
        ranks = np.concatenate([r.reshape(-1, self.params["maxQuesCount"]) for r in self._rank_walk(dataloader, dtype, use_gt)], 0)
        if not as_table:
            return ranks
        from .utils import processRanks
        processRanks(ranks, verbose=verbose)                                       # :170
        return self._round_table(dataloader, dtype, ranks, last_round_only=False)

    def predict(self, dataloader, dtype="val"):                                   # model.lua:191-246
        """Full 100-option rank lists (evaluate.lua's EvalAI dump): every round of a val dialog, the last round of a test one."""
        K = int(self.params.get("numOptions", 100))
        ranks = np.concatenate([r.reshape(-1, self.params["maxQuesCount"], K) for r in self._rank_walk(dataloader, dtype, False)], 0)
        return self._round_table(dataloader, dtype, ranks, last_round_only=(dtype == "test"))

    # ---- beam search / sampling (model.lua:432-613, generate.lua) ---------------------------------------------
    def generateAnswers(self, dataloader, dtype="val", params=None, strict=True):
        """Model:generateAnswers: per dialog, encoder forward on the device, then the decoder driven one step at a time
        through vd_gen_decoder_step; the hypothesis bookkeeping runs on the host exactly as the Lua does (it indexes
        GPU tensors scalar by scalar).  Returns the reference's answerTable with token-id lists (and text when the
        dataloader carries ind2word).  `strict=False` yields None for a round where no beam reached <END> (the
        reference indexes nil there, model.lua:575)."""
        if self.params["decoder"] == "disc":                                            # :434-437
            raise ValueError("Sampling/beam search only for generative model")
        params = params or {}
        sampleWords = bool(params.get("sampleWords", 0) == 1)                           # :443
        temperature = float(params.get("temperature", 1.0))
        beamSize, beamLen = int(params.get("beamSize", 5)), int(params.get("beamLen", 20))
        startToken, endToken = dataloader.word2ind["<START>"], dataloader.word2ind["<END>"]   # :453-454
        numThreads = int(params.get("maxThreads") or dataloader.numThreads[dtype])      # :455
        ind2word = getattr(dataloader, "ind2word", None)
        rng = np.random.default_rng(int(params.get("seed", 1234)))
        eng, H = self.engine, self.params["rnnHiddenSize"]
        words = (lambda ids: " ".join(ind2word.get(int(t), "<UNK>") for t in ids if int(t) > 0)) if ind2word else None
        state_buf = [eng.device_alloc(max(beamSize, self.params["maxQuesCount"]) * H * 4) for _ in range(4)]
        answerTable = []
        try:
            for convId in range(numThreads):
                self.wrapper.evaluate()                                                 # :460
                batch = dataloader.getIndexData(np.array([convId]), self.params, dtype)  # :462-463
                encOut = self.forwardBackward(batch, True, True).numpy()                # :467 (N,H), N = 10
                ques = batch["ques_fwd"].reshape(-1, batch.c.Tq)
                N = encOut.shape[0]
                layers = [eng.encoder_rnn_state(l, N) for l in range(2)]
                has_layers = layers[0][0] is not None
                encH = [(layers[l][0].numpy(), layers[l][1].numpy()) for l in range(2)] if has_layers else None
                threadAnswers = []

                def step(tokens, Hs, Cs):
                    n = len(tokens)
                    for i, a in enumerate(Hs + Cs):
                        eng.upload(state_buf[i], a[:n])
                    return eng.gen_decoder_step(tokens, state_buf[0:2], state_buf[2:4])

                if not sampleWords and not params.get("hostBeam"):
                    # All N rounds of the dialog are searched at once (N * beamSize hypotheses per decoder step).  The LSTM state
                    # and the (rows, V) log-probabilities stay on the device (vd_gen_beam_step): per step the tokens and parent
                    # indices go up, the beamSize best (log-prob, class) pairs of every hypothesis come down; the candidate merge
                    # of model.lua:529-569 — including its quirks — runs on those few numbers.
                    bs = beamSize
                    beams = np.zeros((N, beamLen, bs), dtype=np.int64)                  # :479
                    beams[:, 0, :] = startToken                                         # :506
                    scores = np.zeros((N, bs), dtype=np.float64)                        # :507
                    finish = [[] for _ in range(N)]                                     # :508
                    if has_layers:                                                      # :482-491
                        iH = [np.repeat(encH[0][0], bs, 0), np.repeat(encOut, bs, 0)]
                        iC = [np.repeat(encH[0][1], bs, 0), np.repeat(encH[1][1], bs, 0)]
                    else:                                                               # :493-501
                        z = np.zeros((N * bs, H), np.float32)
                        iH, iC = [z, np.repeat(encOut, bs, 0)], [z, z]
                    parent = None
                    for stp in range(1, beamLen):                                       # :510
                        topv, topi = eng.gen_beam_step(beams[:, stp - 1, :].reshape(-1), parent, iH, iC, bs)   # :519-542
                        parent = -1 - np.arange(N * bs, dtype=np.int32)                 # default: the column keeps its old content
                        exploreSize = 1 if stp == 1 else bs                             # :516
                        for it in range(N):
                            cands = []
                            for wordId in range(exploreSize):                           # :529
                                r = it * bs + wordId
                                for candId in range(bs):                                # :544
                                    tok = int(topi[r, candId]) + 1
                                    sc = float(scores[it, wordId]) + float(topv[r, candId])
                                    if tok == endToken:                                 # :548
                                        cb = beams[it, :, wordId].copy()
                                        cb[stp] = tok
                                        finish[it].append({"beam": cb, "length": stp + 1, "score": sc})
                                    else:
                                        cands.append((sc, wordId, tok))
                            cands.sort(key=lambda t: -t[0])                             # :558 (stable)
                            old = beams[it].copy()
                            for candId in range(min(len(cands), bs)):                   # :560-569
                                sc, wordId, tok = cands[candId]
                                beams[it, :, candId] = old[:, wordId]
                                beams[it, stp, candId] = tok
                                scores[it, candId] = sc
                                parent[it * bs + candId] = it * bs + wordId
                    for it in range(N):
                        finish[it].sort(key=lambda d: -d["score"])                      # :572
                        if not finish[it]:
                            if strict:
                                raise IndexError("no beam reached <END> within beamLen (model.lua:575 indexes nil here)")
                            threadAnswers.append(None)
                            continue
                        best = finish[it][0]
                        entry = {"question": ques[it].tolist(), "answer": best["beam"].tolist(), "score": best["score"],
                                 "length": best["length"]}
                        if words:
                            entry["question_text"], entry["answer_text"] = words(ques[it]), words(best["beam"])
                        threadAnswers.append(entry)
                elif not sampleWords:
                    # the reference's own loop structure (one round at a time, log-probabilities and state through the host):
                    # kept as the cross-check of the batched search above (params.hostBeam = 1)
                    for it in range(N):                                                 # :472
                        beams = np.zeros((beamLen, beamSize), dtype=np.int64)           # :479
                        if has_layers:                                                  # :482-491
                            Hs = [encH[0][0][it], encOut[it]]
                            Cs = [encH[0][1][it], encH[1][1][it]]
                        else:                                                           # :493-501
                            Hs = [np.zeros(H, np.float32), encOut[it]]
                            Cs = [np.zeros(H, np.float32), np.zeros(H, np.float32)]
                        Hs = [np.repeat(h[None], beamSize, 0).astype(np.float32) for h in Hs]
                        Cs = [np.repeat(c[None], beamSize, 0).astype(np.float32) for c in Cs]
                        beams[0] = startToken                                           # :506
                        scores = np.zeros(beamSize, dtype=np.float64)                   # :507
                        finishBeams = []
                        for stp in range(1, beamLen):                                   # :510
                            cands = []
                            exploreSize = 1 if stp == 1 else beamSize                   # :516
                            decOut, nH, nC = step(beams[stp - 1], Hs, Cs)               # :519-526
                            for wordId in range(exploreSize):                           # :529
                                order = np.argsort(-decOut[wordId], kind="stable")[:beamSize]   # :538-542 topk, sorted
                                for candId in range(beamSize):                          # :544
                                    candBeam = beams[:, wordId].copy()
                                    tok = int(order[candId]) + 1
                                    candBeam[stp] = tok
                                    sc = float(scores[wordId]) + float(decOut[wordId, order[candId]])
                                    if tok == endToken:                                 # :548
                                        finishBeams.append({"beam": candBeam, "length": stp + 1, "score": sc})
                                    else:
                                        cands.append((sc, candBeam, [h[wordId].copy() for h in nH], [c[wordId].copy() for c in nC]))
                            cands.sort(key=lambda t: -t[0])                             # :558
                            for candId in range(min(len(cands), beamSize)):             # :560-569
                                beams[:, candId] = cands[candId][1]
                                for lv in range(2):
                                    Hs[lv][candId] = cands[candId][2][lv]
                                    Cs[lv][candId] = cands[candId][3][lv]
                                scores[candId] = cands[candId][0]
                        finishBeams.sort(key=lambda d: -d["score"])                     # :572
                        if not finishBeams:
                            if strict:
                                raise IndexError("no beam reached <END> within beamLen (model.lua:575 indexes nil here)")
                            threadAnswers.append(None)
                            continue
                        best = finishBeams[0]
                        entry = {"question": ques[it].tolist(), "answer": best["beam"].tolist(), "score": best["score"],
                                 "length": best["length"]}
                        if words:
                            entry["question_text"], entry["answer_text"] = words(ques[it]), words(best["beam"])
                        threadAnswers.append(entry)
                else:                                                                   # :581-602
                    if has_layers:                                                      # forwardConnect, gen.lua:30-42
                        Hs, Cs = [encH[0][0], encOut], [encH[0][1], encH[1][1]]
                    else:
                        Hs, Cs = [np.zeros((N, H), np.float32), encOut], [np.zeros((N, H), np.float32)] * 2
                    tok = np.full(N, startToken, dtype=np.int64)
                    seq = [tok.copy()]
                    for _ in range(beamLen):
                        decOut, Hs, Cs = step(tok, Hs, Cs)                              # :586-588 (+ decoderConnect)
                        p = np.exp(decOut.astype(np.float64) / temperature)             # :590
                        p /= p.sum(1, keepdims=True)
                        tok = np.array([rng.choice(p.shape[1], p=p[i]) + 1 for i in range(N)], dtype=np.int64)
                        seq.append(tok.copy())
                    ans = np.stack(seq, 1)
                    for it in range(N):
                        entry = {"question": ques[it].tolist(), "answer": ans[it].tolist()}
                        if words:
                            entry["question_text"], entry["answer_text"] = words(ques[it]), words(ans[it])
                        threadAnswers.append(entry)
                self.wrapper.training()                                                 # :605
                img = getattr(dataloader, "unique_img_" + dtype, None)
                # getIndexData hands out this rank's slice of the split: the image list is indexed by the global dialog id
                gid = convId + (dataloader.part[dtype][0] if hasattr(dataloader, "part") and dtype in getattr(dataloader, "part", {}) else 0)
                answerTable.append({"image_id": _image_id(img[gid]) if img is not None else gid, "dialog": threadAnswers})   # :606
        finally:
            for p in state_buf:
                eng.device_free(p)
        return answerTable

    # ---- checkpoints (train.lua:33-34,78-80,99-102,120-121; evaluate.lua:58-91) -------------------------------
    def save(self, path: str, final: bool = False):
        """torch.save(path, {modelW, optims, modelParams}) — `final` = the model_final.t7 form (float weights, no optims)."""
        from .checkpoint import save_checkpoint
        save_checkpoint(self, path, final)

    def load(self, path: str, permutation=None, restore_adam_state: bool = False):
        """wrapperW:copy(savedModel.modelW); optims.learningRate = savedModel.optims.learningRate."""
        from .checkpoint import load_checkpoint, restore
        ck = load_checkpoint(path, permutation)
        restore(self, ck, restore_adam_state)
        return ck
