"""CPU restatement of the reference dataloader's tensor preparation and batch indexing.

TEST INFRASTRUCTURE ONLY (see oracle/__init__.py): imported by tests/ and bench.py's CPU leg, never by the product path.
PARITY UNPINNED: the reference ships no tests or golden vectors for this path and Lua/Torch7 cannot run here, so
this file restates /root/reference/dataloader.lua + utils.lua literally (same loops, same order of writes, 1-based
index arithmetic kept in the comments) and is the only anchor for the device implementation in
visdial_b200/csrc/corpus.cu.

Input: the per-split datasets of visdial_data.h5 as data/prepro.py:105-183 writes them (dict of numpy arrays).
Every function cites the reference lines it follows.
"""
from __future__ import annotations

from typing import Dict

import numpy as np


def right_align(sequences: np.ndarray, lengths: np.ndarray) -> np.ndarray:
    """utils.rightAlign, utils.lua:6-45 — including the `break` at the first zero-length round (:20-22)."""
    out = np.zeros_like(sequences)                                   # :9
    if sequences.ndim == 3:
        n, cnt, M = sequences.shape                                  # :13-15
        for im in range(n):
            for q in range(cnt):
                L = int(lengths[im, q])
                if L == 0:                                           # :20-22
                    break
                out[im, q, M - L:M] = sequences[im, q, 0:L]          # :25-26
    elif sequences.ndim == 2:
        n, M = sequences.shape                                       # :31-32
        for im in range(n):
            L = int(lengths[im])
            if L > 0:                                                # :36
                out[im, M - L:M] = sequences[im, 0:L]                # :38-39
    return out


def process_answers(ans: np.ndarray, ans_len: np.ndarray, start: int, end: int):
    """dataloader:processAnswers, dataloader.lua:159-199.  Returns (ans_in, ans_out, ans_len + 1)."""
    n, R, La = ans.shape                                             # :164-166
    dec_in = np.zeros((n, R, La + 1), dtype=np.int64)                # :168-169
    dec_out = np.zeros((n, R, La + 1), dtype=np.int64)
    dec_in[:, :, 0] = start                                          # :172
    for th in range(n):
        for rd in range(R):
            L = int(ans_len[th, rd])                                 # :178
            if L > 0:                                                # :181
                dec_in[th, rd, 1:L + 1] = ans[th, rd, 0:L]           # :182-183
                dec_out[th, rd, 0:L] = ans[th, rd, 0:L]              # :185-186
            dec_out[th, rd, L] = end                                 # :193 (also when the answer is empty)
    return dec_in, dec_out, ans_len + 1                              # :196-198


def process_history(cap, cap_len, ques, ques_len, ans, ans_len, concat: bool, end: int, max_history_len: int):
    """dataloader:processHistory, dataloader.lua:202-278.  Returns (hist, hist_len, maxHistoryLen)."""
    n, R, Lq = ques.shape                                            # :207,:211-213
    La = ans.shape[2]
    if concat:
        max_history_len = min(R * (Lq + La), 300)                    # :217
        history = np.zeros((n, R, max_history_len), dtype=np.int64)  # :219-221
    else:
        history = np.zeros((n, R, Lq + La), dtype=np.int64)          # :223-225
    hist_len = np.zeros((n, R), dtype=np.int64)
    for th in range(n):
        lenC = int(cap_len[th])                                      # :230
        lenH = 0
        for rd in range(R):
            if rd == 0:
                history[th, rd, 0:Lq + La] = cap[th, 0:Lq + La]      # :235-236 (raises if the caption tensor is narrower)
                lenH = min(lenC, Lq + La)                            # :237
            else:
                lenQ = int(ques_len[th, rd - 1])                     # :239-240
                lenA = int(ans_len[th, rd - 1])
                if concat:
                    history[th, rd, 0:lenH] = history[th, rd - 1, 0:lenH]        # :243-244
                    history[th, rd, lenH] = end                                   # :245 (IndexError past the width, as Lua)
                    if lenQ > 0:
                        if lenH + 1 + lenQ > history.shape[2]:
                            raise IndexError("history overflows maxHistoryLen (dataloader.lua:247 would raise)")
                        history[th, rd, lenH + 1:lenH + 1 + lenQ] = ques[th, rd - 1, 0:lenQ]                # :246-249
                    if lenA > 0:
                        if lenH + 1 + lenQ + lenA > history.shape[2]:
                            raise IndexError("history overflows maxHistoryLen (dataloader.lua:251 would raise)")
                        history[th, rd, lenH + 1 + lenQ:lenH + 1 + lenQ + lenA] = ans[th, rd - 1, 0:lenA]    # :250-253
                    lenH = lenH + lenQ + lenA + 1                                 # :254
                else:
                    if lenQ > 0:
                        history[th, rd, 0:lenQ] = ques[th, rd - 1, 0:lenQ]       # :257-260
                    if lenA > 0:
                        history[th, rd, lenQ:lenQ + lenA] = ans[th, rd - 1, 0:lenA]   # :261-264
                    lenH = lenA + lenQ                                            # :265
            hist_len[th, rd] = lenH                                  # :269
    return right_align(history, hist_len), hist_len, max_history_len  # :274-276


def process_options(opt_list: np.ndarray, opt_len: np.ndarray, start: int, end: int):
    """dataloader:processOptions, dataloader.lua:281-321.  Returns (opt_in, opt_out, opt_len + 1)."""
    m, La = opt_list.shape                                           # :283-288
    dec_in = np.zeros((m, La + 1), dtype=np.int64)                   # :290-291
    dec_out = np.zeros((m, La + 1), dtype=np.int64)
    dec_in[:, 0] = start                                             # :294
    for i in range(m):
        L = int(opt_len[i])                                          # :303
        if L > 0:                                                    # :306
            dec_in[i, 1:L + 1] = opt_list[i, 0:L]                    # :307
            dec_out[i, 0:L] = opt_list[i, 0:L]                       # :309
            dec_out[i, L] = end                                      # :310 (only for non-empty options)
    return dec_in, dec_out, opt_len + 1                              # :316-318


def prepare_images(images: np.ndarray, img_norm: bool, att: bool) -> np.ndarray:
    """dataloader.lua:59-73: optional L2 norm over dim 2, then N x C x S x S -> N x S x S x C for attention encoders."""
    x = images.astype(np.float32)
    if img_norm:
        nm = np.sqrt(np.sum(x * x, axis=1, keepdims=True, dtype=np.float32))     # :65
        x = (x / nm).astype(np.float32)                                          # :66
    if att:
        x = np.ascontiguousarray(np.transpose(x, (0, 2, 3, 1)))                  # :70-72
    return x


class DataloaderOracle:
    """dataloader:initialize's per-split state after prepareDataset (dataloader.lua:132-156), one split."""

    def __init__(self, raw: Dict[str, np.ndarray], *, use_history: bool, concat_history: bool, use_im: bool,
                 start: int, end: int, max_history_len: int = 60, img_norm: bool = False, att: bool = False):
        self.raw = raw
        self.useHistory, self.concatHistory, self.useIm = use_history, concat_history, use_im
        self.maxHistoryLen = max_history_len                                      # :142
        self.ques_len = raw["ques_length"].astype(np.int64)
        self.ques_fwd = right_align(raw["ques"].astype(np.int64), self.ques_len)  # :146-147
        if use_history:                                                           # :150
            self.hist, self.hist_len, self.maxHistoryLen = process_history(
                raw["cap"], raw["cap_length"], raw["ques"], raw["ques_length"], raw["ans"], raw["ans_length"],
                concat_history, end, max_history_len)
        self.opt_in, self.opt_out, self.opt_len = process_options(
            raw["opt_list"].astype(np.int64), raw["opt_length"].astype(np.int64), start, end)      # :153
        self.ans_in, self.ans_out, self.ans_len = process_answers(
            raw["ans"].astype(np.int64), raw["ans_length"].astype(np.int64), start, end)           # :155
        if use_im:
            self.img_fv = prepare_images(raw["images"], img_norm, att)            # :59-73
            self.img_pos = raw["img_pos"].astype(np.int64)                        # 0-based here (:76-77 adds Lua's 1)

    def get_index_data(self, inds: np.ndarray) -> Dict[str, np.ndarray]:
        """dataloader.getIndexData, dataloader.lua:378-433.  `inds` 0-based."""
        out = {}
        max_q = int(self.ques_len[inds].max())                                    # :380-381
        out["ques_fwd"] = self.ques_fwd[inds][:, :, self.ques_fwd.shape[2] - max_q:]   # :383-384 {-maxQuesLen,-1}
        if self.useHistory:
            max_h = min(int(self.hist_len[inds].max()), self.maxHistoryLen)       # :388-389
            out["hist"] = self.hist[inds][:, :, self.hist.shape[2] - max_h:]      # :390-391
        if self.useIm:
            out["img_feat"] = self.img_fv[self.img_pos[inds]]                     # :395-397
        max_a = int(self.ans_len[inds].max())                                     # :401-402
        out["answer_in"] = self.ans_in[inds][:, :, :max_a]                        # :404-407
        out["answer_out"] = self.ans_out[inds][:, :, :max_a]
        if "ans_index" in self.raw and self.raw["ans_index"] is not None:
            out["answer_ind"] = self.raw["ans_index"][inds].astype(np.int64)      # :427-430
        return out

    def get_index_option(self, inds: np.ndarray, decoder: str):
        """dataloader.getIndexOption, dataloader.lua:436-478."""
        opt_inds = self.raw["opt"][inds].astype(np.int64)                         # :441 / :465  (1-based rows)
        vec = opt_inds.reshape(-1) - 1                                            # :442 / :466
        B, R, K = opt_inds.shape
        if decoder == "gen":
            max_o = int(self.opt_len[vec].max())                                  # :444-445
            oi = self.opt_in[vec].reshape(B, R, K, -1)[:, :, :, :max_o]           # :447-450
            oo = self.opt_out[vec].reshape(B, R, K, -1)[:, :, :, :max_o]          # :452-455
            return {"option_in": oi, "option_out": oo}
        return self.raw["opt_list"][vec].reshape(B, R, K, -1).astype(np.int64)    # :468-470

    def get_batch(self, inds: np.ndarray, decoder: str, test_batch: bool) -> Dict[str, np.ndarray]:
        """getTrainBatch (dataloader.lua:324-341, given the sampled inds) / getTestBatch (:344-375)."""
        out = self.get_index_data(inds)
        if decoder == "disc":                                                     # :331-337 / :362-368
            o = self.get_index_option(inds, "disc")
            out["options"] = o.reshape(o.shape[0] * o.shape[1], o.shape[2], -1)
            if "answer_ind" in out:
                out["answer_ind"] = out["answer_ind"].reshape(-1)
        elif test_batch:                                                          # :369-371
            out.update(self.get_index_option(inds, "gen"))
        if test_batch and self.raw.get("num_rounds") is not None:
            out["num_rounds"] = self.raw["num_rounds"][inds].astype(np.int64)     # :373
        return out
