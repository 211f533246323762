"""TEST INFRASTRUCTURE ONLY (see oracle/__init__.py) — PARITY UNPINNED.

Op-for-op CPU restatement (PyTorch-CPU tensors, fp32 with an fp64 switch) of the
reference hot path: `Model:forwardBackward` (/root/reference/model.lua:249-342),
`Model:retrieveBatch` (model.lua:344-430), the encoder graphs
encoders/{lf-ques,lf-ques-im-hist,hrea-ques-im-hist,mn-att-ques-im-hist}.lua, the
decoder graphs decoders/{disc,gen}.lua, the custom mask modules model_utils/*.lua,
clamp + adam (model.lua:96-105, model_utils/optim_updates.lua:62-91) and
utils.computeLhood / utils.computeRanks (utils.lua:86-128).

The third-party modules the reference instantiates (nn.Linear, nn.SeqLSTM,
nn.LookupTableMaskZero, nn.MM, nn.SoftMax, nn.Dropout, the criterions) are NOT in
/root/reference and are not version-pinned by it (README.md:45-60: `luarocks install
nn nngraph`, Element-Research/rnn at floating HEAD).  Their published semantics are
restated here; every such function says [upstream].

"Reference structure" is kept on purpose (per-timestep addmm LSTM loop, 100
sequential option-LSTM passes, materialised repeatTensor) so that the same code is
the CPU baseline timed by bench.py (kind "port").

Parameters are a dict name -> tensor; the names are the segment names of the
engine's flat parameter vector (include/visdial_b200.h).
"""
from __future__ import annotations

import math
from typing import Callable, Dict, Optional

import torch

Tensor = torch.Tensor

# Dropout sites: the mask of a site is indexed by the row-major linear
# index of the reference tensor at that site.
SITE_QEMBED = 0       # mn-att-ques-im-hist.lua:24   (Tq,N,E)
SITE_HEMBED = 1       # mn-att-ques-im-hist.lua:25   (Th,N,E)
SITE_HATT = 2         # mn-att-ques-im-hist.lua:64   (N,H)
SITE_IMG_TR = 3       # mn-att-ques-im-hist.lua:74   (N,196,H)
SITE_U_OUT = 5        # mn-att-ques-im-hist.lua:106  (N,H)
SITE_FUSION = 6       # lf-ques.lua:30, lf-ques-im-hist.lua:55  (N,K)
SITE_IMG_FC7 = 7      # hrea-ques-im-hist.lua:48     (N,4096)
SITE_HOP0 = 16        # mn-att-ques-im-hist.lua:92   (N,196,Cm), + hop index


class Ctx:
    """Run mode: training (dropout on, masks from `mask_fn(site, shape)->0/1 tensor`) or eval."""

    def __init__(self, train: bool = False,
                 mask_fn: Optional[Callable[[int, tuple], Tensor]] = None,
                 structure: str = "reference"):
        self.train = train
        self.mask_fn = mask_fn
        self.structure = structure  # "reference" = 100 sequential option passes; "batched"


# ----------------------------------------------------------------------------------------
# [upstream] primitive modules
# ----------------------------------------------------------------------------------------

def lookup_table_mask_zero(weight: Tensor, ids: Tensor) -> Tensor:
    """[upstream rnn] nn.LookupTableMaskZero(V,E): table (V+1,E); id 0 (pad) reads row 1 of the
    Lua table, which is zeroed at every forward.  Here ids index `weight` directly (row 0 = pad).
    Built at mn-att-ques-im-hist.lua:21, lf-ques.lua:12, hrea-ques-im-hist.lua:20; shared with
    the decoders at disc.lua:12 and gen.lua:10.  The gradient is accumulated into every row,
    the pad row included (accGradParameters is the parent's on ids+1)."""
    z = torch.zeros_like(weight)
    z[0] = weight[0].detach()
    w = weight - z            # forward sees a zero pad row; d/dweight is the identity on every row
    return w[ids]


def dropout(ctx: Ctx, x: Tensor, p: float, site: int) -> Tensor:
    """[upstream nn] nn.Dropout(p) v2: train x*Bernoulli(1-p)/(1-p), eval identity."""
    if not ctx.train or p <= 0:
        return x
    m = ctx.mask_fn(site, tuple(x.shape)).to(x.dtype)
    return x * m / (1.0 - p)


def linear(x: Tensor, w: Tensor, b: Tensor) -> Tensor:
    """[upstream nn] nn.Linear: y = x W^T + b, W (out,in)."""
    return torch.addmm(b, x, w.t())


class _SeqLSTMFn(torch.autograd.Function):
    """[upstream rnn] nn.SeqLSTM forward + BPTT.  weight (D+H,4H): x rows
    then h rows; gate column blocks [i f o g]; maskzero resets (h,c,gates) to 0 on rows whose
    input vector is all-zero.  Per-timestep addmm loop like the Lua `for t=1,T`."""

    @staticmethod
    def forward(ctx, x, W, b, h0, c0, mask):
        T, N, D = x.shape
        H = b.numel() // 4
        Wx, Wh = W[:D], W[D:]
        h = x.new_zeros(T, N, H)
        c = x.new_zeros(T, N, H)
        gates = x.new_zeros(T, N, 4 * H)
        prev_h, prev_c = h0, c0
        for t in range(T):
            a = gates[t]
            a.copy_(b.expand(N, 4 * H))
            a.addmm_(x[t], Wx)
            a.addmm_(prev_h, Wh)
            a[:, :3 * H].sigmoid_()
            a[:, 3 * H:].tanh_()
            i, f, o, g = a[:, :H], a[:, H:2 * H], a[:, 2 * H:3 * H], a[:, 3 * H:]
            torch.mul(f, prev_c, out=c[t])
            c[t].addcmul_(i, g)
            torch.mul(o, torch.tanh(c[t]), out=h[t])
            if mask is not None:
                m = mask[t]
                h[t][m] = 0
                c[t][m] = 0
                a[m] = 0
            prev_h, prev_c = h[t], c[t]
        ctx.save_for_backward(x, W, h0, c0, h, c, gates)
        ctx.mask = mask
        return h, c

    @staticmethod
    def backward(ctx, grad_h, grad_c):
        x, W, h0, c0, h, c, gates = ctx.saved_tensors
        mask = ctx.mask
        T, N, D = x.shape
        H = h.shape[2]
        Wx, Wh = W[:D], W[D:]
        dx = torch.zeros_like(x)
        dW = torch.zeros_like(W)
        db = W.new_zeros(4 * H)
        dh_next = x.new_zeros(N, H)
        dc_next = x.new_zeros(N, H)
        for t in range(T - 1, -1, -1):
            prev_h = h[t - 1] if t > 0 else h0
            prev_c = c[t - 1] if t > 0 else c0
            a = gates[t]
            i, f, o, g = a[:, :H], a[:, H:2 * H], a[:, 2 * H:3 * H], a[:, 3 * H:]
            dh = grad_h[t] + dh_next
            dc = grad_c[t] + dc_next
            if mask is not None:
                m = mask[t]
                dh = dh.clone()
                dc = dc.clone()
                dh[m] = 0
                dc[m] = 0
            tc = torch.tanh(c[t])
            dc = dc + dh * o * (1 - tc * tc)
            da = torch.empty_like(a)
            da[:, :H] = dc * g * i * (1 - i)
            da[:, H:2 * H] = dc * prev_c * f * (1 - f)
            da[:, 2 * H:3 * H] = dh * tc * o * (1 - o)
            da[:, 3 * H:] = dc * i * (1 - g * g)
            dx[t] = da @ Wx.t()
            dW[:D].addmm_(x[t].t(), da)
            dW[D:].addmm_(prev_h.t(), da)
            db += da.sum(0)
            dh_next = da @ Wh.t()
            dc_next = dc * f
        return dx, dW, db, dh_next, dc_next, None


def seq_lstm(x: Tensor, W: Tensor, b: Tensor, h0: Optional[Tensor] = None,
             c0: Optional[Tensor] = None, maskzero: bool = False, manual_bptt: bool = True):
    """nn.SeqLSTM(D,H) on time-major x (T,N,D) -> (output (T,N,H), cell (T,N,H)).
    maskzero(): the mask is derived from the INPUT vector (all-zero row), as upstream does."""
    T, N, D = x.shape
    H = b.numel() // 4
    if h0 is None:
        h0 = x.new_zeros(N, H)
    if c0 is None:
        c0 = x.new_zeros(N, H)
    mask = (x.detach().abs().sum(-1) == 0) if maskzero else None
    if manual_bptt:
        return _SeqLSTMFn.apply(x, W, b, h0, c0, mask)
    # pure-autograd variant, used only to validate the hand-written BPTT above
    Wx, Wh = W[:D], W[D:]
    hs, cs = [], []
    ph, pc = h0, c0
    for t in range(T):
        a = b + x[t] @ Wx + ph @ Wh
        i, f, o = (torch.sigmoid(a[:, k * H:(k + 1) * H]) for k in range(3))
        g = torch.tanh(a[:, 3 * H:])
        nc = f * pc + i * g
        nh = o * torch.tanh(nc)
        if mask is not None:
            keep = (~mask[t]).to(x.dtype).unsqueeze(1)
            nc = nc * keep
            nh = nh * keep
        hs.append(nh)
        cs.append(nc)
        ph, pc = nh, nc
    return torch.stack(hs), torch.stack(cs)


def lstm_p(P: Dict[str, Tensor], name: str):
    return P[name + ".weight"], P[name + ".bias"]


# ----------------------------------------------------------------------------------------
# custom modules, fully specified in the reference tree
# ----------------------------------------------------------------------------------------

def mask_softmax(data: Tensor, mask: Tensor) -> Tensor:
    """model_utils/MaskSoftMax.lua:5-21: maskedFill(mask,-9999999) then SoftMax over the last dim
    of the 2-D input; backward (:23-46) is the softmax gradient, zero gradient for the mask."""
    return torch.softmax(data.masked_fill(mask.bool(), -9999999.0), dim=-1)


def mask_time(ques: Tensor, img_embed: Tensor) -> Tensor:
    """model_utils/MaskTime.lua:12-28: out[t,n,:] = imgEmbed[n,:] if ques[t,n] != 0 else 0;
    backward (:30-40) zeroes masked positions and sums over t."""
    keep = (ques != 0).to(img_embed.dtype).unsqueeze(-1)           # (T,N,1)
    return img_embed.unsqueeze(0) * keep


def mask_future(x: Tensor) -> Tensor:
    """model_utils/MaskFuture.lua:4-31: strict upper triangle (j>i) of each (n,n) slice -> 0."""
    n = x.shape[-1]
    tri = torch.triu(torch.ones(n, n, dtype=torch.bool), 1)
    return x.masked_fill(tri, 0.0)


def replace_zero(x: Tensor, constant: float) -> Tensor:
    """model_utils/ReplaceZero.lua:13-25: exact zeros -> constant; gradient 0 there."""
    return torch.where(x == 0, torch.full_like(x, constant), x)


# ----------------------------------------------------------------------------------------
# encoders
# ----------------------------------------------------------------------------------------

def _two_layer_lstm(P, prefix, x, maskzero=True):
    o1, c1 = seq_lstm(x, *lstm_p(P, prefix + ".lstm1"), maskzero=maskzero)
    o2, c2 = seq_lstm(o1, *lstm_p(P, prefix + ".lstm2"), maskzero=maskzero)
    return (o1, c1), (o2, c2)


def encoder_lf_ques(ctx: Ctx, cfg, P, inputs):
    """encoders/lf-ques.lua:3-36."""
    ques = inputs["ques"]                                            # (Tq,N)
    x = lookup_table_mask_zero(P["wordEmbed.weight"], ques)         # :12-13
    l1, l2 = _two_layer_lstm(P, "ques", x)                           # :16-24
    q = l2[0][-1]                                                    # :25 Select(1,-1)
    q = dropout(ctx, q, cfg["dropout"], SITE_FUSION)                 # :29-31
    out = torch.tanh(linear(q, P["fusion.weight"], P["fusion.bias"]))  # :32-33
    return out, {"rnnLayers": [l1, l2]}


def encoder_lf_ques_im_hist(ctx: Ctx, cfg, P, inputs):
    """encoders/lf-ques-im-hist.lua:3-61; concat order [question | image | history] (:49-54)."""
    ques, img, hist = inputs["ques"], inputs["img"], inputs["hist"]
    x = lookup_table_mask_zero(P["wordEmbed.weight"], ques)          # :15-16
    l1, l2 = _two_layer_lstm(P, "ques", x)                           # :19-28
    xh = lookup_table_mask_zero(P["wordEmbed.weight"], hist)         # :31-36
    _, h2 = _two_layer_lstm(P, "hist", xh)                           # :37-47
    j = torch.cat([l2[0][-1], img, h2[0][-1]], 1)                    # :49-54 JoinTable(2)
    j = dropout(ctx, j, cfg["dropout"], SITE_FUSION)                 # :55-57
    out = torch.tanh(linear(j, P["fusion.weight"], P["fusion.bias"]))  # :58-59
    return out, {"rnnLayers": [l1, l2]}


def encoder_hrea_ques_im_hist(ctx: Ctx, cfg, P, inputs):
    """encoders/hrea-ques-im-hist.lua:7-140."""
    ques, img, hist = inputs["ques"], inputs["img"], inputs["hist"]
    R = cfg["maxQuesCount"]
    H = cfg["rnnHiddenSize"]
    W = lookup_table_mask_zero(P["wordEmbed.weight"], ques)          # :20-21  (Tq,N,E)
    ie = linear(dropout(ctx, img, 0.5, SITE_IMG_FC7),                # :45-50  Dropout(0.5)->Linear
                P["img.embed.weight"], P["img.embed.bias"])
    I = mask_time(ques, ie)                                          # :52-55
    xh = lookup_table_mask_zero(P["wordEmbed.weight"], hist)         # :24-31
    _, h2 = _two_layer_lstm(P, "hist", xh)                           # :32-41
    Hs = h2[0][-1]                                                   # (N,H)
    Qi_in = torch.cat([W, I], -1)                                    # :67-69 JoinTable(-1)
    l1, l2 = _two_layer_lstm(P, "ques", Qi_in)                       # :72-81
    Qi = l2[0][-1]                                                   # (N,H)
    # attention over history, :89-118
    sq = linear(Qi, P["att.q.weight"], P["att.q.bias"]).view(-1, R)  # (B,R)   [b,i]
    sh = linear(Hs, P["att.h.weight"], P["att.h.bias"]).view(-1, R)  # (B,R)   [b,j]
    score = sq.unsqueeze(2) + sh.unsqueeze(1)                        # Replicate(10,3)/(10,2), CAddTable
    score = mask_future(score)                                       # :101
    score = replace_zero(score.reshape(-1, R), -math.inf)            # :102-103
    prob = torch.softmax(score, -1).view(-1, R, R)                   # :104-105
    Hv = Hs.view(-1, R, H)
    # Replicate(512,4) * Replicate(10,2) -> CMulTable -> Sum(3)  (:106-127)
    att = (prob.unsqueeze(-1) * Hv.unsqueeze(1)).sum(2).reshape(-1, H)
    j = torch.cat([att, Qi], -1)                                     # :132 JoinTable(-1)
    j = j.view(-1, R, 2 * H).transpose(0, 1)                         # :133-134 (R,B,2H)
    d, _ = seq_lstm(j, *lstm_p(P, "dialog.lstm"), maskzero=False)    # :135
    out = d.transpose(0, 1).reshape(-1, H)                           # :136-137
    return out, {"rnnLayers": [l1, l2]}


def encoder_mn_att_ques_im_hist(ctx: Ctx, cfg, P, inputs):
    """encoders/mn-att-ques-im-hist.lua:5-115."""
    ques, img, hist, mask = inputs["ques"], inputs["img"], inputs["hist"], inputs["mask"]
    R = cfg["maxQuesCount"]
    H = cfg["rnnHiddenSize"]
    S2 = cfg["imgSpatialSize"] ** 2
    C = cfg["imgFeatureSize"]
    emb = P["wordEmbed.weight"]
    qE = dropout(ctx, lookup_table_mask_zero(emb, ques), 0.5, SITE_QEMBED)   # :24
    hE = dropout(ctx, lookup_table_mask_zero(emb, hist), 0.5, SITE_HEMBED)   # :25
    _, h2 = _two_layer_lstm(P, "hist", hE)                           # :27-34
    h3 = h2[0][-1]                                                   # :35
    _, q2 = _two_layer_lstm(P, "ques", qE)                           # :37-44
    q3 = q2[0][-1]                                                   # :45
    qV = q3.view(-1, R, H)                                           # :48
    hV = h3.view(-1, R, H)                                           # :49
    qh = torch.bmm(qV, hV.transpose(1, 2))                           # :54 MM(false,true)
    probs = mask_softmax(qh.reshape(-1, R), mask).view(-1, R, R)     # :55-57
    hAtt = torch.bmm(probs, hV).reshape(-1, H)                       # :61-62
    hAttTr = torch.tanh(linear(dropout(ctx, hAtt, 0.5, SITE_HATT),   # :64
                               P["mn.fact.weight"], P["mn.fact.bias"]))
    qh2 = torch.tanh(linear(hAttTr + q3, P["mn.query.weight"], P["mn.query.bias"]))  # :65
    # SAN, :67-106
    u = qh2
    N = q3.shape[0]
    img_tr = torch.tanh(linear(img.reshape(-1, C), P["san.img.weight"], P["san.img.bias"]))
    img_tr = dropout(ctx, img_tr.view(N, S2, H), 0.5, SITE_IMG_TR)   # :74-78
    for hop in range(cfg["numAttentionLayers"]):
        pre = "san.hop%d." % (hop + 1)
        img_common = linear(img_tr.reshape(-1, H), P[pre + "img_common.weight"],
                            P[pre + "img_common.bias"]).view(N, S2, -1)       # :83-85
        ques_common = linear(u, P[pre + "ques_common.weight"], P[pre + "ques_common.bias"])  # :88
        iq = torch.tanh(img_common + ques_common.unsqueeze(1))       # :89-92 Replicate + CAddTable
        iq = dropout(ctx, iq, 0.5, SITE_HOP0 + hop)
        s = linear(iq.reshape(-1, iq.shape[-1]), P[pre + "score.weight"], P[pre + "score.bias"])  # :93
        p = torch.softmax(s.view(N, S2), -1)                         # :94
        att = torch.bmm(p.unsqueeze(1), img_tr).reshape(N, H)        # :97-99
        u = att + u                                                  # :102
    out = torch.tanh(linear(dropout(ctx, u, 0.5, SITE_U_OUT), P["san.out.weight"], P["san.out.bias"]))  # :106
    return out, {"rnnLayers": None}


def encoder_lf_ques_im(ctx: Ctx, cfg, P, inputs):
    """encoders/lf-ques-im.lua:3-43: JoinTable(2) [question | image] -> Dropout -> Linear(H+F,H) -> Tanh."""
    x = lookup_table_mask_zero(P["wordEmbed.weight"], inputs["ques"])   # :15-16
    l1, l2 = _two_layer_lstm(P, "ques", x)                              # :19-28
    j = torch.cat([l2[0][-1], inputs["img"]], 1)                        # :31-35
    j = dropout(ctx, j, cfg["dropout"], SITE_FUSION)                    # :36-38
    out = torch.tanh(linear(j, P["fusion.weight"], P["fusion.bias"]))   # :39-40
    return out, {"rnnLayers": [l1, l2]}


def encoder_lf_ques_hist(ctx: Ctx, cfg, P, inputs):
    """encoders/lf-ques-hist.lua:3-62: JoinTable(2) [question | history] -> Dropout -> Linear(2H,H) -> Tanh."""
    x = lookup_table_mask_zero(P["wordEmbed.weight"], inputs["ques"])   # :15-16
    l1, l2 = _two_layer_lstm(P, "ques", x)                              # :19-28
    xh = lookup_table_mask_zero(P["wordEmbed.weight"], inputs["hist"])  # :31-36
    _, h2 = _two_layer_lstm(P, "hist", xh)                              # :37-47
    j = torch.cat([l2[0][-1], h2[0][-1]], 1)                            # :50-54
    j = dropout(ctx, j, cfg["dropout"], SITE_FUSION)                    # :55-57
    out = torch.tanh(linear(j, P["fusion.weight"], P["fusion.bias"]))   # :58-59
    return out, {"rnnLayers": [l1, l2]}


def _dialog_lstm(cfg, P, j):
    """View(-1,10,2H) -> Transpose(1,2) -> SeqLSTM(2H,H) (no maskZero) -> Transpose(1,2) -> View(-1,H)
    (hre-ques-hist.lua:63-68, hre-ques-im-hist.lua:90-94)."""
    R, H = cfg["maxQuesCount"], cfg["rnnHiddenSize"]
    j = j.view(-1, R, 2 * H).transpose(0, 1)
    d, _ = seq_lstm(j, *lstm_p(P, "dialog.lstm"), maskzero=False)
    return d.transpose(0, 1).reshape(-1, H)


def encoder_hre_ques_hist(ctx: Ctx, cfg, P, inputs):
    """encoders/hre-ques-hist.lua:3-71: [question | history] per round -> dialog-level SeqLSTM over the rounds."""
    x = lookup_table_mask_zero(P["wordEmbed.weight"], inputs["ques"])   # :16-17
    l1, l2 = _two_layer_lstm(P, "ques", x)                              # :20-29
    xh = lookup_table_mask_zero(P["wordEmbed.weight"], inputs["hist"])  # :33-38
    _, h2 = _two_layer_lstm(P, "hist", xh)                              # :39-50
    j = torch.cat([l2[0][-1], h2[0][-1]], -1)                           # :59 JoinTable(1,1)
    return _dialog_lstm(cfg, P, j), {"rnnLayers": [l1, l2]}             # :63-68


def encoder_hre_ques_im_hist(ctx: Ctx, cfg, P, inputs):
    """encoders/hre-ques-im-hist.lua:5-97: hrea-ques-im-hist without the history attention, and without the Dropout
    in front of the image Linear (commented out upstream, :46)."""
    ques, img, hist = inputs["ques"], inputs["img"], inputs["hist"]
    W = lookup_table_mask_zero(P["wordEmbed.weight"], ques)             # :18-19
    ie = linear(img, P["img.embed.weight"], P["img.embed.bias"])        # :43-48
    I = mask_time(ques, ie)                                             # :50-53
    xh = lookup_table_mask_zero(P["wordEmbed.weight"], hist)            # :23-28
    _, h2 = _two_layer_lstm(P, "hist", xh)                              # :29-39
    l1, l2 = _two_layer_lstm(P, "ques", torch.cat([W, I], -1))          # :65-79
    j = torch.cat([l2[0][-1], h2[0][-1]], -1)                           # :82-86 JoinTable(-1)
    return _dialog_lstm(cfg, P, j), {"rnnLayers": [l1, l2]}             # :90-94


def _memory_block(ctx, cfg, P, q, h3, mask):
    """MM(q, h^T) -> MaskSoftMax -> MM(probs, h) -> Dropout(0.5) -> Linear -> Tanh; + q; Linear -> Tanh
    (mn-ques-hist.lua:46-63, mn-ques-im-hist.lua:51-68)."""
    R, H = cfg["maxQuesCount"], cfg["rnnHiddenSize"]
    qV, hV = q.view(-1, R, H), h3.view(-1, R, H)
    qh = torch.bmm(qV, hV.transpose(1, 2))
    probs = mask_softmax(qh.reshape(-1, R), mask).view(-1, R, R)
    hAtt = torch.bmm(probs, hV).reshape(-1, H)
    hAttTr = torch.tanh(linear(dropout(ctx, hAtt, 0.5, SITE_HATT), P["mn.fact.weight"], P["mn.fact.bias"]))
    return torch.tanh(linear(hAttTr + q, P["mn.query.weight"], P["mn.query.bias"]))


def _embdrop_lstms(ctx, P, inputs):
    """Dropout(0.5) on both embedded sequences, two 2-layer maskZero LSTM stacks, last step of each
    (mn-ques-hist.lua:22-43, mn-ques-im-hist.lua:24-45, lf-att-ques-im-hist.lua:20-41)."""
    emb = P["wordEmbed.weight"]
    qE = dropout(ctx, lookup_table_mask_zero(emb, inputs["ques"]), 0.5, SITE_QEMBED)
    hE = dropout(ctx, lookup_table_mask_zero(emb, inputs["hist"]), 0.5, SITE_HEMBED)
    _, h2 = _two_layer_lstm(P, "hist", hE)
    _, q2 = _two_layer_lstm(P, "ques", qE)
    return q2[0][-1], h2[0][-1]


def encoder_mn_ques_hist(ctx: Ctx, cfg, P, inputs):
    """encoders/mn-ques-hist.lua:5-71."""
    q3, h3 = _embdrop_lstms(ctx, P, inputs)                             # :22-43
    return _memory_block(ctx, cfg, P, q3, h3, inputs["mask"]), {"rnnLayers": None}   # :46-63


def encoder_mn_ques_im_hist(ctx: Ctx, cfg, P, inputs):
    """encoders/mn-ques-im-hist.lua:5-76: the memory query (and the residual) is Tanh(Linear(F+H,H)([q | fc7])), :47-48."""
    q3, h3 = _embdrop_lstms(ctx, P, inputs)                             # :24-45
    qi = torch.tanh(linear(torch.cat([q3, inputs["img"]], 1), P["mn.qi.weight"], P["mn.qi.bias"]))   # :47-48
    return _memory_block(ctx, cfg, P, qi, h3, inputs["mask"]), {"rnnLayers": None}   # :51-68


def encoder_lf_att_ques_im_hist(ctx: Ctx, cfg, P, inputs):
    """encoders/lf-att-ques-im-hist.lua:3-93: qh = Tanh(Linear(2H,H)([q | h])) (:43), then ONE attention hop over pool5
    (`num_attention_layer = 1` is hard-wired upstream, :49 — params.numAttentionLayers is not read) and the output layer."""
    H, S2, C = cfg["rnnHiddenSize"], cfg["imgSpatialSize"] ** 2, cfg["imgFeatureSize"]
    q3, h3 = _embdrop_lstms(ctx, P, inputs)                             # :20-41
    u = torch.tanh(linear(torch.cat([q3, h3], 1), P["fusion.weight"], P["fusion.bias"]))   # :43
    N = u.shape[0]
    img_tr = torch.tanh(linear(inputs["img"].reshape(-1, C), P["san.img.weight"], P["san.img.bias"]))
    img_tr = dropout(ctx, img_tr.view(N, S2, H), 0.5, SITE_IMG_TR)      # :52-56
    pre = "san.hop1."
    img_common = linear(img_tr.reshape(-1, H), P[pre + "img_common.weight"], P[pre + "img_common.bias"]).view(N, S2, -1)  # :61-63
    ques_common = linear(u, P[pre + "ques_common.weight"], P[pre + "ques_common.bias"])    # :66
    iq = dropout(ctx, torch.tanh(img_common + ques_common.unsqueeze(1)), 0.5, SITE_HOP0)   # :67-70
    s = linear(iq.reshape(-1, iq.shape[-1]), P[pre + "score.weight"], P[pre + "score.bias"])   # :71
    p = torch.softmax(s.view(N, S2), -1)                                # :72
    u = torch.bmm(p.unsqueeze(1), img_tr).reshape(N, H) + u             # :75-80
    out = torch.tanh(linear(dropout(ctx, u, 0.5, SITE_U_OUT), P["san.out.weight"], P["san.out.bias"]))   # :84
    return out, {"rnnLayers": None}


ENCODERS = {
    "lf-ques-im": encoder_lf_ques_im,
    "lf-ques-hist": encoder_lf_ques_hist,
    "hre-ques-hist": encoder_hre_ques_hist,
    "hre-ques-im-hist": encoder_hre_ques_im_hist,
    "mn-ques-hist": encoder_mn_ques_hist,
    "mn-ques-im-hist": encoder_mn_ques_im_hist,
    "lf-att-ques-im-hist": encoder_lf_att_ques_im_hist,
    "lf-ques": encoder_lf_ques,
    "lf-ques-im-hist": encoder_lf_ques_im_hist,
    "hrea-ques-im-hist": encoder_hrea_ques_im_hist,
    "mn-att-ques-im-hist": encoder_mn_att_ques_im_hist,
}


# ----------------------------------------------------------------------------------------
# decoders
# ----------------------------------------------------------------------------------------

def decoder_disc(ctx: Ctx, cfg, P, options: Tensor, encOut: Tensor) -> Tensor:
    """decoders/disc.lua:3-32: 100 weight-sharing branches run sequentially by nn.Concat(2):
    Select(2,i) -> shared embed -> SeqLSTM(E->H, batchfirst, NO maskzero) -> Select(2,-1);
    then MM with encOut (N,H,1) -> scores (N,100).  options: (N,100,To) ids."""
    N, K, To = options.shape
    W, b = lstm_p(P, "opt.lstm")
    emb = P["wordEmbed.weight"]
    if ctx.structure == "reference":
        feats = []
        for k in range(K):                                           # :9-20
            x = lookup_table_mask_zero(emb, options[:, k, :])        # (N,To,E) batch-first
            o, _ = seq_lstm(x.transpose(0, 1), W, b, maskzero=False)
            feats.append(o[-1].unsqueeze(1))                         # Select(2,-1), Reshape(1,H)
        feat = torch.cat(feats, 1)                                   # (N,100,H)
    else:                                                            # "batched CPU" structure
        x = lookup_table_mask_zero(emb, options.reshape(N * K, To))
        o, _ = seq_lstm(x.transpose(0, 1), W, b, maskzero=False)
        feat = o[-1].view(N, K, -1)
    return torch.bmm(feat, encOut.unsqueeze(2)).squeeze(2)           # :22-29


def gen_forward_connect(state, encOut):
    """decoders/gen.lua:30-42: per-layer (h0,c0) for the decoder LSTMs."""
    H0 = [None, None]
    C0 = [None, None]
    rl = state.get("rnnLayers")
    if rl is not None:
        for ii in range(len(rl)):
            H0[ii] = rl[ii][0][-1]          # enc.rnnLayers[ii].output[seqLen]
            C0[ii] = rl[ii][1][-1]          # enc.rnnLayers[ii].cell[seqLen]
        H0[len(rl) - 1] = encOut            # :37-38
    else:
        H0[1] = encOut                      # :40
    return H0, C0


def decoder_gen(ctx: Ctx, cfg, P, answer_in: Tensor, H0, C0) -> Tensor:
    """decoders/gen.lua:3-27: shared embed -> 2x SeqLSTM (maskzero, seeded state) ->
    Sequencer(MaskZero(Linear(H,V))) -> Sequencer(MaskZero(LogSoftMax)).  answer_in (Ta,N).
    Returns log-probs (Ta,N,V); rows of pad steps are exactly zero."""
    x = lookup_table_mask_zero(P["wordEmbed.weight"], answer_in)     # :10-11
    o1, _ = seq_lstm(x, *lstm_p(P, "dec.lstm1"), h0=H0[0], c0=C0[0], maskzero=True)
    o2, _ = seq_lstm(o1, *lstm_p(P, "dec.lstm2"), h0=H0[1], c0=C0[1], maskzero=True)
    T, N, H = o2.shape
    flat = o2.reshape(T * N, H)
    keep = (flat.detach().abs().sum(1) != 0).to(flat.dtype).unsqueeze(1)   # MaskZero(.,1)
    logits = linear(flat, P["dec.out.weight"], P["dec.out.bias"]) * keep   # :23
    logp = torch.log_softmax(logits, 1) * keep                             # :24
    return logp.view(T, N, -1)


def gen_criterion(logp: Tensor, answer_out: Tensor) -> Tensor:
    """model.lua:33-36: SequencerCriterion(MaskZeroCriterion(ClassNLLCriterion(sum),1)): rows whose
    input (log-prob row) is all-zero are skipped; token id c (1-based) is class c."""
    T, N, V = logp.shape
    flat = logp.reshape(T * N, V)
    tgt = answer_out.reshape(T * N)
    keep = (flat.detach().abs().sum(1) != 0) & (tgt > 0)
    picked = flat.gather(1, (tgt.clamp(min=1) - 1).unsqueeze(1)).squeeze(1)
    return -(picked * keep.to(flat.dtype)).sum()


def compute_lhood(words: Tensor, logp: Tensor) -> Tensor:
    """utils.lua:86-102: gather the log-prob of every target token, zero where target == 0,
    sum over time -> (N)."""
    T, N, V = logp.shape
    idx = words.reshape(-1, 1)
    m = idx == 0
    lp = logp.reshape(-1, V).gather(1, (idx.clamp(min=1) - 1))
    lp = lp.masked_fill(m, 0.0).view(T, N)
    return lp.sum(0)


def compute_ranks(scores: Tensor, gt_pos: Optional[Tensor] = None) -> Tensor:
    """utils.lua:106-128: descending sort, inverse permutation = rank of every option (1 = best).
    The reference's TH sort is not stable; the tie rule pinned here (and in the CUDA kernel) is
    'lower index wins': rank[k] = 1 + #{j: s_j > s_k or (s_j == s_k and j < k)}."""
    order = torch.argsort(-scores.double(), dim=1, stable=True)
    ranks = torch.empty_like(order)
    ar = torch.arange(1, scores.shape[1] + 1).expand_as(order)
    ranks.scatter_(1, order, ar)
    if gt_pos is not None:
        ranks = ranks.gather(1, (gt_pos.view(-1, 1).long() - 1)).squeeze(1)
    return ranks


def process_ranks(ranks: Tensor) -> Dict[str, float]:
    """utils.lua:131-160 metric formulae (R@1/5/10, median, mean rank, MRR)."""
    r = ranks.double().view(-1)
    n = r.numel()
    return {"r1": float((r <= 1).sum()) / n, "r5": float((r <= 5).sum()) / n,
            "r10": float((r <= 10).sum()) / n, "medianR": float(r.median()),
            "meanR": float(r.mean()), "meanRR": float((1.0 / r).mean())}


# ----------------------------------------------------------------------------------------
# Model:forwardBackward / retrieveBatch / clamp + adam
# ----------------------------------------------------------------------------------------

def prepare_inputs(cfg, batch: Dict[str, Tensor]) -> Dict[str, Tensor]:
    """model.lua:252-294: time-major views, 10x repeated image, MN mask."""
    inputs = {}
    R = cfg["maxQuesCount"]
    q = batch["ques_fwd"]
    inputs["ques"] = q.reshape(-1, q.shape[2]).t()                    # :255-257
    enc = cfg["encoder"]
    if "im" in enc:                                                   # opts.lua:56
        im = batch["img_feat"]
        if "att" in enc:                                              # :262-265 (materialised repeat)
            S, C = cfg["imgSpatialSize"], cfg["imgFeatureSize"]
            im = im.view(-1, 1, S, S, C).repeat(1, R, 1, 1, 1).view(-1, S, S, C)
        else:                                                         # :267-269
            F = cfg["imgFeatureSize"]
            im = im.view(-1, 1, F).repeat(1, R, 1).view(-1, F)
        inputs["img"] = im
    if "hist" in enc:                                                 # opts.lua:55
        h = batch["hist"]
        inputs["hist"] = h.reshape(-1, h.shape[2]).t()                # :274-278
    if "mn" in enc:                                                   # :280-294
        m = torch.ones(R, R, dtype=torch.uint8)
        for i in range(R):
            for j in range(R):
                if j <= i:
                    m[i, j] = 0
        inputs["mask"] = m.repeat(batch["hist"].shape[0], 1)
    return inputs


def forward_backward(ctx: Ctx, cfg, P: Dict[str, Tensor], batch: Dict[str, Tensor],
                     only_forward: bool = False):
    """Model:forwardBackward (model.lua:249-342).  Returns dict(loss, encOut, decOut, grads)."""
    if not only_forward:
        P = {k: v.detach().clone().requires_grad_(True) for k, v in P.items()}
    inputs = prepare_inputs(cfg, batch)
    encOut, state = ENCODERS[cfg["encoder"]](ctx, cfg, P, inputs)     # :297
    out = {"encOut": encOut.detach()}
    if cfg["decoder"] == "gen":
        H0, C0 = gen_forward_connect(state, encOut)                   # :300
        a_in = batch["answer_in"]
        a_in = a_in.reshape(-1, a_in.shape[2]).t()                    # :307-308
        a_out = batch["answer_out"]
        a_out = a_out.reshape(-1, a_out.shape[2]).t()                 # :310-311
        logp = decoder_gen(ctx, cfg, P, a_in, H0, C0)                 # :313
        loss = gen_criterion(logp, a_out)                             # :314
        out["decOut"] = logp.detach()
    else:
        scores = decoder_disc(ctx, cfg, P, batch["options"], encOut)  # :329
        tgt = batch["answer_ind"].reshape(-1).long() - 1
        loss = torch.nn.functional.cross_entropy(scores, tgt)         # :330 CrossEntropyCriterion (mean)
        out["decOut"] = scores.detach()
    out["loss"] = float(loss.detach())
    if not only_forward:
        loss.backward()                                               # :316-338
        out["grads"] = {k: (v.grad if v.grad is not None else torch.zeros_like(v)) for k, v in P.items()}
    return out


def retrieve_batch(ctx: Ctx, cfg, P, batch, use_gt: bool = True) -> Tensor:
    """Model:retrieveBatch (model.lua:344-430)."""
    with torch.no_grad():
        inputs = prepare_inputs(cfg, batch)
        encOut, state = ENCODERS[cfg["encoder"]](ctx, cfg, P, inputs)  # :389
        gt = batch["answer_ind"] if use_gt else None
        if cfg["decoder"] == "gen":
            oi = batch["option_in"]
            oo = batch["option_out"]
            oi = oi.reshape(-1, oi.shape[2], oi.shape[3]).transpose(0, 1).transpose(1, 2)  # (100,To,N)
            oo = oo.reshape(-1, oo.shape[2], oo.shape[3]).transpose(0, 1).transpose(1, 2)
            lh = []
            for k in range(oi.shape[0]):                              # :405-415
                H0, C0 = gen_forward_connect(state, encOut)
                logp = decoder_gen(ctx, cfg, P, oi[k], H0, C0)
                lh.append(compute_lhood(oo[k], logp))
            return compute_ranks(torch.stack(lh).t(), gt)             # :420
        scores = decoder_disc(ctx, cfg, P, batch["options"], encOut)  # :423
        return compute_ranks(scores, gt)                              # :427


def decoder_gen_step(cfg, P, tokens: Tensor, H, C):
    """One time step of decoders/gen.lua:3-27 with explicit state, as Model:generateAnswers drives it through
    `.userPrevOutput / .userPrevCell` (model.lua:517-524): tokens (N) -> (log-probs (N,V), [h1,h2], [c1,c2])."""
    x = lookup_table_mask_zero(P["wordEmbed.weight"], tokens.view(1, -1))
    o1, c1 = seq_lstm(x, *lstm_p(P, "dec.lstm1"), h0=H[0], c0=C[0], maskzero=True, manual_bptt=False)
    o2, c2 = seq_lstm(o1, *lstm_p(P, "dec.lstm2"), h0=H[1], c0=C[1], maskzero=True, manual_bptt=False)
    flat = o2[0]
    keep = (flat.abs().sum(1) != 0).to(flat.dtype).unsqueeze(1)
    logits = linear(flat, P["dec.out.weight"], P["dec.out.bias"]) * keep
    return torch.log_softmax(logits, 1) * keep, [o1[0], o2[0]], [c1[0], c2[0]]


def _initial_beam_state(state, encOut: Tensor, it: int, beam: int):
    """model.lua:478-503: decoder state of round `it` replicated over the beams."""
    Hd = encOut.shape[1]
    rl = state.get("rnnLayers")
    if rl is not None:                                                # encoders that expose .rnnLayers
        H = [rl[l][0][-1][it] for l in range(len(rl))]                # :482 output[Tq][iter]
        Cc = [rl[l][1][-1][it] for l in range(len(rl))]               # :483 cell[Tq][iter]
        H[len(rl) - 1] = encOut[it]                                   # :484-486
    else:                                                             # :491-501
        H = [encOut.new_zeros(Hd), encOut[it]]
        Cc = [encOut.new_zeros(Hd), encOut.new_zeros(Hd)]
    return [h.repeat(beam, 1) for h in H], [c.repeat(beam, 1) for c in Cc]


def generate_answers(ctx: Ctx, cfg, P, batch, start_token: int, end_token: int, beam_size: int = 5, beam_len: int = 20,
                     sample_words: bool = False, temperature: float = 1.0, generator: Optional[torch.Generator] = None,
                     strict: bool = True):
    """Model:generateAnswers for ONE dialog (model.lua:432-613; the reference loops convId = 1..numThreads with a batch
    of one dialog, :462-465).  Beam search (:472-579) restated literally, quirks included: step 2 expands one beam only
    (:510), finished hypotheses leave the pool with their un-normalised score (:546-547), columns beyond the number of
    surviving candidates keep their previous content (:559), the answer is the best FINISHED beam (:575-579; the
    reference indexes nil — raises here — when none finished).  Sampling (:581-602): multinomial over
    exp(logp / T) with the decoder fed its own samples (`decoderConnect`, gen.lua:63-68).
    Returns a list of 10 dicts {answer: LongTensor (beam_len) zero-padded, score, length} (beam) or
    {answer: LongTensor (beam_len + 1)} (sampling)."""
    assert cfg["decoder"] == "gen", "Sampling/beam search only for generative model"        # :434-437
    with torch.no_grad():
        inputs = prepare_inputs(cfg, batch)
        encOut, state = ENCODERS[cfg["encoder"]](Ctx(train=False), cfg, P, inputs)           # :468 (evaluate mode)
        R = encOut.shape[0]
        out = []
        if not sample_words:
            for it in range(R):                                                              # :474
                beams = torch.zeros(beam_len, beam_size, dtype=torch.long)                   # :479
                H, Cc = _initial_beam_state(state, encOut, it, beam_size)
                beams[0] = start_token                                                       # :506
                scores = torch.zeros(beam_size, dtype=torch.float64)                         # :507
                finished = []                                                                # :508
                for step in range(1, beam_len):                                              # :510 (Lua step = step + 1)
                    cands = []
                    explore = 1 if step == 1 else beam_size                                  # :516
                    logp, nH, nC = decoder_gen_step(cfg, P, beams[step - 1], H, Cc)          # :519-526
                    for w in range(explore):                                                 # :529
                        # :538-542 torch.topk(beamSize): sorted descending; ties (an all-zero MaskZero'd row, duplicate
                        # beams) are resolved 'lower class index first' — TH's order among ties is unspecified, the rule is
                        # pinned here and in visdial_b200/model.py so that both walk the same hypotheses
                        top_i = torch.argsort(-logp[w], stable=True)[:beam_size]
                        top_p = logp[w][top_i]
                        for cnd in range(beam_size):                                         # :544
                            cb = beams[:, w].clone()
                            tok = int(top_i[cnd]) + 1                                        # class index -> 1-based token id
                            cb[step] = tok
                            sc = float(scores[w]) + float(top_p[cnd])
                            if tok == end_token:                                             # :548-549
                                finished.append({"answer": cb, "length": step + 1, "score": sc})
                            else:                                                            # :550-553
                                cands.append((sc, cb, [h[w].clone() for h in nH], [c[w].clone() for c in nC]))
                    cands.sort(key=lambda t: -t[0])                                          # :558 (ties: stable here)
                    for k in range(min(len(cands), beam_size)):                              # :560-569
                        beams[:, k] = cands[k][1]
                        for lv in range(2):
                            H[lv][k] = cands[k][2][lv]
                            Cc[lv][k] = cands[k][3][lv]
                        scores[k] = cands[k][0]
                finished.sort(key=lambda d: -d["score"])                                     # :572
                if not finished:
                    if strict:
                        raise IndexError("no beam reached <END> within beamLen (model.lua:575 indexes nil here)")
                    out.append(None)
                    continue
                out.append(finished[0])                                                      # :575
        else:
            H0, C0 = gen_forward_connect(state, encOut)                                      # forwardBackward(batch, true, true)
            H = [h if h is not None else encOut.new_zeros(R, encOut.shape[1]) for h in H0]
            Cc = [c if c is not None else encOut.new_zeros(R, encOut.shape[1]) for c in C0]
            tok = torch.full((R,), start_token, dtype=torch.long)                            # :582
            seq = [tok.clone()]
            for _ in range(beam_len):                                                        # :584
                logp, H, Cc = decoder_gen_step(cfg, P, tok, H, Cc)                           # :586-588
                probs = torch.exp(logp / temperature)                                        # :590
                tok = torch.multinomial(probs, 1, generator=generator).squeeze(1) + 1
                seq.append(tok.clone())
            ans = torch.stack(seq, 1)                                                        # :594
            out = [{"answer": ans[i]} for i in range(R)]
    return out


def clamp_adam(W: Tensor, dW: Tensor, state: dict, lr: float, beta1=0.9, beta2=0.999, eps=1e-8):
    """model.lua:96-99 + model_utils/optim_updates.lua:62-91 on the flat vectors (in place)."""
    dW.clamp_(-5.0, 5.0)
    if "m" not in state:
        state["t"] = 0
        state["m"] = torch.zeros_like(dW)
        state["v"] = torch.zeros_like(dW)
    state["m"].mul_(beta1).add_(dW, alpha=1 - beta1)
    state["v"].mul_(beta2).addcmul_(dW, dW, value=1 - beta2)
    tmp = state["v"].sqrt().add_(eps)
    state["t"] += 1
    bc1 = 1 - beta1 ** state["t"]
    bc2 = 1 - beta2 ** state["t"]
    step = lr * math.sqrt(bc2) / bc1
    W.addcdiv_(state["m"], tmp, value=-step)
    return W


def decay_lr(lr: float, cfg) -> float:
    """model.lua:102-105."""
    if lr > cfg["minLRate"]:
        lr = lr * cfg["lrDecayRate"]
    return lr
