-- Device-resident replacement for the tensor preparation and batch indexing of dataloader.lua (:143-478).
-- AUTHORED, NOT EXECUTED (no Lua runtime in this image).  Usage inside dataloader:initialize, after the
-- quesFile:read / imgFile:read calls of dataloader.lua:45-129 and INSTEAD of self:prepareDataset(dtype) (:135):
--
--     local b200 = require 'dataloader_b200'
--     self.corpus = self.corpus or {}
--     self.corpus[dtype] = b200.create(model.wrapper.engine, self, dtype, opt)
--
-- and in getTrainBatch / getTestBatch (:324-375), instead of getIndexData + getIndexOption:
--
--     return b200.getBatch(self.corpus[dtype], inds, params.decoder, isTest)     -- a vd_batch of device pointers
local ffi = require 'ffi'
local vd = require 'visdial_ffi'

local M = {}

-- the h5 tensors arrive as Float/Double/Long tensors; the C ABI takes contiguous int32 / float32
local function i32(t) return t:int():contiguous() end

function M.create(engine, dl, dtype, opt)
  local keep = {                                             -- referenced until vd_corpus_create has copied them
    ques = i32(dl[dtype .. '_ques']), ques_len = i32(dl[dtype .. '_ques_len']),
    ans = i32(dl[dtype .. '_ans']), ans_len = i32(dl[dtype .. '_ans_len']),
    opt = i32(dl[dtype .. '_opt']), opt_list = i32(dl[dtype .. '_opt_list']), opt_len = i32(dl[dtype .. '_opt_len']),
  }
  local d = ffi.new('vd_corpus_desc')
  d.numThreads, d.numRounds, d.maxQuesLen = keep.ques:size(1), keep.ques:size(2), keep.ques:size(3)
  d.maxAnsLen = keep.ans:size(3)
  d.numOptions, d.numOptList = keep.opt:size(3), keep.opt_list:size(1)
  d.useHistory = opt.useHistory and 1 or 0
  d.concatHistory = opt.concatHistory and 1 or 0
  d.useIm = opt.useIm and 1 or 0
  d.maxHistoryLen = opt.maxHistoryLen or 60                  -- dataloader.lua:142
  d.imgNorm = opt.imgNorm or 0
  d.imgAtt = string.match(opt.encoder, 'att') and 1 or 0     -- :70
  d.startToken, d.endToken = dl.word2ind['<START>'], dl.word2ind['<END>']
  d.ques, d.ques_len, d.ans, d.ans_len = keep.ques:data(), keep.ques_len:data(), keep.ans:data(), keep.ans_len:data()
  d.opt, d.opt_list, d.opt_len = keep.opt:data(), keep.opt_list:data(), keep.opt_len:data()
  if dtype ~= 'test' then
    keep.ans_index = i32(dl[dtype .. '_ans_ind']); d.ans_index = keep.ans_index:data()
  end
  if opt.useHistory then
    keep.cap, keep.cap_len = i32(dl[dtype .. '_cap']), i32(dl[dtype .. '_cap_len'])
    d.maxCapLen = keep.cap:size(2)
    d.cap, d.cap_len = keep.cap:data(), keep.cap_len:data()
  end
  if opt.useIm then
    -- pass the features AS READ from the h5 file (before the norm / permute of :64-72: the library does both) and the
    -- 0-based img_pos (before the +1 of :77)
    keep.images = dl[dtype .. '_img_raw']:float():contiguous()
    keep.img_pos = i32(dl[dtype .. '_img_pos'] - 1)
    d.numImages, d.imgChannels = keep.images:size(1), keep.images:size(2)
    d.imgSpatial = d.imgAtt == 1 and keep.images:size(3) or 0
    d.images, d.img_pos = keep.images:data(), keep.img_pos:data()
  end
  local h = ffi.new('vd_corpus*[1]')
  vd.check(vd.C.vd_corpus_create(engine, d, h))
  return {h = ffi.gc(h[0], vd.C.vd_corpus_destroy), num_rounds = dl[dtype .. '_num_rounds']}
end

-- inds: 1-based LongTensor as drawn by getTrainBatch (:326) or filled by getTestBatch (:356-357)
function M.getBatch(corpus, inds, decoder, isTest)
  local n = inds:size(1)
  local inds0 = ffi.new('int64_t[?]', n)
  for i = 1, n do inds0[i - 1] = inds[i] - 1 end
  local mode = (decoder == 'disc') and 0 or (isTest and 2 or 1)
  local batch = ffi.new('vd_batch')
  vd.check(vd.C.vd_corpus_get_batch(corpus.h, inds0, n, mode, batch))
  local out = {cbatch = batch}
  if isTest and corpus.num_rounds then out.num_rounds = corpus.num_rounds:index(1, inds):long() end   -- :373
  return out
end

return M
