-- Drop-in for /root/reference/encoders/lf-ques-hist.lua: same table shape ({model = ...}), same file name, loaded by
-- model.lua:19-20 with dofile().  The graph itself runs in libvisdial_b200.so (csrc/engine.cu).
local mod = require 'module_b200'
local encoderNet = {}

function encoderNet.model(params)
  assert(params.encoder == 'lf-ques-hist')
  local enc = mod.newHalf('enc', params, 'lf-ques-hist')
  enc.wordEmbed = 'wordEmbed.weight'            -- shared table lives in the engine (disc.lua:12, gen.lua:10)
  enc.rnnLayers = {'ques.lstm1', 'ques.lstm2'}   -- gen.lua:30-42 reads this
  return enc
end

return encoderNet
