-- Drop-in for /root/reference/decoders/disc.lua (table shape of disc.lua:3-38).
local mod = require 'module_b200'
local decoderNet = {}

function decoderNet.model(params, enc)
  return mod.newHalf('dec', params, 'disc')
end
function decoderNet.forwardConnect(enc, dec, encOut, seqLen) end   -- disc.lua:35
function decoderNet.backwardConnect(enc, dec) end                  -- disc.lua:38

return decoderNet
