-- Drop-in for /root/reference/decoders/gen.lua (table shape of gen.lua:3-68).
local ffi = require 'ffi'
local vd = require 'visdial_ffi'
local mod = require 'module_b200'
local decoderNet = {}

function decoderNet.model(params, enc)
  local dec = mod.newHalf('dec', params, 'gen')
  dec.rnnLayers = {'dec.lstm1', 'dec.lstm2'}
  return dec
end
function decoderNet.forwardConnect(enc, dec, encOut, seqLen)       -- gen.lua:30-42
  vd.check(vd.C.vd_forward_connect(dec.engine))
end
function decoderNet.backwardConnect(enc, dec)                      -- gen.lua:45-60 -> gradient wrt encOut
  local g = ffi.new('const float*[1]')
  vd.check(vd.C.vd_backward_connect(dec.engine, g))
  return g[0]
end
function decoderNet.decoderConnect(dec)                            -- gen.lua:63-68
  -- the reference copies each layer's last output / cell into userPrevOutput / userPrevCell; vd_gen_decoder_step takes the
  -- previous (h, c) explicitly (the device pointers vd_gen_decoder_step returned for the previous token), nothing to copy
end

return decoderNet
