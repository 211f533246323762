-- nn.Module-protocol objects backed by the C engine: what encoders/<name>.lua and decoders/<name>.lua return
-- instead of an nn / nngraph module (model.lua:25-26, :42-57).  AUTHORED, NOT EXECUTED (no Lua runtime here).
local ffi = require 'ffi'
local vd = require 'visdial_ffi'

local Half = {}
Half.__index = Half

local function newHalf(kind, params, name)
  return setmetatable({kind = kind, params = params, name = name}, Half)
end

-- `cbatch` = the vd_batch cdata (`visdial_ffi.batch(b).c`); the caller keeps the table visdial_ffi.batch returned alive
-- for the whole step (it anchors the converted host tensors the struct points into)
function Half:forward(cbatch)                       -- encoder:forward(inputs) / decoder:forward(x), model.lua:297,313,329
  local out = ffi.new('const float*[1]')
  if self.kind == 'enc' then vd.check(vd.C.vd_encoder_forward(self.engine, cbatch, out))
  else vd.check(vd.C.vd_decoder_forward(self.engine, cbatch, out)) end
  self.output = out[0]
  return self.output
end

function Half:backward(cbatch, grad)                -- model.lua:319,323,335,337
  if self.kind == 'enc' then
    vd.check(vd.C.vd_encoder_backward(self.engine, cbatch, grad))
  else
    vd.check(vd.C.vd_decoder_backward(self.engine, cbatch))
    local g = ffi.new('const float*[1]')
    vd.check(vd.C.vd_backward_connect(self.engine, g))
    return {nil, g[0]}                              -- disc: {gradOptions, gradEncOut}; the caller uses [2]
  end
end

-- nn.Sequential():add(enc):add(dec) (model.lua:42): creates the engine = the flat parameter vector
local Wrapper = {}
Wrapper.__index = Wrapper

function Wrapper.new(enc, dec, params)
  local h = ffi.new('vd_engine*[1]')
  vd.check(vd.C.vd_create(vd.params(params), h))
  local self = setmetatable({engine = ffi.gc(h[0], vd.C.vd_destroy), enc = enc, dec = dec}, Wrapper)
  enc.engine, dec.engine = self.engine, self.engine
  return self
end
function Wrapper:cuda() return self end             -- the engine already lives on params.gpuid
function Wrapper:get(i) return i == 1 and self.enc or self.dec end
function Wrapper:getParameters()                    -- model.lua:55: flat device buffers
  local W, dW = ffi.new('float*[1]'), ffi.new('float*[1]')
  vd.check(vd.C.vd_param_buffers(self.engine, W, dW))
  return W[0], dW[0]
end
function Wrapper:training() vd.check(vd.C.vd_set_training(self.engine, 1)) end
function Wrapper:evaluate() vd.check(vd.C.vd_set_training(self.engine, 0)) end
function Wrapper:zeroGradParameters() vd.check(vd.C.vd_zero_grad(self.engine)) end

-- What `dofile('encoders/<name>.lua')` returns (model.lua:19-20): {model = function(params) ... end}.  The graph runs in the engine;
-- the only per-encoder fact left on this side is whether upstream exports enc.rnnLayers (Sequential encoders do, the nngraph
-- gModule ones — mn-*, lf-att-* — do not), which decoders/gen.lua:30-42 reads to decide on the state copy.
local exportsRnnLayers = {
  ['lf-ques'] = true, ['lf-ques-im'] = true, ['lf-ques-hist'] = true, ['lf-ques-im-hist'] = true, ['lf-att-ques-im-hist'] = false,
  ['hre-ques-hist'] = true, ['hre-ques-im-hist'] = true, ['hrea-ques-im-hist'] = true,
  ['mn-ques-hist'] = false, ['mn-ques-im-hist'] = false, ['mn-att-ques-im-hist'] = false,
}

local function encoder(name)
  assert(exportsRnnLayers[name] ~= nil, 'unknown encoder ' .. name)
  return {model = function(params)
    assert(params.encoder == name)
    local enc = newHalf('enc', params, name)
    enc.wordEmbed = 'wordEmbed.weight'            -- the shared table lives in the engine (disc.lua:12, gen.lua:10)
    enc.rnnLayers = exportsRnnLayers[name] and {'ques.lstm1', 'ques.lstm2'} or nil
    return enc
  end}
end

return {newHalf = newHalf, Wrapper = Wrapper, encoder = encoder}
