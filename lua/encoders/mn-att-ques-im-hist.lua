-- Drop-in for encoders/mn-att-ques-im-hist.lua (same file name and table shape; model.lua:19-20 loads it with dofile()).
return require('module_b200').encoder('mn-att-ques-im-hist')
