-- Drop-in for encoders/lf-ques-hist.lua (same file name and table shape; model.lua:19-20 loads it with dofile()).
return require('module_b200').encoder('lf-ques-hist')
