"""TEST INFRASTRUCTURE ONLY — CPU oracle for the visdial hot path.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference
legs may import this package.  The product (visdial_b200/) never does.

PARITY UNPINNED: the reference (/root/reference) ships no tests, golden vectors or
fixtures, and its arithmetic lives in un-vendored, un-pinned Lua rocks (torch7 nn,
nngraph, Element-Research/rnn) that cannot run in this image (no lua/luajit/th).
The oracle restates the reference graphs op-for-op (file:line cited per function)
and pins itself against maths instead: fp64 finite differences, closed-form
mini-cases and the structural invariants of SURVEY.md §8c (tests/test_oracle_*.py).
"""
