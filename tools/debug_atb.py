"""GPU debug helper (not a test): structured inputs through vd_gemm_atb in TF32 mode."""
import ctypes as C, sys, os
import numpy as np
sys.path.insert(0, os.path.join(os.path.dirname(__file__), "..")); sys.path.insert(0, os.path.join(os.path.dirname(__file__), "..", "tests"))
from helpers import small_params
from visdial_b200 import Engine, VD_MATH_TF32
from visdial_b200._lib import check
from test_tensorcore_gpu import Dev

eng = Engine(small_params("lf-ques", "disc"))
eng.set_math_mode(VD_MATH_TF32)

def run(A, B, C0):
    K, M = A.shape; N = B.shape[1]
    dA, dB, dC = Dev(eng, A), Dev(eng, B), Dev(eng, C0)
    check(eng.lib.vd_gemm_atb(eng.h, M, N, K, dA.p, M, dB.p, N, dC.p, N))
    eng.synchronize()
    return dC.get()

M, N, K = 128, 256, 64
A = np.ones((K, M), np.float32); B = np.ones((K, N), np.float32)
got = run(A, B, np.zeros((M, N), np.float32))
print("ones: unique", np.unique(got)[:10], "expect", K)
A = np.zeros((K, M), np.float32); A[:, :] = np.arange(M)[None, :]
got = run(A, B, np.zeros((M, N), np.float32))
print("A=m index: got[:8,0]/K", got[:8, 0] / K, "got[32:36,0]/K", got[32:36, 0] / K, "row const?", np.abs(got - got[:, :1]).max())
A = np.ones((K, M), np.float32); B = np.zeros((K, N), np.float32); B[:, :] = np.arange(N)[None, :]
got = run(A, B, np.zeros((M, N), np.float32))
print("B=n index: got[0,:8]/K", got[0, :8] / K, "got[0,32:36]/K", got[0, 32:36] / K, got[0, 250:256] / K)
A = np.zeros((K, M), np.float32); A[3, :] = 1; B = np.zeros((K, N), np.float32); B[:, :] = np.arange(K)[:, None]
got = run(A, B, np.zeros((M, N), np.float32))
print("A one-hot k=3, B=k index: expect 3 ->", np.unique(got)[:8])
rng = np.random.default_rng(0)
A = rng.standard_normal((K, M)).astype(np.float32); B = rng.standard_normal((K, N)).astype(np.float32)
got = run(A, B, np.zeros((M, N), np.float32)); ref = A.astype(np.float64).T @ B
print("random: max err", np.abs(got - ref).max(), "got[0,:4]", got[0, :4], "ref[0,:4]", ref[0, :4])
