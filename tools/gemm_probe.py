"""Tile-GEMM timing probe: full tiles vs partially filled tiles (block skipping), through gadj_test_gemm."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
from dynadjust_b200 import engine
adj = engine.Adjustment()
rng = np.random.default_rng(1)
def run(M, N, K, reps=5):
    A = rng.standard_normal((M, K)); B = rng.standard_normal((N, K))
    C, ms = adj.test_gemm(A, B, reps=reps)
    ms /= reps
    err = np.abs(C - A @ B.T).max()
    tiles = -(-M // 128) * -(-N // 128)
    print(f"M={M:6d} N={N:5d} K={K:5d} tiles={tiles:5d} ms={ms:8.4f} us/tile/SM={ms*1e3/max(1,-(-tiles//148)):8.2f} TF/s={2.0*M*N*K/ms/1e9:6.2f} err={err:.1e}", flush=True)
for N in (128, 64, 32, 16):
    run(148 * 128, N, 1024)
for M in (128, 64, 62, 16):
    run(M, 148 * 128, 1024)
run(148 * 128, 128, 240)
run(148 * 128, 32, 240)
run(148 * 128, 128, 80)
run(148 * 128, 32, 80)
run(4096, 4096, 4096, reps=3)
