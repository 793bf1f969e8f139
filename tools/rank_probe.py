#!/usr/bin/env python
"""Time (and let ncu profile) the tcgen05 ranking kernel on a C4-shaped block: B users x 2 M tracks."""
import sys, time, os
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from yue_b200 import synth
from yue_b200.engine import Engine, RANK_TC
B = int(sys.argv[1]) if len(sys.argv) > 1 else 18944
n = int(sys.argv[2]) if len(sys.argv) > 2 else 2_000_000
reps = int(sys.argv[3]) if len(sys.argv) > 3 else 3
N = int(sys.argv[4]) if len(sys.argv) > 4 else 10
D = int(sys.argv[5]) if len(sys.argv) > 5 else 64
eng = Engine(0)
indptr, uq = synth.mask_csr_torch(B, n, 50, 7)
P, Q = synth.init_factors(B, n, D, 8)
eng.set_interactions(B, n, np.zeros(B + 1, np.int64), np.zeros(0, np.int32), indptr, uq)
eng.set_factors(P, Q)
users = np.arange(B, dtype=np.int32)
eng.rank_topn(users[:256], N, RANK_TC)
for rep in range(reps):
    t0 = time.perf_counter(); eng.rank_topn(users, N, RANK_TC); dt = time.perf_counter() - t0
    print("N=%d " % N + "TC B=%d n=%d  %.4f s  %.0f users/s  %.1f dense TFLOP/s  (fallback rows, spilled rows) = %s" % (B, n, dt, B / dt, 2.0 * B * n * D / dt / 1e12, eng.rank_stats()), flush=True)
