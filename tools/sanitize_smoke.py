#!/usr/bin/env python
"""Small end-to-end pass over every kernel family, for compute-sanitizer --tool memcheck (one tool per call)."""
import os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
os.environ.setdefault("YUE_SGD_HOT_MIN_COUNT", "1")          # force the hot-row table on a small log
os.environ.setdefault("YUE_SGD_MIN_EVENTS_PER_WARP", "512")  # several warps
from yue_b200 import synth
from yue_b200.engine import MODE_HOGWILD, MODE_SERIAL, RANK_EXACT, RANK_TC, Engine
eng = Engine(0)
log = synth.power_law_log(700, 500, 40000, seed=3)
for d in (32, 64, 128, 10):
    P, Q = synth.init_factors(log.m, log.n, d, seed=d)
    eng.set_interactions(log.m, log.n, log.ev_indptr, log.ev_items, log.uq_indptr, log.uq_items)
    eng.set_factors(P, Q)
    l0 = eng.bpr_epoch(0.05, 0.01, 0.01, 1, 0, MODE_HOGWILD)
    l1 = sum(eng.bpr_epoch_part(0.05, 0.01, 0.01, 1, 1, k, 3, MODE_HOGWILD) for k in range(3))
    l2 = eng.apr_epoch(0.003, 0.002, 0.01, 0.5, 2.0, 1, 2, 0, MODE_HOGWILD)
    l3 = eng.bpr_epoch(0.05, 0.01, 0.01, 1, 3, MODE_SERIAL) if d == 10 else 0.0
    print("d=%d losses %.1f %.1f %.1f %.1f" % (d, l0, l1, l2, l3), flush=True)
    users = log.test_users()[:300]
    eng.set_test_set(log.test_indptr, log.test_items)
    for algo, N in ((RANK_EXACT, 10), (RANK_EXACT, 50)) + (((RANK_TC, 10), (RANK_TC, 20)) if d <= 64 else ()):
        ids, sc = eng.rank_topn(users, N, algo)
        sums, distinct = eng.rank_metrics([5, N])
    print("  ranked, hits@%d = %d" % (N, sums[1, 0]), flush=True)
rng = np.random.default_rng(1)
u = rng.integers(0, 300, 20000).astype(np.int32); it = np.minimum(rng.zipf(1.3, 20000) - 1, 399).astype(np.int32)
eng.ingest_events(300, 400, u, it, (rng.random(20000) < 0.2).astype(np.uint8))
print("ingest sizes", eng.interaction_sizes(), flush=True)
eng.close()
print("sanitize smoke done")
