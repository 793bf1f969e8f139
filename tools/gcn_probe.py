#!/usr/bin/env python
"""Time K9 (LightGCN steps, csrc/lightgcn.cuh) at config/LightGCN.conf's settings on config C1's shape; the ncu target.
usage: python tools/gcn_probe.py [steps]   (YUE_GCN_TIMING=1 prints the per-phase times of CTA 0)"""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from yue_b200 import synth  # noqa: E402
from yue_b200.engine import Engine  # noqa: E402
from yue_b200.lightgcn import truncated_normal  # noqa: E402

steps = int(sys.argv[1]) if len(sys.argv) > 1 else 0
users, tracks, plays = (int(x) for x in os.environ.get("PROBE_SIZE", "4000,50000,100000").split(","))
log = synth.power_law_log(users, tracks, plays, 20260142, test_ratio=0.2)
ev_user = np.repeat(np.arange(log.m, dtype=np.int32), np.diff(log.ev_indptr))
perm = np.random.default_rng(7).permutation(len(ev_user))
rng = np.random.default_rng(8)
eng = Engine(0)
eng.set_interactions(log.m, log.n, log.ev_indptr, log.ev_items, log.uq_indptr, log.uq_items)
eng.set_factors(truncated_normal((log.m, 50), 0.005, rng), truncated_normal((log.n, 50), 0.005, rng))
eng.gcn_set_events(ev_user[perm], log.ev_items[perm])
n = (log.train_size + 127) // 128
end = min(n, steps) if steps else n
eng.gcn_epoch(128, 0.002, 0.001, 1, 0, step_end=min(end, 20))
for ep in range(1, 4):
    eng.sync()
    eng.timer_start()
    loss = eng.gcn_epoch(128, 0.002, 0.001, 1, ep, step_end=end)
    ms = eng.timer_stop()
    print("pass %d: %d steps in %.2f ms = %.1f us per step, loss %.4f -> %.4f" % (ep, end, ms, 1e3 * ms / end, loss[0], loss[-1]), flush=True)
