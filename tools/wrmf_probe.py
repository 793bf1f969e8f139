#!/usr/bin/env python
"""Time the WRMF half-sweeps (K7) on a synthetic power-law log; the ncu target for wrmf_solve_kernel.
usage: python tools/wrmf_probe.py [users tracks plays d [iterations]]   (default: a quarter of config C2's shape)"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from yue_b200 import synth  # noqa: E402
from yue_b200.engine import Engine  # noqa: E402

a = [int(x) for x in sys.argv[1:]]
users, tracks, plays, d = (a + [250000, 50000, 12500000, 64][len(a):])[:4]
iters = a[4] if len(a) > 4 else 2
log = synth.power_law_log_torch(users, tracks, plays, 20260107, device="cuda")
P, Q = synth.init_factors(log.m, log.n, d, 20261107)
eng = Engine(0)
eng.set_interactions(log.m, log.n, log.ev_indptr, log.ev_items, log.uq_indptr, log.uq_items)
eng.set_factors(P * 10, Q * 10)
nnz = eng.interaction_sizes()[3]
print("users %d tracks %d unique pairs %d d %d" % (log.m, log.n, nnz, d), flush=True)
for it in range(iters):
    for side, name in ((0, "user"), (1, "track")):
        eng.sync()
        eng.timer_start()
        loss = eng.wrmf_sweep(side, 1.0, 10.0, want_loss=(side == 0 and bool(os.environ.get("PROBE_LOSS"))))
        print("iteration %d %-5s sweep %8.3f ms%s" % (it, name, eng.timer_stop(), "" if loss is None else "  loss %.6g" % loss), flush=True)
