#!/usr/bin/env python
"""Time one Hogwild epoch of CUNE's two-level BPR (K8, yue_cune_epoch) on a synthetic power-law log; the ncu target for
cune_sgd_kernel.  Implicit positives are synthetic: for every user, the unplayed tracks of two other users (the shape
CUNE.py:111-113 produces), none for every fifth user.
usage: python tools/cune_probe.py [--small] [users tracks plays d [epochs]]   (default: a tenth of config C2's shape)"""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from yue_b200 import synth  # noqa: E402
from yue_b200.cune import implicit_positive_lists  # noqa: E402
from yue_b200.engine import MODE_HOGWILD, Engine  # noqa: E402

argv = [x for x in sys.argv[1:] if x != "--small"]
a = [int(x) for x in argv]
default = [20000, 8000, 500000, 64] if "--small" in sys.argv else [100000, 20000, 5000000, 64]
users, tracks, plays, d = (a + default[len(a):])[:4]
epochs = a[4] if len(a) > 4 else 2
log = synth.power_law_log(users, tracks, plays, 20260108, test_ratio=0.0)
P, Q = synth.init_factors(log.m, log.n, d, 20261108)
m = log.m
top = {u: [(u * 7 + 3) % m, (u * 11 + 5) % m] for u in range(m) if u % 5}
top = {u: [f for f in fr if f != u] for u, fr in top.items()}
ip_indptr, ip_items = implicit_positive_lists(m, log.uq_indptr, log.uq_items, top)
eng = Engine(0)
eng.set_interactions(log.m, log.n, log.ev_indptr, log.ev_items, log.uq_indptr, log.uq_items)
eng.set_factors(P, Q)
eng.cune_set_implicit(ip_indptr, ip_items)
T = int(log.ev_indptr[-1])
print("users %d tracks %d events %d implicit positives %d d %d" % (m, log.n, T, len(ip_items), d), flush=True)
for ep in range(epochs):
    eng.sync()
    eng.timer_start()
    loss = eng.cune_epoch(0.02, 0.01, 0.01, 2.0, 7, ep, MODE_HOGWILD)
    ms = eng.timer_stop()
    # 3 repeats x 4 Q rows moved + Q[i] twice per event
    print("epoch %d  %9.3f ms  %.3e events/s  %.3e repeats/s  %.1f GB/s algorithmic  loss %.6g"
          % (ep, ms, T / ms * 1e3, 3 * T / ms * 1e3, T * 14 * d * 4 / ms / 1e6, loss), flush=True)
Pg, Qg = eng.get_factors()
assert np.isfinite(Pg).all() and np.isfinite(Qg).all()
