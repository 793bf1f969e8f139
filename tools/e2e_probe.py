#!/usr/bin/env python
"""Where does the end-to-end step of bench.py spend its time?  (C2 shape, pinned host buffers)"""
import os, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from yue_b200 import synth
from yue_b200.engine import MODE_HOGWILD, Engine, PinnedArray
log = synth.power_law_log_torch(1_000_000, 200_000, 50_000_000, 20260103, device="cuda")
for name in ("ev_indptr", "ev_items", "uq_indptr", "uq_items"):
    a = getattr(log, name); pa = PinnedArray(a.shape, a.dtype); pa.array[:] = a; setattr(log, name, pa.array); setattr(log, "_pin_" + name, pa)
m, n = log.m, log.n
pP, pQ = PinnedArray((m, 64), np.float32), PinnedArray((n, 64), np.float32)
P0, Q0 = synth.init_factors(m, n, 64, 1)
pP.array[:], pQ.array[:] = P0, Q0
eng = Engine(0)
for rep in range(3):
    t = [time.perf_counter()]
    eng.set_interactions(m, n, log.ev_indptr, log.ev_items, log.uq_indptr, log.uq_items); t.append(time.perf_counter())
    eng.set_factors(pP.array, pQ.array); t.append(time.perf_counter())
    loss = eng.bpr_epoch(0.02, 0.01, 0.01, 1, rep, MODE_HOGWILD); t.append(time.perf_counter())
    eng.frob2(); t.append(time.perf_counter())
    eng.get_factors(pP.array, pQ.array); t.append(time.perf_counter())
    print("set_interactions %.1f ms  set_factors %.1f  epoch %.1f  frob2 %.1f  get_factors %.1f  total %.1f" %
          tuple([1e3 * (t[i + 1] - t[i]) for i in range(5)] + [1e3 * (t[-1] - t[0])]), flush=True)
