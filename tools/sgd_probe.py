#!/usr/bin/env python
"""Time the SGD epoch kernel at config C2 (or a given size) under several env settings, one log.
usage: python tools/sgd_probe.py "K=V,K=V" "K=V" ...   (each argument = one configuration)
env PROBE_SIZE="users,tracks,plays,d" overrides the C2 shape; PROBE_EPOCHS the timed epochs."""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from yue_b200 import synth  # noqa: E402
from yue_b200.engine import MODE_HOGWILD, Engine  # noqa: E402

users, tracks, plays, d = (int(x) for x in os.environ.get("PROBE_SIZE", "1000000,200000,50000000,64").split(","))
epochs = int(os.environ.get("PROBE_EPOCHS", "5"))
beta = float(os.environ.get("PROBE_BETA", "1.0"))
log = synth.power_law_log_torch(users, tracks, plays, 20260103, device="cuda", beta=beta)
print("beta", beta, "train events", log.train_size, flush=True)
P, Q = synth.init_factors(log.m, log.n, d, int(os.environ.get('PROBE_PQ_SEED', '20261103')))
SS = int(os.environ.get('PROBE_SEED', '1'))
if os.environ.get("PROBE_EMPTY_CACHE"):
    import torch
    torch.cuda.empty_cache()
if os.environ.get("PROBE_PIN"):
    from yue_b200.engine import PinnedArray
    keep = []
    for name in ("ev_indptr", "ev_items", "uq_indptr", "uq_items"):
        a = getattr(log, name); pa = PinnedArray(a.shape, a.dtype); pa.array[:] = a; setattr(log, name, pa.array); keep.append(pa)
    if os.environ.get("PROBE_PIN") == "2":
        pP, pQ = PinnedArray(P.shape, np.float32), PinnedArray(Q.shape, np.float32)
        pP.array[:], pQ.array[:] = P, Q
        P, Q = pP.array, pQ.array
for cfg in sys.argv[1:] or [""]:
    keys = []
    for kv in filter(None, cfg.split(",")):
        k, v = kv.split("=")
        os.environ[k] = v
        keys.append(k)
    eng = Engine(0)
    eng.set_interactions(log.m, log.n, log.ev_indptr, log.ev_items, log.uq_indptr, log.uq_items)
    eng.set_factors(P, Q)
    eng.bpr_epoch(0.02, 0.01, 0.01, SS, 0, MODE_HOGWILD, want_loss=False)
    ms = []
    for ep in range(1, 1 + epochs):
        eng.sync()
        eng.timer_start()
        eng.bpr_epoch(0.02, 0.01, 0.01, SS, ep, MODE_HOGWILD, want_loss=False)
        ms.append(eng.timer_stop())
    if os.environ.get("PROBE_TRACE"):
        print("   per epoch ms:", " ".join("%.2f" % x for x in ms), flush=True)
    loss = eng.bpr_epoch(0.02, 0.01, 0.01, 1, epochs + 1, MODE_HOGWILD)
    print("%-70s ms/epoch min %.2f mean %.2f  -> %.3e triplets/s   loss %.1f" %
          (cfg or "(default)", min(ms), float(np.mean(ms)), log.train_size / (min(ms) * 1e-3), loss), flush=True)
    eng.close()
    for k in keys:
        os.environ.pop(k, None)
