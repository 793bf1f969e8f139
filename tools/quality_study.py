#!/usr/bin/env python
"""Recall@10 / NDCG@10 of the Hogwild throughput mode against the serial-order mode (which
reproduces the reference loop) on one synthetic log, sweeping learning rate, concurrency
(events per warp) and the shared-memory hot-row path.  Evidence for DESIGN.md's stability notes;
run on a B200: python tools/quality_study.py > gpurun_out/quality_study.txt"""
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from yue_b200 import synth  # noqa: E402
from yue_b200.engine import MODE_HOGWILD, MODE_SERIAL, RANK_EXACT, Engine  # noqa: E402


def run(log, P, Q, mode, epochs, lr, seed, env):
    for k, v in env.items():
        os.environ[k] = str(v)
    eng = Engine(0)
    for k in env:
        os.environ.pop(k, None)
    try:
        eng.set_interactions(log.m, log.n, log.ev_indptr, log.ev_items, log.uq_indptr, log.uq_items)
        eng.set_factors(P, Q)
        t0 = time.time()
        for ep in range(epochs):
            try:
                loss = eng.bpr_epoch(lr, 0.01, 0.01, seed, ep, mode)
            except Exception:                       # NaN/inf loss: report as diverged
                loss = float("inf")
                break
        dt = time.time() - t0
        users = log.test_users()
        eng.rank_topn(users, 10, RANK_EXACT)
        eng.set_test_set(log.test_indptr, log.test_items)
        sums, _ = eng.rank_metrics([10])               # K6 (checked against the oracle in tests/test_metrics_gpu.py)
        _, Qf = eng.get_factors()
    finally:
        eng.close()
    return sums[0, 1] / len(users), sums[0, 3] / len(users), loss, dt, float(np.linalg.norm(Qf[0]))


def main():
    users, tracks, plays, d = 12000, 3000, 900000, 32
    if len(sys.argv) > 1:
        users, tracks, plays = (int(x) for x in sys.argv[1:4])
    if len(sys.argv) > 4:
        d = int(sys.argv[4])
    c2_epochs = int(sys.argv[5]) if len(sys.argv) > 5 else 0
    log = (synth.power_law_log_torch(users, tracks, plays, seed=33, test_ratio=0.2) if plays > 3_000_000
           else synth.power_law_log(users, tracks, plays, seed=33))
    P, Q = synth.init_factors(log.m, log.n, d, seed=5)
    share = np.bincount(log.ev_items, minlength=log.n).max() / log.train_size
    print("log: %d users x %d tracks, %d train events, hottest track share %.3f, d=%d" % (log.m, log.n, log.train_size, share, d))
    print("%-44s %8s %8s %12s %7s %8s" % ("config", "recall", "ndcg", "last loss", "sec", "|Q[0]|"))
    big = log.train_size > 2_000_000
    sweeps = ((0.02, 10),) if big else ((0.02, 20), (0.05, 10))
    hot = dict(YUE_SGD_HOT_MAX=64, YUE_SGD_HOT_MIN_COUNT=16384, YUE_SGD_HOT_FLUSH=4)
    nohot = dict(YUE_SGD_HOT_MAX=0)
    configs = [("serial seed 99", MODE_SERIAL, {})]
    if not big:
        configs.append(("serial seed 100", MODE_SERIAL, {}))
    for epw in (65536, 16384, 4096, 2048) if big else (16384, 4096, 1024):
        configs.append(("hogwild %d ev/warp, no hot" % epw, MODE_HOGWILD, dict(YUE_SGD_MIN_EVENTS_PER_WARP=epw, **nohot)))
        configs.append(("hogwild %d ev/warp, hot flush 4" % epw, MODE_HOGWILD, dict(YUE_SGD_MIN_EVENTS_PER_WARP=epw, **hot)))
    if c2_epochs:       # full-GPU concurrency at the true scale
        # STUDY_CONFIGS="name:K=V,K=V;name2:K=V" (env of each configuration), STUDY_REPS runs of each;
        # STUDY_SERIAL=1 trains the serial-order reference first (minutes at C2)
        sweeps = ((0.02, c2_epochs),)
        configs = []
        if os.environ.get("STUDY_SERIAL", "1") == "1":
            configs.append(("serial seed 99", MODE_SERIAL, {}))
        reps = int(os.environ.get("STUDY_REPS", "3"))
        for spec in os.environ.get("STUDY_CONFIGS", "default:").split(";"):
            name, _, kvs = spec.partition(":")
            env = dict(kv.split("=") for kv in kvs.split(",") if kv)
            configs += [("%s, run %d" % (name, k), MODE_HOGWILD, env) for k in range(reps)]
    for lr, epochs in sweeps:
        base = tuple(float(x) for x in os.environ["STUDY_BASE"].split(",")) if os.environ.get("STUDY_BASE") else None
        for name, mode, env in configs:
            seed = 100 if name.endswith("100") else 99
            r, n, loss, dt, qn = run(log, P, Q, mode, epochs, lr, seed, env)
            if base is None:
                base = (r, n)
            print("lr %.2f x%-2d %-36s %8.4f %8.4f %14.1f %7.2f %8.3f   d_recall %+.4f d_ndcg %+.4f"
                  % (lr, epochs, name, r, n, loss, dt, qn, r - base[0], n - base[1]), flush=True)


if __name__ == "__main__":
    main()
