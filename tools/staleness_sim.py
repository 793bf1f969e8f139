#!/usr/bin/env python
"""CPU experiment (not product code, no oracle import): how much read-to-write delay on the SHARED rows of Q does the BPR
trajectory tolerate before Recall@10 / NDCG@10 leave the 0.5-point gate?

Every asynchronous schedule -- one GPU's 1 776 warps, or N GPUs sharing the hot rows over NVLink -- computes an update from
rows read some time before its own change becomes visible.  Model: events in the reference's order; event e reads Q as it
was `delay` events ago (P[u] is private to its warp and always fresh), its changes of Q[i], Q[j] become visible `delay`
events later.  The number that matters for a hot row is D = delay x (its share of the positives) = touches in flight.
On one B200 the blocked kernel reads a block ahead: ~1.85 us x 3.85e9 triplets/s ~ 7 K events in flight, D ~ 550 for config
C2's most played track (7.8 %).  A row shared by N GPUs over NVLink adds >= 2 us of round trip at N times the touch rate.

usage: python tools/staleness_sim.py [--delays 0,2000,8000,32000] [--users 20000 --tracks 5000 --plays 400000 --epochs 4]
"""
import argparse
import os
import sys
import time
from collections import deque

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from tools.mgpu_exchange_sim import metrics, negatives  # noqa: E402
from yue_b200 import synth  # noqa: E402


def train(log, P, Q, negs, lr, reg, delay, only_hot=None):
    ev_user = np.repeat(np.arange(log.m), np.diff(log.ev_indptr))
    c = 1.0 - lr * reg
    pend = deque()
    for neg in negs:
        for e in range(len(ev_user)):
            u, i, j = ev_user[e], log.ev_items[e], neg[e]
            Pu, Qi, Qj = P[u], Q[i], Q[j]
            x = float(Pu.dot(Qi) - Pu.dot(Qj))
            g = lr / (1.0 + np.exp(x))
            Pn = Pu + g * (Qi - Qj)
            di = (Qi + g * Pn) * c - Qi
            dj = (Qj - g * Pn) * c - Qj
            P[u] = Pn * c
            if delay == 0:
                Q[i] += di; Q[j] += dj
            else:
                # rows outside `only_hot` (when given) are updated at once: isolates the effect of the hot rows
                if only_hot is not None and not only_hot[i]:
                    Q[i] += di
                else:
                    pend.append((i, di))
                if only_hot is not None and not only_hot[j]:
                    Q[j] += dj
                else:
                    pend.append((j, dj))
                pend.append(None)                      # event boundary
                if e >= delay:
                    while True:
                        it = pend.popleft()
                        if it is None:
                            break
                        Q[it[0]] += it[1]
        while pend:                                    # epoch end: everything becomes visible (kernel boundary)
            it = pend.popleft()
            if it is not None:
                Q[it[0]] += it[1]
    return P, Q


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--users", type=int, default=20000)
    ap.add_argument("--tracks", type=int, default=5000)
    ap.add_argument("--plays", type=int, default=400000)
    ap.add_argument("--d", type=int, default=16)
    ap.add_argument("--epochs", type=int, default=4)
    ap.add_argument("--lr", type=float, default=0.05)
    ap.add_argument("--reg", type=float, default=0.01)
    ap.add_argument("--delays", type=str, default="0,2000,8000,32000")
    ap.add_argument("--hot", type=int, default=0, help="> 0: only the H hottest rows are delayed, the others are always fresh")
    ap.add_argument("--seed", type=int, default=0)
    args = ap.parse_args()
    log = synth.power_law_log(args.users, args.tracks, args.plays, seed=77)
    ev_user = np.repeat(np.arange(log.m), np.diff(log.ev_indptr))
    played = np.zeros((log.m, log.n), dtype=bool)
    played[np.repeat(np.arange(log.m), np.diff(log.uq_indptr)), log.uq_items] = True
    counts = np.bincount(log.ev_items, minlength=log.n)
    share = counts.max() / log.train_size
    print("log: %d users x %d tracks, %d train events; hottest track %.1f %% of the positives; %d epochs, d=%d, lr %.3f"
          % (log.m, log.n, log.train_size, 100.0 * share, args.epochs, args.d, args.lr), flush=True)
    rng = np.random.default_rng(1000 + args.seed)
    negs = [negatives(rng, ev_user, played, log.n) for _ in range(args.epochs)]
    P0, Q0 = synth.init_factors(log.m, log.n, args.d, seed=5)
    only_hot = None
    if args.hot > 0:
        only_hot = np.zeros(log.n, dtype=bool)
        only_hot[np.argsort(-counts)[:args.hot]] = True
    base = None
    for delay in (int(x) for x in args.delays.split(",")):
        t0 = time.time()
        P, Q = train(log, P0.copy(), Q0.copy(), negs, args.lr, args.reg, delay, only_hot)
        rec, nd = metrics(P, Q, log)
        if base is None:
            base = (rec, nd)
        print("delay %6d events (D = %5.0f touches of the hottest row in flight%s)  recall@10 %.4f (%+.4f)  ndcg@10 %.4f (%+.4f)  |Q[hottest]| %.3f  [%.0f s]"
              % (delay, delay * share, ", only the %d hottest rows delayed" % args.hot if args.hot else "", rec, rec - base[0], nd, nd - base[1],
                 float(np.linalg.norm(Q[np.argmax(counts)])), time.time() - t0), flush=True)


if __name__ == "__main__":
    main()
