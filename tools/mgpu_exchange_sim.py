#!/usr/bin/env python
"""CPU simulation of the multi-GPU SGD exchange schemes (DESIGN.md sections 6 and 9, item 1) -- an experiment, not product
code and not a parity claim: its own 15-line numpy BPR loop, no CUDA, no oracle import.

Question it answers before GPU time is spent on the round-2 plan: if the ranks share ONE copy of the most played tracks'
rows (peer memory) and reconcile only the long tail by summing deltas a few times per epoch, does Recall@10 / NDCG@10 stay
within 0.5 points of the serial order -- where summing the deltas of ALL rows does not?

G ranks are emulated in one process: users interleaved over the ranks (sharding.interleaved_users), every rank walks its
users in stream order, the ranks advance in turns of `--quantum` events (stand-in for running at the same time), P rows
are private by construction, Q rows of the H hottest tracks are one shared array, the others one copy per rank that is
reconciled `--exchanges` times per epoch:  Q <- snapshot + sum_r (Q_r - snapshot).

usage: python tools/mgpu_exchange_sim.py [--users 20000 --tracks 5000 --plays 400000 --d 16 --epochs 6 --ranks 2]
"""
import argparse
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from yue_b200 import synth  # noqa: E402


def negatives(rng, ev_user, played, n):
    j = rng.integers(0, n, len(ev_user))
    bad = played[ev_user, j]
    while bad.any():
        j[bad] = rng.integers(0, n, int(bad.sum()))
        bad = played[ev_user, j]
    return j


def step(Pu, Qi, Qj, lr, reg):
    """BPR.py:50-57 on three row views (in place)."""
    x = float(Pu.dot(Qi) - Pu.dot(Qj))
    g = lr / (1.0 + np.exp(x))                 # lr (1 - sigmoid(x))
    d = Qi - Qj
    Pu += g * d
    Qi += g * Pu
    Qj -= g * Pu
    Pu *= 1 - lr * reg
    Qi *= 1 - lr * reg
    Qj *= 1 - lr * reg


def metrics(P, Q, log, N=10):
    users = log.test_users()
    rec = ndcg = 0.0
    disc = 1.0 / np.log2(np.arange(2, N + 2))
    for b0 in range(0, len(users), 2048):
        ub = users[b0:b0 + 2048]
        S = P[ub] @ Q.T
        for r, u in enumerate(ub):
            S[r, log.uq_items[log.uq_indptr[u]:log.uq_indptr[u + 1]]] = -np.inf
        top = np.argpartition(-S, N, axis=1)[:, :N]
        order = np.argsort(-np.take_along_axis(S, top, 1), axis=1, kind="stable")
        top = np.take_along_axis(top, order, 1)
        for r, u in enumerate(ub):
            held = log.test_items[log.test_indptr[u]:log.test_indptr[u + 1]]
            hit = np.isin(top[r], held)
            rec += hit.sum() / len(held)
            ndcg += (hit * disc).sum() / disc[:min(len(held), N)].sum()
    return rec / len(users), ndcg / len(users)


def train(log, P, Q, negs, lr, reg, ranks, hot, exchanges, quantum, merge=0):
    """ranks == 1: the serial order.  hot: boolean mask of the tracks whose rows are shared -- as ONE copy every rank reads
    and writes (merge == 0), or as a local copy per rank that is merged with a master copy every `merge` events of the
    rank: delta = local - base is added to the master, the master becomes the new local copy and base (what a rank's
    merger warp would do with remote REDs and one remote load of the hot table, without any barrier)."""
    ev_user = np.repeat(np.arange(log.m), np.diff(log.ev_indptr))
    if ranks == 1:
        for neg in negs:
            for e in range(len(ev_user)):
                step(P[ev_user[e]], Q[log.ev_items[e]], Q[neg[e]], lr, reg)
        return P, Q
    order = [np.concatenate([np.arange(log.ev_indptr[u], log.ev_indptr[u + 1]) for u in range(r, log.m, ranks)]) for r in range(ranks)]
    Qr = [Q.copy() for _ in range(ranks)]          # per-rank copies (tail rows); hot rows live in Q itself
    tail = ~hot
    base = [Q[hot].copy() for _ in range(ranks)]
    since = [0] * ranks
    for neg in negs:
        snap = Q.copy()
        pos = [0] * ranks
        bounds = [[len(o) * (x + 1) // exchanges for x in range(exchanges)] for o in order]
        for x in range(exchanges):
            live = True
            while live:
                live = False
                for r in range(ranks):
                    end = min(pos[r] + quantum, bounds[r][x])
                    for e in order[r][pos[r]:end]:
                        i, j = log.ev_items[e], neg[e]
                        if merge:
                            step(P[ev_user[e]], Qr[r][i], Qr[r][j], lr, reg)
                        else:
                            step(P[ev_user[e]], (Q if hot[i] else Qr[r])[i], (Q if hot[j] else Qr[r])[j], lr, reg)
                    if merge:
                        since[r] += end - pos[r]
                        if since[r] >= merge:
                            since[r] = 0
                            Q[hot] += Qr[r][hot] - base[r]
                            Qr[r][hot] = Q[hot]
                            base[r] = Q[hot].copy()
                    live = live or end > pos[r]
                    pos[r] = end
            if merge:                               # bring every rank's hot rows up to date before the tail exchange
                for r in range(ranks):
                    Q[hot] += Qr[r][hot] - base[r]
                    Qr[r][hot] = Q[hot]
                    base[r] = Q[hot].copy()
                    since[r] = 0
                for r in range(ranks):
                    Qr[r][hot] = Q[hot]
                    base[r] = Q[hot].copy()
            new = snap + sum(q - snap for q in Qr)  # the all-reduce of the deltas (tail rows; hot rows are not in Qr)
            Q[tail] = new[tail]
            for q in Qr:
                q[tail] = Q[tail]
            snap = Q.copy()
    return P, Q


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--users", type=int, default=20000)
    ap.add_argument("--tracks", type=int, default=5000)
    ap.add_argument("--plays", type=int, default=400000)
    ap.add_argument("--d", type=int, default=16)
    ap.add_argument("--epochs", type=int, default=6)
    ap.add_argument("--ranks", type=int, default=2)
    ap.add_argument("--quantum", type=int, default=64)
    ap.add_argument("--lr", type=float, default=0.05)
    ap.add_argument("--reg", type=float, default=0.01)
    ap.add_argument("--seeds", type=int, default=2)
    ap.add_argument("--hot", type=str, default="0,64,512", help="comma list: how many of the hottest rows are one shared copy")
    ap.add_argument("--exchanges", type=str, default="4,32", help="comma list: tail exchanges per epoch")
    ap.add_argument("--merge", type=str, default="", help="comma list of merge periods (events per rank) for 64 local hot rows")
    args = ap.parse_args()
    log = synth.power_law_log(args.users, args.tracks, args.plays, seed=77)
    ev_user = np.repeat(np.arange(log.m), np.diff(log.ev_indptr))
    played = np.zeros((log.m, log.n), dtype=bool)
    played[np.repeat(np.arange(log.m), np.diff(log.uq_indptr)), log.uq_items] = True
    counts = np.bincount(log.ev_items, minlength=log.n)
    by_count = np.argsort(-counts)
    print("log: %d users x %d tracks, %d train events; hottest track %.1f %% of the positives; %d rank(s), %d epochs, d=%d"
          % (log.m, log.n, log.train_size, 100.0 * counts.max() / log.train_size, args.ranks, args.epochs, args.d), flush=True)
    configs = [("serial", 1, 0, 1, 0)]
    if args.merge:
        for K in (int(x) for x in args.merge.split(",")):
            configs.append(("%d ranks, 64 hot rows local, merged with the master copy every %6d events, 8 exchanges/epoch" %
                            (args.ranks, K), args.ranks, 64, 8, K))
    else:
        for H in (int(x) for x in args.hot.split(",")):
            for X in (int(x) for x in args.exchanges.split(",")):
                configs.append(("%d ranks, %4d shared hot rows (%4.1f %% of the positives), %2d exchanges/epoch" %
                                (args.ranks, H, 100.0 * counts[by_count[:H]].sum() / log.train_size, X), args.ranks, H, X, 0))
    for seed in range(args.seeds):
        rng = np.random.default_rng(1000 + seed)
        negs = [negatives(rng, ev_user, played, log.n) for _ in range(args.epochs)]
        P0, Q0 = synth.init_factors(log.m, log.n, args.d, seed=5)
        base = None
        for name, ranks, H, X, K in configs:
            hot = np.zeros(log.n, dtype=bool)
            hot[by_count[:H]] = True
            t0 = time.time()
            P, Q = train(log, P0.copy(), Q0.copy(), negs, args.lr, args.reg, ranks, hot, X, args.quantum, K)
            rec, nd = metrics(P, Q, log)
            if base is None:
                base = (rec, nd)
            print("seed %d  %-78s recall@10 %.4f (%+.4f)  ndcg@10 %.4f (%+.4f)  |Q[hottest]| %.3f  [%.0f s]"
                  % (seed, name, rec, rec - base[0], nd, nd - base[1], float(np.linalg.norm(Q[by_count[0]])), time.time() - t0), flush=True)


if __name__ == "__main__":
    main()
