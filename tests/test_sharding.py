"""Host-side logic of the user-sharded multi-GPU path on CPU: the event-balanced user split, the
shard-independent sampler, and the Q reconciliation over a real 2-process gloo group."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from oracle import bpr_ref, philox, record_ref
from yue_b200 import sharding, synth


def test_shard_users_by_events_balances_power_law():
    log = synth.power_law_log(5000, 800, 200000, seed=2)
    for world in (2, 4, 8):
        b = sharding.shard_users_by_events(log.ev_indptr, world)
        assert b[0] == 0 and b[-1] == log.m and (np.diff(b) >= 0).all()
        ev = np.diff(log.ev_indptr[b])
        assert ev.sum() == log.train_size
        assert ev.max() <= log.train_size / world + np.diff(log.ev_indptr).max()
    # degenerate: more ranks than users with events
    b = sharding.shard_users_by_events(np.array([0, 5, 5, 5]), 4)
    assert b[0] == 0 and b[-1] == 3 and (np.diff(b) >= 0).all()


def test_local_shard_is_consistent_and_sampler_is_shard_independent():
    log = synth.power_law_log(600, 300, 30000, seed=3)
    b = sharding.shard_users_by_events(log.ev_indptr, 3)
    ev_user = record_ref.ev_users(log.ev_indptr)
    full = philox.sample_negatives(9, 1, ev_user, log.n, log.uq_indptr, log.uq_items)
    got = []
    for r in range(3):
        s = sharding.local_shard(log.ev_indptr, log.ev_items, log.uq_indptr, log.uq_items, b, r)
        assert s["ev_indptr"][0] == 0 and s["uq_indptr"][0] == 0 and len(s["ev_items"]) == s["ev_indptr"][-1]
        got.append(philox.sample_negatives(9, 1, record_ref.ev_users(s["ev_indptr"]), log.n, s["uq_indptr"],
                                           s["uq_items"], event_base=s["event_base"]))
    assert np.array_equal(np.concatenate(got), full)


def _worker(rank, world, port, ret):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    log = synth.power_law_log(400, 200, 20000, seed=5)
    P, Q = synth.init_factors(log.m, log.n, 16, seed=6)
    b = sharding.shard_users_by_events(log.ev_indptr, world)
    s = sharding.local_shard(log.ev_indptr, log.ev_items, log.uq_indptr, log.uq_items, b, rank)
    u0 = s["user_begin"]
    Pl, Ql = P[u0:u0 + s["m_local"]].copy(), Q.copy()
    ev_user = record_ref.ev_users(s["ev_indptr"])
    for ep in range(2):                                   # local serial epoch (the oracle plays the kernel), then exchange
        snap = torch.from_numpy(Ql.copy())
        neg = philox.sample_negatives(7, ep, ev_user, log.n, s["uq_indptr"], s["uq_items"], event_base=s["event_base"])
        bpr_ref.sgd_epoch(Pl, Ql, ev_user, s["ev_items"], neg, 0.02, 0.01, 0.01)
        Ql = sharding.reconcile_q(torch.from_numpy(Ql), snap, lambda t: dist.all_reduce(t)).numpy().copy()
    gathered = [None] * world
    dist.all_gather_object(gathered, (u0, Pl, Ql))
    if rank == 0:
        ret["out"] = gathered
    dist.destroy_process_group()


def test_two_rank_gloo_reconciliation_equals_sum_of_deltas():
    with socket.socket() as sk:
        sk.bind(("127.0.0.1", 0))
        port = sk.getsockname()[1]
    mgr = mp.Manager()
    ret = mgr.dict()
    mp.spawn(_worker, args=(2, port, ret), nprocs=2, join=True)
    (u0a, Pa, Qa), (u0b, Pb, Qb) = ret["out"]
    assert np.array_equal(Qa, Qb)                          # replicas agree bit for bit after the exchange
    # single-process emulation of the same schedule: both shards from the same snapshot, deltas summed
    log = synth.power_law_log(400, 200, 20000, seed=5)
    P, Q = synth.init_factors(log.m, log.n, 16, seed=6)
    b = sharding.shard_users_by_events(log.ev_indptr, 2)
    shards = [sharding.local_shard(log.ev_indptr, log.ev_items, log.uq_indptr, log.uq_items, b, r) for r in range(2)]
    Pl = [P[s["user_begin"]:s["user_begin"] + s["m_local"]].copy() for s in shards]
    for ep in range(2):
        qs = []
        for r, s in enumerate(shards):
            ql = Q.copy()
            eu = record_ref.ev_users(s["ev_indptr"])
            neg = philox.sample_negatives(7, ep, eu, log.n, s["uq_indptr"], s["uq_items"], event_base=s["event_base"])
            bpr_ref.sgd_epoch(Pl[r], ql, eu, s["ev_items"], neg, 0.02, 0.01, 0.01)
            qs.append(ql)
        Q = Q + ((qs[0] - Q) + (qs[1] - Q))
    assert np.array_equal(Qa, Q)
    assert np.array_equal(Pa, Pl[0]) and np.array_equal(Pb, Pl[1])


def test_interleaved_shards_partition_the_log():
    """interleaved_users + local_shard_of_users: every user on exactly one rank, events and play rows intact and in
    order, event offsets map local event indices back to the unsharded log's."""
    from yue_b200 import synth
    log = synth.power_law_log(101, 60, 5000, seed=2)
    seen = []
    for world in (1, 2, 4):
        total_events = 0
        for rank in range(world):
            users = sharding.interleaved_users(log.m, world, rank)
            sh = sharding.local_shard_of_users(log.ev_indptr, log.ev_items, log.uq_indptr, log.uq_items, users)
            assert sh["m_local"] == len(users) and sh["ev_indptr"][0] == 0 and sh["uq_indptr"][0] == 0
            for k, u in enumerate(users):
                a, b = sh["ev_indptr"][k], sh["ev_indptr"][k + 1]
                assert np.array_equal(sh["ev_items"][a:b], log.ev_items[log.ev_indptr[u]:log.ev_indptr[u + 1]])
                assert a + sh["event_offsets"][k] == log.ev_indptr[u]
                a, b = sh["uq_indptr"][k], sh["uq_indptr"][k + 1]
                assert np.array_equal(sh["uq_items"][a:b], log.uq_items[log.uq_indptr[u]:log.uq_indptr[u + 1]])
            total_events += int(sh["ev_indptr"][-1])
            seen.append(users)
        assert total_events == log.train_size
    assert np.array_equal(np.sort(np.concatenate(seen[-4:])), np.arange(log.m))


def test_saturation_weights_limits():
    """1 for tracks that are (almost) never played between two exchanges, 1/G for the most played ones, monotone."""
    counts = np.array([0, 1, 100, 10_000, 1_000_000, 100_000_000])
    for world in (2, 4, 8):
        w = sharding.saturation_weights(counts, world, sub_epochs=8, kappa=1e-3)
        assert w[0] == 1.0 and abs(w[1] - 1.0) < 1e-3
        assert abs(w[-1] - 1.0 / world) < 1e-6
        assert np.all(np.diff(w) <= 1e-7)
    # the closed form: G ranks each contracting by a reproduce one stream contracting by a^G
    a, G = 0.7, 4
    w = sharding.saturation_weights(np.array([-np.log(a) / 1e-3 * G * 8]), G, 8, 1e-3)[0]
    assert abs(w * G * (1 - a) - (1 - a ** G)) < 1e-6


# ---- WRMF: row-sharded half-sweeps (the oracle plays the kernel), rows exchanged over a 2-process gloo group ----
def _wrmf_half_rows(out, other, indptr, idx, cnt, lo, hi, reg, want_loss=False):
    """oracle half-sweep restricted to rows [lo, hi): the Gram matrix is of the WHOLE other table, like the kernel's"""
    from oracle import wrmf_ref
    sub = out[lo:hi]                                       # view: solved in place
    return wrmf_ref.half_sweep(sub, other, indptr[lo:hi + 1], idx, cnt, reg, gram="f64", want_loss=want_loss)


def _wrmf_worker(rank, world, port, ret):
    from oracle import wrmf_ref
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    log = synth.power_law_log(300, 150, 9000, seed=21)
    X, Y = synth.init_factors(log.m, log.n, 12, seed=22)
    X, Y = X * 10, Y * 10
    cnt = wrmf_ref.pair_counts(log.ev_indptr, log.ev_items, log.uq_indptr, log.uq_items)
    itp, itu, itc = wrmf_ref.transpose(log.m, log.n, log.uq_indptr, log.uq_items, cnt)
    ub, tb = sharding.shard_users_by_events(log.uq_indptr, world), sharding.shard_users_by_events(itp, world)
    Xt, Yt = torch.from_numpy(X), torch.from_numpy(Y)     # share memory with X, Y
    losses = []
    for _ in range(2):
        loss = _wrmf_half_rows(X, Y, log.uq_indptr, log.uq_items, cnt, ub[rank], ub[rank + 1], 0.5, want_loss=True)
        sharding.exchange_rows(Xt, ub, dist)
        _wrmf_half_rows(Y, X, itp, itu, itc, tb[rank], tb[rank + 1], 0.5)
        sharding.exchange_rows(Yt, tb, dist)
        lt = torch.tensor([loss], dtype=torch.float64)
        dist.all_reduce(lt)
        losses.append(float(lt.item()))
    gathered = [None] * world
    dist.all_gather_object(gathered, (X, Y, losses))
    if rank == 0:
        ret["out"] = gathered
    dist.destroy_process_group()


def test_two_rank_gloo_wrmf_row_exchange_is_exact():
    from oracle import wrmf_ref
    with socket.socket() as sk:
        sk.bind(("127.0.0.1", 0))
        port = sk.getsockname()[1]
    mgr = mp.Manager()
    ret = mgr.dict()
    mp.spawn(_wrmf_worker, args=(2, port, ret), nprocs=2, join=True)
    (Xa, Ya, la), (Xb, Yb, lb) = ret["out"]
    assert np.array_equal(Xa, Xb) and np.array_equal(Ya, Yb) and la == lb
    log = synth.power_law_log(300, 150, 9000, seed=21)
    X, Y = synth.init_factors(log.m, log.n, 12, seed=22)
    X, Y = X * 10, Y * 10
    cnt = wrmf_ref.pair_counts(log.ev_indptr, log.ev_items, log.uq_indptr, log.uq_items)
    itp, itu, itc = wrmf_ref.transpose(log.m, log.n, log.uq_indptr, log.uq_items, cnt)
    losses = [wrmf_ref.iteration(X, Y, log.uq_indptr, log.uq_items, cnt, itp, itu, itc, 0.5, gram="f64") for _ in range(2)]
    assert np.array_equal(Xa, X) and np.array_equal(Ya, Y)           # no reduction anywhere: bit-identical to one process
    assert la == pytest.approx(losses, rel=1e-12)
