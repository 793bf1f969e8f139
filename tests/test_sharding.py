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


# ---- round 2: shared hot rows + overlapped exchange (yue_b200/sharding.py: SharedHotTrainer), host logic ------------------
def test_select_hot_tracks_rule_and_order():
    c = np.zeros(1000, np.int64)
    c[[5, 7, 9, 11]] = [50000, 90000, 50000, 16000]          # 11 is below min_count
    c[100:600] = 3000                                          # the tail: 1.5 M events
    tracks, counts, total = sharding.select_hot_tracks(c)
    assert total == int(c.sum())
    assert tracks.tolist() == [7, 5, 9] and counts.tolist() == [90000, 50000, 50000]      # most played first, ties by id
    assert sharding.select_hot_tracks(c, hot_max=2)[0].tolist() == [7, 5]
    # a track must also carry more than 1/hot_div of ALL events: 1/128 is the one-GPU rule, 1/4096 the sharded trainer's default
    c2 = c.copy(); c2[100:600] = 30000
    assert sharding.select_hot_tracks(c2, hot_div=128)[0].tolist() == []
    wide = sharding.select_hot_tracks(c2)[0]
    assert len(wide) == 248 and wide[:3].tolist() == [7, 5, 9] and set(wide[3:].tolist()) <= set(range(100, 600))


def test_wrmf_row_ranges_are_balanced_by_cost_and_default_rules():
    """shard_rows_by_cost: every row pays a factorisation besides its entries, so the ranges hold about the same COST; by
    entries alone the last rank of a power-law log gets several times the rows.  And the sharded SGD trainer's defaults."""
    log = synth.power_law_log(20000, 3000, 600000, seed=8)
    e = np.diff(log.uq_indptr).astype(np.float64)
    cost = np.where(e == 0, 0.0, np.where(e <= 16, 39.0 + 0.5 * e, 59.0 + e))
    for world in (2, 4, 8):
        b = sharding.shard_rows_by_cost(log.uq_indptr, world)
        assert b[0] == 0 and b[-1] == log.m and (np.diff(b) > 0).all()
        per = np.array([cost[b[r]:b[r + 1]].sum() for r in range(world)])
        assert per.max() <= 1.02 * per.mean() + cost.max()
        by_entries = np.diff(sharding.shard_users_by_events(log.uq_indptr, world))
        assert np.diff(b).max() < by_entries.max()
    assert [sharding.default_sub_epochs(w) for w in (1, 2, 4, 8)] == [32, 32, 64, 256]
    assert [sharding.default_asynchrony(w) for w in (1, 2, 8)] == [1.0, 0.25, 0.25]


class _FakeEngine:
    """Records the order of the C-ABI calls SharedHotTrainer makes."""
    device = 0

    def __init__(self):
        self.calls = []

    def __getattr__(self, name):
        def f(*a, **k):
            self.calls.append(name)
            if name == "hot_table_export":
                return b"\0" * 64, 0x1000
            if name in ("bpr_epoch_part", "apr_epoch_part"):
                return 1.5
            return None
        return f


def test_shared_hot_trainer_pipeline_order_two_thread_ranks():
    """Two ranks as threads (ThreadCtl): hot set agreed from the summed counts, tables exchanged, and per part
    epoch -> finish(previous) -> begin -> reduce, i.e. the sum of part k is applied after part k+1."""
    import threading
    shared = sharding.ThreadCtl.Shared(2)
    engs = [_FakeEngine(), _FakeEngine()]
    counts = np.zeros(500, np.int64); counts[3] = 40000; counts[10:400] = 2000
    out = [None, None]
    reduced = [0, 0]

    def run(r):
        ctl = sharding.ThreadCtl(shared, r)

        def reduce(e):
            reduced[r] += 1
            e.calls.append("reduce")
        tr = sharding.SharedHotTrainer(engs[r], ctl, counts // 2, sub_epochs=3, asynchrony=1.0, reduce=reduce)
        l = tr.epoch(0.02, 0.01, 0.01, 7, 0, want_loss=True)
        l2 = tr.epoch(0.003, 0.002, 0.01, 7, 1, want_loss=True, apr=(0.5, 2.0), finalize=True)
        out[r] = (tr, l, l2)
    th = [threading.Thread(target=run, args=(r,)) for r in range(2)]
    [t.start() for t in th]
    [t.join() for t in th]
    for r in range(2):
        tr, l, l2 = out[r]
        assert tr.hot_tracks.tolist() == [3] and l == 4.5 and l2 == 4.5 and reduced[r] == 6
        c = engs[r].calls
        setup = c[:c.index("bpr_epoch_part")]
        assert setup == ["set_hot_tracks", "hot_table_export", "enable_peer", "hot_share", "sync", "set_delta_weights",
                         "q_snapshot", "set_sgd_concurrency"]
        body = c[len(setup):]
        part = ["q_exchange_finish", "q_exchange_begin", "reduce"]
        assert body == (["bpr_epoch_part", "q_exchange_begin", "reduce"] + ["bpr_epoch_part"] + part + ["bpr_epoch_part"] + part
                        + (["apr_epoch_part"] + part) * 3 + ["q_exchange_finish", "sync", "hot_pull"])
        # concurrency: both ranks together = what one GPU would give the whole log (here 1.2 M events -> 50 warps)
        assert tr.n_warps == max(1, round(min(148 * 12, int(counts.sum()) // 16384) / 2)) and tr.n_ctas == 140


def _ctl_worker(rank, world, port, ret):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    ctl = sharding.TorchCtl(dist)
    got = ctl.allgather(("rank", rank, b"\1" * 4))
    s = ctl.allreduce_sum(np.arange(5, dtype=np.int64) * (rank + 1))
    ctl.barrier()
    if rank == 0:
        ret["got"], ret["sum"] = got, s
    dist.destroy_process_group()


def test_torch_ctl_over_gloo_world_size_2():
    with socket.socket() as sk:
        sk.bind(("127.0.0.1", 0))
        port = sk.getsockname()[1]
    ret = mp.Manager().dict()
    mp.spawn(_ctl_worker, args=(2, port, ret), nprocs=2, join=True)
    assert ret["got"] == [("rank", 0, b"\1" * 4), ("rank", 1, b"\1" * 4)]
    assert ret["sum"].tolist() == [0, 3, 6, 9, 12]
