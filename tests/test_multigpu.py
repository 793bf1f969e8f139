"""The Q reconciliation of the user-sharded path on real hardware: two handles on one GPU emulate
two ranks (always runs), and a 2-process NCCL run on 2 GPUs (skipped on a single-GPU box)."""
import os
import socket
import sys

import numpy as np
import pytest

from yue_b200 import sharding, synth
from yue_b200.engine import MODE_SERIAL, Engine

pytestmark = pytest.mark.gpu


def _shards(world):
    log = synth.power_law_log(3000, 900, 150000, seed=12)
    P, Q = synth.init_factors(log.m, log.n, 64, seed=13)
    b = sharding.shard_users_by_events(log.ev_indptr, world)
    return log, P, Q, [sharding.local_shard(log.ev_indptr, log.ev_items, log.uq_indptr, log.uq_items, b, r) for r in range(world)]


def test_two_handles_emulate_two_ranks():
    """delta = Q - snapshot, Q <- snapshot + sum of deltas, with the pack/apply kernels and the
    reduction done on the host in place of NCCL.  Serial-order epochs make both sides deterministic."""
    import ctypes as C
    from yue_b200._lib import BUF_Q_DELTA
    log, P, Q, shards = _shards(2)
    engs = [Engine(0), Engine(0)]
    try:
        Ploc = []
        for e, s in zip(engs, shards):
            e.set_interactions(s["m_local"], log.n, s["ev_indptr"], s["ev_items"], s["uq_indptr"], s["uq_items"],
                               user_begin=s["user_begin"], event_base=s["event_base"])
            e.set_factors(P[s["user_begin"]:s["user_begin"] + s["m_local"]], Q)
            e.q_snapshot()
        Qref = Q.copy()
        for ep in range(2):
            locals_ = []
            for e in engs:
                e.bpr_epoch(0.02, 0.01, 0.01, 5, ep, MODE_SERIAL)
                locals_.append(e.get_factors()[1])
            Qref = Qref + ((locals_[0] - Qref) + (locals_[1] - Qref))       # what the all-reduce must produce
            # exchange: pack on each handle, sum the two delta buffers on the host, write back, apply
            import torch
            deltas = []
            for e in engs:
                e.q_delta_pack()
                e.sync()
                ptr, nbytes = e.device_buffer(BUF_Q_DELTA)
                t = torch.as_tensor(sharding._DevAlias(ptr, nbytes), device="cuda:0")
                deltas.append(t)
            total = deltas[0] + deltas[1]
            torch.cuda.synchronize()
            for e, t in zip(engs, deltas):
                t.copy_(total)
                torch.cuda.synchronize()
                e.q_delta_apply()
            q0, q1 = engs[0].get_factors()[1], engs[1].get_factors()[1]
            assert np.array_equal(q0, q1)
            assert np.allclose(q0, Qref, rtol=1e-6, atol=1e-8)
    finally:
        for e in engs:
            e.close()


def _rank_main(rank, world, port, out):
    import torch
    import torch.distributed as dist
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", rank))
    log, P, Q, shards = _shards(world)
    s = shards[rank]
    eng = Engine(rank)
    eng.set_interactions(s["m_local"], log.n, s["ev_indptr"], s["ev_items"], s["uq_indptr"], s["uq_items"],
                         user_begin=s["user_begin"], event_base=s["event_base"])
    eng.set_factors(P[s["user_begin"]:s["user_begin"] + s["m_local"]], Q)
    tr = sharding.ShardedTrainer(eng, dist, torch.device("cuda", rank))
    for ep in range(2):
        tr.epoch(0.02, 0.01, 0.01, 5, ep, MODE_SERIAL)
    Pl, Ql = eng.get_factors()
    gathered = [None] * world
    dist.all_gather_object(gathered, (Pl, Ql))
    if rank == 0:
        out["res"] = gathered
    eng.close()
    dist.destroy_process_group()


def test_two_gpu_nccl_matches_two_handle_emulation():
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    import torch.multiprocessing as mp
    with socket.socket() as sk:
        sk.bind(("127.0.0.1", 0))
        port = sk.getsockname()[1]
    out = mp.Manager().dict()
    mp.spawn(_rank_main, args=(2, port, out), nprocs=2, join=True)
    (P0, Q0), (P1, Q1) = out["res"]
    assert np.array_equal(Q0, Q1)
    # same schedule on one GPU, serial epochs per shard from the same snapshot
    log, P, Q, shards = _shards(2)
    Qref = Q.copy()
    eng = Engine(0)
    try:
        Pl = [P[s["user_begin"]:s["user_begin"] + s["m_local"]].copy() for s in shards]
        for ep in range(2):
            loc = []
            for r, s in enumerate(shards):
                eng.set_interactions(s["m_local"], log.n, s["ev_indptr"], s["ev_items"], s["uq_indptr"], s["uq_items"],
                                     user_begin=s["user_begin"], event_base=s["event_base"])
                eng.set_factors(Pl[r], Qref)
                eng.bpr_epoch(0.02, 0.01, 0.01, 5, ep, MODE_SERIAL)
                Pl[r], q = eng.get_factors()
                loc.append(q)
            Qref = Qref + ((loc[0] - Qref) + (loc[1] - Qref))
    finally:
        eng.close()
    assert np.allclose(Q0, Qref, rtol=1e-6, atol=1e-8)
    assert np.allclose(P0, Pl[0], rtol=1e-6, atol=1e-8) and np.allclose(P1, Pl[1], rtol=1e-6, atol=1e-8)


# ---- WRMF: row-sharded half-sweeps over NCCL (yue_b200/sharding.py: WrmfShardedTrainer) --------------------------
def _wrmf_rank_main(rank, world, port, out):
    import torch
    import torch.distributed as dist
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", rank))
    log = synth.power_law_log(3000, 900, 150000, seed=12)
    P, Q = synth.init_factors(log.m, log.n, 40, seed=13)
    eng = Engine(rank)
    eng.set_interactions(log.m, log.n, log.ev_indptr, log.ev_items, log.uq_indptr, log.uq_items)
    eng.set_factors(P * 10, Q * 10)
    tr = sharding.WrmfShardedTrainer(eng, dist, torch.device("cuda", rank), log.uq_indptr, eng.wrmf_pair_counts()[1])
    losses = [tr.iteration(1.0) for _ in range(2)]
    X, Y = eng.get_factors()
    gathered = [None] * world
    dist.all_gather_object(gathered, (X, Y, losses))
    if rank == 0:
        out["res"] = gathered
    eng.close()
    dist.destroy_process_group()


def test_two_gpu_nccl_wrmf_is_bit_identical_to_one_gpu():
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    import torch.multiprocessing as mp
    with socket.socket() as sk:
        sk.bind(("127.0.0.1", 0))
        port = sk.getsockname()[1]
    out = mp.Manager().dict()
    mp.spawn(_wrmf_rank_main, args=(2, port, out), nprocs=2, join=True)
    (X0, Y0, l0), (X1, Y1, l1) = out["res"]
    assert np.array_equal(X0, X1) and np.array_equal(Y0, Y1) and l0 == l1
    log = synth.power_law_log(3000, 900, 150000, seed=12)
    P, Q = synth.init_factors(log.m, log.n, 40, seed=13)
    eng = Engine(0)
    try:
        eng.set_interactions(log.m, log.n, log.ev_indptr, log.ev_items, log.uq_indptr, log.uq_items)
        eng.set_factors(P * 10, Q * 10)
        ref = []
        for _ in range(2):
            ref.append(eng.wrmf_sweep(0, 1.0, 10.0, want_loss=True))
            eng.wrmf_sweep(1, 1.0, 10.0)
        X, Y = eng.get_factors()
    finally:
        eng.close()
    assert np.array_equal(X0, X) and np.array_equal(Y0, Y)
    assert l0 == pytest.approx(ref, rel=1e-12)
