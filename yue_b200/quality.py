"""End-to-end quality check of a training schedule (north_star check 4): Recall@10 / NDCG@10 of the tables a schedule
produces against the tables of the SERIAL-ORDER run on the same log, same initial tables and same sampler seed (the serial
mode reproduces the reference's loop, recommender/cf/BPR.py:40-62, to 1e-5 -- tests/test_bpr_gpu.py).  The gate is 0.5
points absolute on both.  Used by bench.py's `quality` block, tests/test_multigpu.py and tools/quality_mgpu.py.

Everything here runs on the GPU through the C-ABI (training: K1/K2, ranking: K3 exact, metrics: K6); nothing is computed on
the CPU and the oracle is not involved (the serial mode is the product's own parity mode, pinned elsewhere)."""
import time

import numpy as np

from . import sharding, synth
from .engine import MODE_HOGWILD, MODE_SERIAL, RANK_AUTO, Engine

GATE = 0.005
QUALITY_LOG = dict(users=100_000, tracks=20_000, plays=5_000_000, d=64, epochs=4, seed=33, init_seed=5, sampler_seed=99,
                   lr=0.02, reg_u=0.01, reg_i=0.01)


def make_log(spec, device="cuda"):
    log = synth.power_law_log_torch(spec["users"], spec["tracks"], spec["plays"], seed=spec["seed"], test_ratio=0.2, device=device)
    P, Q = synth.init_factors(log.m, log.n, spec["d"], seed=spec["init_seed"])
    return log, P, Q


def _metrics(eng, te_indptr, te_items):
    """(sum of per-user recall@10, sum of per-user NDCG@10, test users) of the handle's tables over its local users."""
    users = np.nonzero(np.diff(te_indptr) > 0)[0].astype(np.int32)
    if len(users) == 0:
        return 0.0, 0.0, 0
    eng.rank_topn(users, 10, RANK_AUTO)
    eng.set_test_set(te_indptr, te_items)
    sums, _ = eng.rank_metrics([10])
    return float(sums[0, 1]), float(sums[0, 3]), int(len(users))


def single_gpu_run(device, log, P, Q, spec, mode):
    """Train the whole log on one handle in `mode`; returns (recall@10, ndcg@10, seconds, last epoch's -log loss)."""
    eng = Engine(device)
    try:
        eng.set_interactions(log.m, log.n, log.ev_indptr, log.ev_items, log.uq_indptr, log.uq_items)
        eng.set_factors(P, Q)
        t0 = time.perf_counter()
        for ep in range(spec["epochs"]):
            loss = eng.bpr_epoch(spec["lr"], spec["reg_u"], spec["reg_i"], spec["sampler_seed"], ep, mode)
        dt = time.perf_counter() - t0
        r, n, cnt = _metrics(eng, log.test_indptr, log.test_items)
    finally:
        eng.close()
    return r / max(cnt, 1), n / max(cnt, 1), dt, loss


def shared_hot_run(device, ctl, log, P, Q, spec, sub_epochs, asynchrony, reduce_factory=None, reserve_sms=8, apr=None, kappa=None,
                   hot_max=248, hot_div=4096):
    """This rank's part of one run of SharedHotTrainer on `log` (every rank holds the same log and initial tables and keeps
    the users rank, rank + world, ...).  kappa: None = plain sum of the tail deltas, else sharding.saturation_weights with
    that per-touch contraction.  Returns a dict on every rank (metric sums are all-reduced through ctl)."""
    rank, world = ctl.rank, ctl.world
    sub_epochs = int(sub_epochs) if sub_epochs else sharding.default_sub_epochs(world)       # 0 / None: the trainer's defaults
    asynchrony = float(asynchrony) if asynchrony else sharding.default_asynchrony(world)
    mine = sharding.interleaved_users(log.m, world, rank)
    sh = sharding.local_shard_of_users(log.ev_indptr, log.ev_items, log.uq_indptr, log.uq_items, mine)
    te = sharding.local_shard_of_users(log.test_indptr, log.test_items, log.test_indptr, log.test_items, mine)
    eng = Engine(device)
    try:
        eng.set_interactions(sh["m_local"], log.n, sh["ev_indptr"], sh["ev_items"], sh["uq_indptr"], sh["uq_items"])
        eng.set_event_offsets(sh["event_offsets"])           # the sampler stream of the unsharded log
        eng.set_factors(np.ascontiguousarray(P[mine]), Q)
        local_counts = np.bincount(sh["ev_items"], minlength=log.n)
        reduce = reduce_factory(eng) if (reduce_factory is not None and world > 1) else None
        w = None
        if kappa is not None and world > 1:
            w = sharding.saturation_weights(np.bincount(log.ev_items, minlength=log.n), world, sub_epochs, kappa)
        tr = sharding.SharedHotTrainer(eng, ctl, local_counts, sub_epochs=sub_epochs, asynchrony=asynchrony, reduce=reduce,
                                       reserve_sms=reserve_sms, row_weights=w, hot_max=hot_max, hot_div=hot_div)
        eng.sync()
        ctl.barrier()
        t0 = time.perf_counter()
        loss = 0.0
        for ep in range(spec["epochs"]):
            loss = tr.epoch(spec["lr"], spec["reg_u"], spec["reg_i"], spec["sampler_seed"], ep, want_loss=True, apr=apr)
        tr.finalize()
        dt = time.perf_counter() - t0
        r, n, cnt = _metrics(eng, te["ev_indptr"], te["ev_items"])
        tot = ctl.allreduce_sum(np.array([r, n, float(cnt), loss], dtype=np.float64))
        out = dict(recall=float(tot[0] / max(tot[2], 1.0)), ndcg=float(tot[1] / max(tot[2], 1.0)), test_users=int(tot[2]),
                   last_epoch_loss=float(tot[3]), seconds=dt, ranks=world, sub_epochs=int(sub_epochs), asynchrony=float(asynchrony),
                   warps_per_rank=int(tr.n_warps), ctas_per_rank=int(tr.n_ctas), hot_tracks=int(len(tr.hot_tracks)),
                   hot_share_of_events=float(tr.hot_share_of_events))
        tr.close()
    except Exception as exc:                            # a diverged run (NaN loss) is a result, not a crash -- on every rank alike
        if "NaN" not in str(exc):
            raise
        out = dict(recall=float("nan"), ndcg=float("nan"), test_users=0, last_epoch_loss=float("nan"), seconds=float("nan"), ranks=world,
                   sub_epochs=int(sub_epochs), asynchrony=float(asynchrony), warps_per_rank=0, ctas_per_rank=0, hot_tracks=0,
                   hot_share_of_events=0.0, diverged=True)
    finally:
        eng.close()
    return out


def verdict(run, base_recall, base_ndcg):
    """Adds the deltas against the serial-order run and the gate's verdict to a run's dict."""
    run = dict(run)
    run["d_recall"] = run["recall"] - base_recall
    run["d_ndcg"] = run["ndcg"] - base_ndcg
    run["gate"] = GATE
    run["in_gate"] = bool(abs(run["d_recall"]) < GATE and abs(run["d_ndcg"]) < GATE)
    return run


def torch_reduce_factory(dist, device):
    """reduce(engine) for SharedHotTrainer: torch.distributed's all-reduce of the handle's delta buffer, ordered on the
    handle's SECOND stream (NCCL waits for the pack, the finish kernel waits for NCCL; the epoch stream runs on)."""
    import torch
    from ._lib import BUF_Q_DELTA

    def factory(eng):
        ptr, nbytes = eng.device_buffer(BUF_Q_DELTA)
        delta = torch.as_tensor(sharding._DevAlias(ptr, nbytes), device=device)
        stream = torch.cuda.ExternalStream(eng.stream2_ptr(), device=device)

        def reduce(_eng):
            with torch.cuda.stream(stream):
                dist.all_reduce(delta)
        return reduce
    return factory
