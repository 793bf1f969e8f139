#!/usr/bin/env python
"""Recall@10 / NDCG@10 of the user-sharded multi-GPU trainer on the C2 quality log (the log, the factor
initialisation and the reference numbers of tools/quality_study.py: serial order 0.0978 / 0.0791).

    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 tools/quality_study_mgpu.py \
        [users tracks plays d epochs] [--sub-epochs S,S,...]

Every rank generates the same log (same seed), keeps users rank, rank+N, rank+2N, ... (yue_b200.sharding.
interleaved_users; --contiguous = event-balanced contiguous ranges, which fail), trains with ShardedTrainer (Q replicated, dQ all-reduced S times per epoch), ranks its own
test users against the final Q and the metric sums are all-reduced."""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from yue_b200 import sharding, synth  # noqa: E402
from yue_b200.engine import MODE_HOGWILD, RANK_AUTO, Engine  # noqa: E402


def main():
    argv = [a for a in sys.argv[1:] if not a.startswith("--")]
    subs = [1]
    for a in sys.argv[1:]:
        if a.startswith("--sub-epochs"):
            subs = [int(x) for x in a.split("=")[1].split(",")]
    users, tracks, plays, d, epochs = (int(x) for x in argv[:5]) if len(argv) >= 5 else (1_000_000, 200_000, 62_500_000, 64, 4)
    rank, world, local = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    log = synth.power_law_log_torch(users, tracks, plays, seed=33, test_ratio=0.2, device="cuda")
    torch.cuda.empty_cache()
    P, Q = synth.init_factors(log.m, log.n, d, seed=5)
    contiguous = "--contiguous" in sys.argv
    if contiguous:          # the sharding that does NOT work (kept for the record)
        bounds = sharding.shard_users_by_events(log.ev_indptr, world)
        mine = np.arange(bounds[rank], bounds[rank + 1], dtype=np.int64)
    else:
        mine = sharding.interleaved_users(log.m, world, rank)
    sh = sharding.local_shard_of_users(log.ev_indptr, log.ev_items, log.uq_indptr, log.uq_items, mine)
    m_local = sh["m_local"]
    te = sharding.local_shard_of_users(log.test_indptr, log.test_items, log.test_indptr, log.test_items, mine)
    te_indptr, te_items = te["ev_indptr"], te["ev_items"]
    if rank == 0:
        print("log: %d users x %d tracks, %d train events, %d ranks, d=%d, %d epochs, lr 0.02" %
              (log.m, log.n, log.train_size, world, d, epochs), flush=True)
    seed = 99
    for a in sys.argv[1:]:
        if a.startswith("--seed"):
            seed = int(a.split("=")[1])
    kappas = [None]
    for a in sys.argv[1:]:
        if a.startswith("--kappa"):
            kappas = [None if x == "none" else float(x) for x in a.split("=")[1].split(",")]
    counts = np.bincount(log.ev_items, minlength=log.n)          # the same global log on every rank
    for S, kappa in [(S, k) for S in subs for k in kappas]:
        eng = Engine(local)
        eng.set_interactions(m_local, log.n, sh["ev_indptr"], sh["ev_items"], sh["uq_indptr"], sh["uq_items"])
        if hasattr(eng, "set_event_offsets"):
            eng.set_event_offsets(sh["event_offsets"])
        eng.set_factors(np.ascontiguousarray(P[mine]), Q)
        w = None if kappa is None else sharding.saturation_weights(counts, world, S, kappa)
        trainer = sharding.ShardedTrainer(eng, dist, torch.device("cuda", local), sub_epochs=S, row_weights=w) if world > 1 else None
        loss = 0.0
        for ep in range(epochs):
            l = trainer.epoch(0.02, 0.01, 0.01, seed, ep, MODE_HOGWILD, want_loss=True) if trainer else \
                eng.bpr_epoch(0.02, 0.01, 0.01, seed, ep, MODE_HOGWILD)
            loss = l
        test_users = np.nonzero(np.diff(te_indptr) > 0)[0].astype(np.int32)
        eng.rank_topn(test_users, 10, RANK_AUTO)
        eng.set_test_set(te_indptr, te_items)
        sums, _ = eng.rank_metrics([10])
        t = torch.tensor([sums[0, 1], sums[0, 3], float(len(test_users)), loss], device="cuda", dtype=torch.float64)
        if world > 1:
            dist.all_reduce(t)
        q0 = float(np.linalg.norm(eng.get_factors()[1][0]))
        eng.close()
        if rank == 0:
            r, n = float(t[0] / t[2]), float(t[1] / t[2])
            print("seed %d " % seed + "%s %d rank(s), %d exchange(s) per epoch, kappa %s: recall@10 %.4f (%+.4f vs serial)  ndcg@10 %.4f (%+.4f)  last epoch loss %.1f  |Q[0]| %.3f"
                  % ("contiguous" if contiguous else "interleaved", world, S, kappa, r, r - 0.0978, n, n - 0.0791, float(t[3]), q0), flush=True)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
