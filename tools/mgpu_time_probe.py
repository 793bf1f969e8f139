#!/usr/bin/env python
"""Where the time of a SharedHotTrainer epoch goes (config C2's shard per rank, weak scaling): sweeps the asynchrony bound,
the parts per epoch and the reduction (torch NCCL on the second stream / none at all) and prints ms per epoch, max over ranks.

    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 tools/mgpu_time_probe.py \
        [--cases A:S:reduce,...]     reduce = nccl | none   (none: the exchange kernels run, the all-reduce does not)
"""
import argparse
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from yue_b200 import quality, sharding, synth  # noqa: E402
from yue_b200.engine import MODE_HOGWILD, Engine  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--cases", default="1:32:nccl,2:32:nccl,1:1:nccl,1:32:none,2:32:none,2:8:nccl")
    ap.add_argument("--size", default="1000000,200000,50000000,64")
    ap.add_argument("--reserve-sms", type=int, default=8)
    ap.add_argument("--epochs", type=int, default=4)
    args = ap.parse_args()
    rank, world, local = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
    os.environ.setdefault("NCCL_MAX_NCHANNELS", "8")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist.init_process_group("nccl", device_id=dev)
    users, tracks, plays, d = (int(x) for x in args.size.split(","))
    log = synth.power_law_log_torch(users, tracks, plays, 20260103 + rank, device="cuda")
    torch.cuda.empty_cache()
    P, _ = synth.init_factors(log.m, 1, d, 5 + rank)
    _, Q = synth.init_factors(1, log.n, d, 4)
    T = log.train_size
    eng = Engine(local)
    eng.set_interactions(log.m, log.n, log.ev_indptr, log.ev_items, log.uq_indptr, log.uq_items, user_begin=rank * log.m, event_base=rank * T)
    counts = np.bincount(log.ev_items, minlength=log.n)
    ctl = sharding.TorchCtl(dist, dev)
    rf = quality.torch_reduce_factory(dist, dev)

    def allmax(x):
        t = torch.tensor([x], device="cuda", dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    eng.set_factors(P, Q)
    for k in range(3):
        eng.sync(); eng.timer_start()
        eng.bpr_epoch(0.02, 0.01, 0.01, 7, k, MODE_HOGWILD, want_loss=False)
        ms = allmax(eng.timer_stop())
    if rank == 0:
        print("plain one-GPU epoch of the rank's shard (%d events): %.2f ms" % (T, ms), flush=True)
    for case in args.cases.split(","):
        A, S, red = case.split(":")
        eng.set_factors(P, Q)
        reduce = rf(eng) if red == "nccl" else (lambda e: None)
        tr = sharding.SharedHotTrainer(eng, ctl, counts, sub_epochs=int(S), asynchrony=float(A), reduce=reduce, reserve_sms=args.reserve_sms)
        times = []
        for ep in range(args.epochs):
            eng.sync(); dist.barrier()
            eng.timer_start()
            tr.epoch(0.02, 0.01, 0.01, 7, ep)
            tr.finalize()
            times.append(allmax(eng.timer_stop()))
        if rank == 0:
            print("asynchrony %s, %2s parts, reduce %-4s: %d warps on %d CTAs per rank, %d hot rows (%.0f %% of the positives): ms/epoch %s -> %.3e triplets/s"
                  % (A, S, red, tr.n_warps, tr.n_ctas, len(tr.hot_tracks), 100 * tr.hot_share_of_events, " ".join("%.2f" % t for t in times),
                     T * world / (min(times[1:]) * 1e-3)), flush=True)
        tr.close()
    eng.close()
    dist.barrier()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
