#!/usr/bin/env python
"""Recall@10 / NDCG@10 of SharedHotTrainer (hot rows shared over NVLink, tail exchanged one part late) against the
serial-order run of the same log, on 1..8 GPUs, sweeping sub-epochs and the asynchrony bound.

    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 tools/quality_mgpu.py \
        [--size users,tracks,plays] [--epochs E] [--sub-epochs 8,32] [--asynchrony 1,2,4] [--no-serial r,n] [--time-c2]

Every rank generates the same log (same seed) and keeps the users rank, rank+N, ...; rank 0 also trains the serial order
(the reference's loop order) unless --no-serial gives its numbers.  One line per configuration; see yue_b200/quality.py."""
import argparse
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from yue_b200 import quality, sharding  # noqa: E402
from yue_b200.engine import MODE_HOGWILD, MODE_SERIAL  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--size", default="")
    ap.add_argument("--epochs", type=int, default=0)
    ap.add_argument("--sub-epochs", default="32")
    ap.add_argument("--asynchrony", default="1")
    ap.add_argument("--no-serial", default="")
    ap.add_argument("--reserve-sms", type=int, default=8)
    ap.add_argument("--kappa", default="none", help="comma list: none = plain sum of the tail deltas, else the per-touch contraction of saturation_weights")
    ap.add_argument("--hot", default="248:4096", help="comma list of hot_max:hot_div -- the shared table holds up to hot_max tracks that are the positive of more than 1/hot_div of the events")
    ap.add_argument("--hogwild-1gpu", action="store_true", help="rank 0 also trains the plain one-GPU Hogwild epochs")
    args = ap.parse_args()
    rank, world, local = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
    os.environ.setdefault("NCCL_MAX_NCHANNELS", "8")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist.init_process_group("nccl" if world > 1 else "gloo", device_id=dev if world > 1 else None,
                            init_method=None if "MASTER_ADDR" in os.environ else "tcp://127.0.0.1:29511",
                            rank=rank, world_size=world)
    spec = dict(quality.QUALITY_LOG)
    if args.size:
        spec["users"], spec["tracks"], spec["plays"] = (int(x) for x in args.size.split(","))
    if args.epochs:
        spec["epochs"] = args.epochs
    log, P, Q = quality.make_log(spec)
    torch.cuda.empty_cache()
    say = (lambda s: print(s, flush=True)) if rank == 0 else (lambda s: None)
    share = np.bincount(log.ev_items, minlength=log.n).max() / log.train_size
    say("log: %d users x %d tracks, %d train events (hottest track %.1f %%), d=%d, %d epochs, lr %.3f, %d rank(s)"
        % (log.m, log.n, log.train_size, 100 * share, spec["d"], spec["epochs"], spec["lr"], world))
    if args.no_serial:
        br, bn = (float(x) for x in args.no_serial.split(","))
    else:
        base = [0.0, 0.0]
        if rank == 0:
            r, n, dt, loss = quality.single_gpu_run(local, log, P, Q, spec, MODE_SERIAL)
            base = [r, n]
            say("serial order (1 warp, %.1f s): recall@10 %.4f ndcg@10 %.4f last-epoch loss %.1f" % (dt, r, n, loss))
        t = torch.tensor(base, dtype=torch.float64, device=dev if world > 1 else "cpu")
        dist.broadcast(t, 0)
        br, bn = float(t[0]), float(t[1])
    if args.hogwild_1gpu and rank == 0:
        r, n, dt, loss = quality.single_gpu_run(local, log, P, Q, spec, MODE_HOGWILD)
        say("one-GPU Hogwild epochs (%.2f s): recall@10 %.4f (%+.4f) ndcg@10 %.4f (%+.4f) loss %.1f" % (dt, r, r - br, n, n - bn, loss))
    ctl = sharding.TorchCtl(dist, dev if world > 1 else None)
    rf = quality.torch_reduce_factory(dist, dev) if world > 1 else None
    for S in (int(x) for x in args.sub_epochs.split(",")):
        for A, K, HM, HD in [(float(x), None if k == "none" else float(k), int(hh.split(":")[0]), int(hh.split(":")[1]))
                             for x in args.asynchrony.split(",") for k in args.kappa.split(",") for hh in args.hot.split(",")]:
            run = quality.verdict(quality.shared_hot_run(local, ctl, log, P, Q, spec, S, A, reduce_factory=rf, reserve_sms=args.reserve_sms, kappa=K,
                                                         hot_max=HM, hot_div=HD), br, bn)
            say("kappa %s: " % K + "%d rank(s), %2d parts/epoch, asynchrony %.2f (%d warps on %d CTAs per rank, %d hot rows = %.0f %% of the events): "
                "recall@10 %.4f (%+.4f) ndcg@10 %.4f (%+.4f) %s  loss %.1f  %.2f s = %.3e triplets/s"
                % (world, S, A, run["warps_per_rank"], run["ctas_per_rank"], run["hot_tracks"], 100 * run["hot_share_of_events"],
                   run["recall"], run["d_recall"], run["ndcg"], run["d_ndcg"], "IN GATE" if run["in_gate"] else "outside",
                   run["last_epoch_loss"], run["seconds"], log.train_size * spec["epochs"] / run["seconds"]))
    dist.barrier()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
