"""User-sharded multi-GPU BPR: one process per GPU, each rank owns a
contiguous range of users (its rows of P and all their events), Q is replicated, and once per
sub-epoch the ranks reconcile it:  Q <- Q_snapshot + sum_ranks (Q_rank - Q_snapshot).

The reference has no counterpart (no collective anywhere).  Plumbing is
torch.distributed; the packing / applying of the deltas and the epoch are CUDA kernels of
libyue_b200.so.  The arithmetic of the exchange is backend-agnostic and is what the CPU tests
(gloo, world_size 2) cover with `reconcile_q`.
"""
import numpy as np


def shard_users_by_events(ev_indptr, world):
    """Contiguous user ranges with (nearly) equal event counts: returns world+1 user boundaries.
    Power-law aware -- a prefix sum over events, not over users."""
    ev_indptr = np.asarray(ev_indptr, dtype=np.int64)
    m, T = len(ev_indptr) - 1, int(ev_indptr[-1])
    bounds = [0]
    for r in range(1, world):
        target = T * r // world
        u = int(np.searchsorted(ev_indptr, target, side="left"))
        bounds.append(min(max(u, bounds[-1]), m))
    bounds.append(m)
    return np.asarray(bounds, dtype=np.int64)


def local_shard(ev_indptr, ev_items, uq_indptr, uq_items, bounds, rank):
    """Rebased CSR slices of rank's users + (user_begin, event_base) for yue_set_interactions_shard."""
    u0, u1 = int(bounds[rank]), int(bounds[rank + 1])
    e0, e1 = int(ev_indptr[u0]), int(ev_indptr[u1])
    q0, q1 = int(uq_indptr[u0]), int(uq_indptr[u1])
    return dict(m_local=u1 - u0, user_begin=u0, event_base=e0,
                ev_indptr=np.ascontiguousarray(ev_indptr[u0:u1 + 1] - e0), ev_items=np.ascontiguousarray(ev_items[e0:e1]),
                uq_indptr=np.ascontiguousarray(uq_indptr[u0:u1 + 1] - q0), uq_items=np.ascontiguousarray(uq_items[q0:q1]))


def interleaved_users(m, world, rank):
    """Users rank, rank + world, rank + 2*world, ... -- the sharding that keeps every rank at the SAME position of
    the reference's user stream (recommender/cf/BPR.py:42) at the same time.  Contiguous ranges do not: with the
    heavy users of a power-law log on one rank and the light ones on another, the light users train for a whole
    epoch against a Q that has not seen the heavy users' updates, which the serial order applies FIRST
    (measured at config C2 on 2 GPUs, profiles/quality_study_r1.md section E: Recall@10 0.0004 instead of 0.098)."""
    return np.arange(rank, m, world, dtype=np.int64)


def local_shard_of_users(ev_indptr, ev_items, uq_indptr, uq_items, users):
    """Rebased CSR of an arbitrary increasing list of users (their events in the original order) and, per local
    user, the offset that turns a local event index into the global one (for yue_set_event_offsets: the sampler
    stream is a function of the GLOBAL event index, so the negatives do not depend on the sharding)."""
    ev_indptr, uq_indptr = np.asarray(ev_indptr, dtype=np.int64), np.asarray(uq_indptr, dtype=np.int64)
    users = np.asarray(users, dtype=np.int64)

    def take(indptr, items):
        deg = indptr[users + 1] - indptr[users]
        loc = np.zeros(len(users) + 1, dtype=np.int64)
        np.cumsum(deg, out=loc[1:])
        delta = indptr[users] - loc[:-1]
        idx = np.repeat(delta, deg) + np.arange(int(loc[-1]), dtype=np.int64)
        return loc, np.ascontiguousarray(np.asarray(items)[idx]), delta
    ev_loc, ev_it, ev_delta = take(ev_indptr, ev_items)
    uq_loc, uq_it, _ = take(uq_indptr, uq_items)
    return dict(m_local=len(users), users=users, ev_indptr=ev_loc, ev_items=ev_it, uq_indptr=uq_loc, uq_items=uq_it,
                event_offsets=np.ascontiguousarray(ev_delta))


def reconcile_q(q_local, q_snapshot, all_reduce_sum):
    """Q <- snapshot + sum over ranks of (Q_rank - snapshot).  `all_reduce_sum(t)` sums a tensor over
    the ranks in place.  Works on any torch tensors (CPU for the gloo tests, the library's device
    buffers on the GPU path, where the two elementwise steps are the q_delta_* kernels instead)."""
    delta = q_local - q_snapshot
    all_reduce_sum(delta)
    q_new = q_snapshot + delta
    return q_new


def saturation_weights(global_counts, world, sub_epochs, kappa):
    """Per-track factor for the summed Q deltas.  A row touched c times between two exchanges on every one of G
    ranks moves, on each rank, towards the same equilibrium: with a per-touch contraction exp(-kappa) the ranks'
    deltas are each (1 - a)(q* - q), a = exp(-kappa c), while the serial order would have moved the row by
    (1 - a^G)(q* - q).  The sum of the deltas is therefore scaled by (1 - a^G) / (G (1 - a)): 1 for rarely played
    tracks (the deltas are independent and add up), 1/G for the most played ones (every rank has already done
    the whole move; summing would overshoot G times -- measured, profiles/quality_study_r1.md section E)."""
    c = np.asarray(global_counts, dtype=np.float64) / float(world * sub_epochs)
    a = np.exp(-kappa * c)
    with np.errstate(invalid="ignore", divide="ignore"):
        w = (1.0 - a ** world) / (world * (1.0 - a))
    w[~np.isfinite(w)] = 1.0
    w[c * kappa < 1e-6] = 1.0
    return w.astype(np.float32)


class _DevAlias:
    """Expose a device buffer of the library as a torch tensor (no copy)."""

    def __init__(self, ptr, nbytes):
        self.__cuda_array_interface__ = {"shape": (nbytes // 4,), "typestr": "<f4", "data": (ptr, False), "version": 2}


class ShardedTrainer:
    """One rank of the user-sharded trainer.  `engine` already holds this rank's shard
    (Engine.set_interactions(..., user_begin, event_base)) and factors (local P rows, full Q)."""

    def __init__(self, engine, dist, device, sub_epochs=1, row_weights=None):
        import torch
        self.sub_epochs = int(sub_epochs)
        engine.set_delta_weights(row_weights)            # applied by the library's q_delta_apply kernel
        from ._lib import BUF_Q_DELTA
        self.eng, self.dist, self.torch = engine, dist, torch
        engine.q_snapshot()
        ptr, nbytes = engine.device_buffer(BUF_Q_DELTA)
        self.delta = torch.as_tensor(_DevAlias(ptr, nbytes), device=device)
        self.stream = torch.cuda.ExternalStream(engine.stream_ptr(), device=device)

    def exchange(self):
        """Q <- snapshot + sum over ranks of (Q_rank - snapshot); the new Q is the next snapshot."""
        self.eng.q_delta_pack()
        with self.torch.cuda.stream(self.stream):        # NCCL ordered on the library's stream
            self.dist.all_reduce(self.delta)
        self.eng.q_delta_apply()                         # Q = snapshot + w * sum of deltas (w: saturation_weights)

    def epoch(self, lr, regU, regI, seed, epoch, mode, want_loss=False):
        """One epoch = sub_epochs parts of the local users' work, the ranks reconciling Q after each part
        (more exchanges = less divergence between the replicas of Q, each costs one all-reduce
        of n*d floats: 51 MB at C2, 1 GB at C3)."""
        if self.sub_epochs == 1:
            loss = self.eng.bpr_epoch(lr, regU, regI, seed, epoch, mode, want_loss=want_loss)
            self.exchange()
            return loss
        loss = 0.0
        for part in range(self.sub_epochs):
            l = self.eng.bpr_epoch_part(lr, regU, regI, seed, epoch, part, self.sub_epochs, mode, want_loss=want_loss)
            loss += l if want_loss else 0.0
            self.exchange()
        return loss if want_loss else None


# ---- WRMF: rows of a half-sweep are independent, so the multi-GPU form is exact -------------
def exchange_rows(table, bounds, dist):
    """Every rank has solved rows [bounds[r], bounds[r+1]) of `table` ([rows, ld] tensor, same shape on every rank) in
    place; afterwards every rank holds every solved row.  One broadcast per rank (ranges are balanced by entries, not
    by rows, so their sizes differ).  Backend-agnostic: gloo on CPU tensors in the tests, NCCL on the library's device
    buffers on the GPU path."""
    for r in range(len(bounds) - 1):
        lo, hi = int(bounds[r]), int(bounds[r + 1])
        if hi > lo:
            dist.broadcast(table[lo:hi], src=r)


class WrmfShardedTrainer:
    """One rank of row-sharded WRMF.  Every rank holds the whole log and both tables (Engine.set_interactions /
    set_factors with the same arrays); per half-sweep a rank solves its range of rows -- ranges balanced by entries
    (shard_users_by_events on the row pointers) -- and the ranks exchange the solved rows.  No reduction is involved, so
    the tables are bit-identical to the single-GPU sweep whatever the number of ranks."""

    def __init__(self, engine, dist, device, uq_indptr, it_indptr):
        import torch
        from ._lib import BUF_P, BUF_Q
        self.eng, self.dist, self.torch = engine, dist, torch
        self.rank, world = dist.get_rank(), dist.get_world_size()
        self.user_bounds = shard_users_by_events(uq_indptr, world)
        self.track_bounds = shard_users_by_events(it_indptr, world)
        ld = (engine.k + 3) & ~3
        self.tables = []
        for which, rows in ((BUF_P, engine.m), (BUF_Q, engine.n)):
            ptr, nbytes = engine.device_buffer(which)
            self.tables.append(torch.as_tensor(_DevAlias(ptr, nbytes), device=device)[:rows * ld].view(rows, ld))
        self.stream = torch.cuda.ExternalStream(engine.stream_ptr(), device=device)

    def iteration(self, reg, alpha=10.0, want_loss=True):
        """WRMF.py:34-83 over all ranks: users, exchange, tracks, exchange.  Returns the loss summed over the ranks."""
        r = self.rank
        loss = self.eng.wrmf_sweep_rows(0, self.user_bounds[r], self.user_bounds[r + 1], reg, alpha, want_loss=want_loss)
        with self.torch.cuda.stream(self.stream):
            exchange_rows(self.tables[0], self.user_bounds, self.dist)
        self.eng.sync()
        self.eng.wrmf_sweep_rows(1, self.track_bounds[r], self.track_bounds[r + 1], reg, alpha)
        with self.torch.cuda.stream(self.stream):
            exchange_rows(self.tables[1], self.track_bounds, self.dist)
            if want_loss:
                lt = self.torch.tensor([loss], dtype=self.torch.float64, device=self.tables[0].device)
                self.dist.all_reduce(lt)
                loss = float(lt.item())
        self.eng.sync()
        return loss if want_loss else None
