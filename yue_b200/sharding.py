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


def shard_rows_by_cost(indptr, world, fixed=59.0, light_max=16, light_fixed=39.0, light_per_entry=0.5):
    """Contiguous row ranges of (nearly) equal WRMF cost: world+1 boundaries.  Fitted to the two sweeps of config C2 on one
    B200 (user sweep 47.0 ms: 480 K Woodbury rows at 22 ns, 520 K factorised rows; track sweep 27.1 ms: 200 K rows, 36.7 M
    entries): a factorised row costs 30.5 ns + 0.52 ns per entry -- the k x k factorisation is a latency chain that does not
    depend on the entries and weighs as much as 59 of them -- a row of 1..light_max entries (Woodbury kernel) 22 ns, an empty
    row nothing.  Balancing by entries alone hands the rank with the lightest rows several times the rows of the others: on
    8 GPUs the iteration took 23.4 ms against 74 / 8 = 9.3 (profiles/r2/bench_r2_n8_session2.json)."""
    e = np.diff(np.asarray(indptr, dtype=np.int64)).astype(np.float64)
    cost = np.where(e == 0, 0.0, np.where(e <= light_max, light_fixed + light_per_entry * e, fixed + e))
    cum = np.concatenate([[0.0], np.cumsum(cost)])
    m = len(e)
    bounds = [0]
    for r in range(1, world):
        u = int(np.searchsorted(cum, cum[-1] * r / world, side="left"))
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
    set_factors with the same arrays); per half-sweep a rank solves its range of rows -- ranges balanced by cost
    (shard_rows_by_cost: entries plus the factorisation every row pays) -- and the ranks exchange the solved rows.  No reduction is involved, so
    the tables are bit-identical to the single-GPU sweep whatever the number of ranks."""

    def __init__(self, engine, dist, device, uq_indptr, it_indptr):
        import torch
        from ._lib import BUF_P, BUF_Q
        self.eng, self.dist, self.torch = engine, dist, torch
        self.rank, world = dist.get_rank(), dist.get_world_size()
        self.user_bounds = shard_rows_by_cost(uq_indptr, world)
        self.track_bounds = shard_rows_by_cost(it_indptr, world)
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


# ---- round 2: hot rows as ONE copy over peer memory, the tail exchanged under the next sub-epoch ---------------------------
def select_hot_tracks(global_counts, hot_max=248, hot_div=4096, min_count=16384):
    """The hot set every rank must agree on: tracks that are the positive of more than 1/hot_div of ALL events and of at least
    min_count, most played first (ties: lower id).  One GPU keeps ~10 rows in its table (hot_max 24, hot_div 128: the rows whose
    L2 slice would saturate); the ranks of a sharded model share MANY more -- every row in the table exists once and is exact,
    every other row is a replica that is reconciled only at the exchanges: at config C2's size 239 shared rows (47 % of the
    positives) are what brings Recall@10 from -0.026 to inside the gate (profiles/r2/quality_c2_n2_hot.log)."""
    c = np.asarray(global_counts, dtype=np.int64)
    total = int(c.sum())
    cand = np.nonzero((c >= min_count) & (c * hot_div > total))[0]
    cand = cand[np.lexsort((cand, -c[cand]))][:hot_max]
    return cand.astype(np.int32), c[cand].astype(np.int64), total


class ThreadCtl:
    """Control plane of R ranks that are threads of one process (one Engine per GPU): the class-API path, and the tests."""

    class Shared:
        def __init__(self, world):
            import threading
            self.world, self.barrier, self.slots = world, threading.Barrier(world), [None] * world

    def __init__(self, shared, rank):
        self.s, self.rank, self.world = shared, rank, shared.world

    def barrier(self):
        self.s.barrier.wait()

    def allgather(self, obj):
        self.s.barrier.wait()
        self.s.slots[self.rank] = obj
        self.s.barrier.wait()
        out = list(self.s.slots)
        self.s.barrier.wait()
        return out

    def allreduce_sum(self, arr):
        return np.sum(self.allgather(np.asarray(arr)), axis=0)


def run_on_ranks(shared, fn):
    """fn(rank, ctl) on one thread per rank of a ThreadCtl group; returns the results in rank order.  A rank that raises
    aborts the group's barrier so that the others do not wait for it forever; the first error is re-raised."""
    import threading
    out, errs = [None] * shared.world, []

    def run(r):
        try:
            out[r] = fn(r, ThreadCtl(shared, r))
        except BaseException as exc:                    # noqa: BLE001
            errs.append(exc)
            shared.barrier.abort()
    th = [threading.Thread(target=run, args=(r,)) for r in range(shared.world)]
    for t in th:
        t.start()
    for t in th:
        t.join()
    if errs:
        first = [e for e in errs if not isinstance(e, threading.BrokenBarrierError)]
        raise (first or errs)[0]
    return out


class TorchCtl:
    """Control plane over torch.distributed (one process per GPU, NCCL or gloo)."""

    def __init__(self, dist, device=None):
        self.dist, self.rank, self.world, self.device = dist, dist.get_rank(), dist.get_world_size(), device

    def barrier(self):
        self.dist.barrier()

    def allgather(self, obj):
        out = [None] * self.world
        self.dist.all_gather_object(out, obj)
        return out

    def allreduce_sum(self, arr):
        import torch
        t = torch.from_numpy(np.ascontiguousarray(arr))
        if self.device is not None:
            t = t.to(self.device)
        self.dist.all_reduce(t)
        return t.cpu().numpy()


def default_sub_epochs(world):
    """Parts per epoch (= exchanges of the tail of Q) that kept SharedHotTrainer inside the 0.5-point gate on the fixed
    5 M-play quality log (profiles/r2/quality_n*.log, asynchrony 1 and 2): 32 on 2 GPUs, 32 / 64 on 4, 64 / 128 / 256 on 8
    (32 parts on 8 GPUs: Recall@10 -0.022).  Every rank adds its tail deltas blind to the other N - 1 ranks' for two parts
    (the sum arrives one part late), so the parts have to shrink as the ranks grow: 32 up to 2 ranks, 4 N^2 above."""
    world = int(world)
    return 32 if world <= 2 else 4 * world * world


def default_asynchrony(world):
    """All ranks together run `asynchrony` x the warps one GPU would give the whole log.  A row shared over NVLink has its
    updates in flight ~4x longer than a local one, so the same bound on updates in flight allows a
    quarter of the warps: at config C2's size 2 GPUs leave the gate at 1.0 and 0.5 (Recall@10 -0.011 / -0.010) and are inside
    at 0.25 and 0.125 (-0.0012 / -0.0005; profiles/r2/quality_c2_n2_async.log)."""
    return 1.0 if int(world) <= 1 else 0.25


class SharedHotTrainer:
    """One rank of the round-2 trainer (include/yue_b200.h "multi-GPU, round 2").

    * users are sharded (interleaved_users + set_event_offsets keep every rank at the same place of the reference's user
      stream with the reference's negatives), P rows are private;
    * the rows of the most played tracks exist ONCE: slot s of the hot-row table lives on rank s % world, every rank's epoch
      kernel loads and adds it over NVLink -- nothing to reconcile, nothing to overshoot;
    * the long tail of Q is replicated and exchanged `sub_epochs` times per epoch: delta = Q - snapshot is packed after a
      part, all-reduced on the handle's second stream WHILE the next part runs, and applied one part late
      (Q += sum - own, snapshot += sum);
    * `asynchrony` bounds how many updates are in flight over all ranks: the ranks together run asynchrony x the warps one
      GPU would give the whole log (tools/staleness_sim.py: the trajectory leaves the 0.5-point gate when the updates of
      the most played row in flight exceed ~1000; one B200 at full speed is at ~550).

    engine: holds this rank's shard and factors.  ctl: ThreadCtl / TorchCtl.  local_counts[n]: plays per track in the shard.
    reduce: a callable(engine) that all-reduces BUF_Q_DELTA on the handle's second stream (bench.py passes
    torch.distributed), or None = chosen here: ranks that are handles of ONE process sum each other's deltas over peer
    memory (yue_q_exchange_reduce_peers), ranks in different processes use the library's NCCL communicator."""

    def __init__(self, engine, ctl, local_counts, sub_epochs=None, asynchrony=None, reserve_sms=8, reduce=None, row_weights=None,
                 hot_max=248, hot_div=4096, sm_count=148, warps_per_sm=12, min_events_per_warp=16384):
        import os
        from . import engine as _eng
        self.eng, self.ctl = engine, ctl
        self.sub_epochs = int(sub_epochs) if sub_epochs else default_sub_epochs(ctl.world)
        asynchrony = float(asynchrony) if asynchrony else default_asynchrony(ctl.world)
        self.asynchrony = asynchrony
        counts = ctl.allreduce_sum(np.asarray(local_counts, dtype=np.int64))
        tracks, tcounts, total = select_hot_tracks(counts, hot_max=hot_max, hot_div=hot_div)
        self.hot_tracks, self.total_events = tracks, total
        self.hot_share_of_events = float(tcounts.sum()) / max(total, 1)
        engine.set_hot_tracks(tracks, tcounts, total)
        # every rank maps every table: a pointer inside one process, a CUDA IPC handle between processes
        ipc, ptr = engine.hot_table_export()
        peers = ctl.allgather((os.getpid(), engine.device, ipc, ptr))
        tables = []
        for r, (pid, dev, handle, p) in enumerate(peers):
            if r == ctl.rank:
                tables.append(None)
            elif pid == os.getpid():
                engine.enable_peer(dev)
                tables.append(p)
            else:
                tables.append(engine.hot_table_open(handle))
        engine.hot_share(ctl.world, ctl.rank, tables)
        engine.sync()
        ctl.barrier()                                  # every owner's rows are in its table before anyone trains
        engine.set_delta_weights(row_weights)
        engine.q_snapshot()
        # concurrency: all ranks together = asynchrony x one GPU's automatic choice for the whole log
        auto = max(1, min(sm_count * warps_per_sm, total // min_events_per_warp))
        self.n_ctas = max(1, sm_count - (reserve_sms if ctl.world > 1 else 0))
        self.n_warps = max(1, min(self.n_ctas * warps_per_sm, int(round(asynchrony * auto / ctl.world))))
        engine.set_sgd_concurrency(self.n_warps, self.n_ctas)
        if reduce is None and ctl.world > 1:
            if all(pid == os.getpid() for pid, _, _, _ in peers):
                # all ranks are handles of this process: every rank sums the ranks' packed deltas itself over peer memory
                from ._lib import BUF_Q_DELTA
                ptrs = ctl.allgather(engine.device_buffer(BUF_Q_DELTA)[0])
                deltas = [None if r == ctl.rank else p for r, p in enumerate(ptrs)]

                def reduce(e):
                    e.sync()                           # my pack is done ...
                    ctl.barrier()                      # ... and so is everybody's
                    e.q_exchange_reduce_peers(deltas)
                    ctl.barrier()                      # everybody has read my delta: the next pack may overwrite it
            else:
                uid = ctl.allgather(_eng.comm_unique_id() if ctl.rank == 0 else None)[0]
                engine.comm_init(ctl.world, ctl.rank, uid)
                reduce = lambda e: e.q_exchange_reduce()   # noqa: E731
        self.reduce = reduce
        self.pending = False

    def _exchange(self):
        if self.ctl.world == 1:
            return
        if self.pending:
            self.eng.q_exchange_finish()               # last part's sum has had a whole part to arrive
        self.eng.q_exchange_begin()
        self.reduce(self.eng)
        self.pending = True

    def epoch(self, lr, regU, regI, seed, epoch, want_loss=False, apr=None, finalize=False):
        """One epoch = sub_epochs parts; returns the local loss (sum over the parts) when want_loss.  apr = (eps, regA[, slot])
        trains the APR variant (K2a; slot = which of the positive's negatives, APR draws 3).  finalize: see `finalize`."""
        from ._lib import E_NUMERIC, MODE_HOGWILD, YueError
        loss, diverged = 0.0, False
        for part in range(self.sub_epochs):
            try:
                if apr is None:
                    l = self.eng.bpr_epoch_part(lr, regU, regI, seed, epoch, part, self.sub_epochs, MODE_HOGWILD, want_loss=want_loss)
                else:
                    l = self.eng.apr_epoch_part(lr, regU, regI, apr[0], apr[1], seed, epoch, part, self.sub_epochs,
                                                apr[2] if len(apr) > 2 else 0, MODE_HOGWILD, want_loss=want_loss)
            except YueError as exc:                    # a non-finite loss on THIS rank: keep the collectives of the epoch in step,
                if exc.code != E_NUMERIC:              # then fail on every rank together
                    raise
                l, diverged = float("nan"), True
            loss += l if want_loss else 0.0
            self._exchange()
        if want_loss and self.ctl.world > 1:
            diverged = bool(self.ctl.allreduce_sum(np.array([1 if diverged else 0], dtype=np.int64))[0] > 0)
        if diverged:
            raise YueError(E_NUMERIC, "Loss = NaN or Infinity: current settings does not fit the recommender!")
        if finalize:
            self.finalize()
        return loss if want_loss else None

    def finalize(self):
        """Apply the exchange still in flight and bring this rank's Q up to date with the shared hot rows (the handle's Q
        is then complete: yue_get_factors, yue_frob2, ranking).  Training may go on afterwards."""
        if self.pending:
            self.eng.q_exchange_finish(quiescent=True)     # nothing trained since the pack: every rank ends with the same bits
            self.pending = False
        self.eng.sync()
        self.ctl.barrier()                             # nobody is still adding to a table
        self.eng.hot_pull()
        self.ctl.barrier()                             # nobody's next epoch changes a table while another rank reads it

    def close(self):
        """Back to private per-launch tables (the handle keeps its complete Q)."""
        self.finalize()
        self.eng.hot_unshare()
        self.eng.set_sgd_concurrency(0, 0)
