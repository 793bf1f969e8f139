"""Seeded synthetic play logs shaped like the BASELINE.json configs.

The reference ships no data (``/root/reference/.gitignore:1-4``), so every workload here is
generated: user activity ``deg_u ~ rank_u^-0.8`` (min 2 plays), track popularity
``P(track = t) ~ 1/(t+1)`` (track id = popularity rank: low ids are hot, the worst case for
Hogwild conflicts), plays drawn with replacement so repeat plays stay in the log as separate
events, and a per-event Bernoulli(test_ratio) hold-out like ``tool/dataSplit.py:9-23`` followed
by the "drop test pairs already seen in training" rule of ``data/record.py:195-202``.

Two forms:

* ``power_law_log`` -> ``PlayLog`` of integer arrays (the form the C ABI consumes) for the large
  configs, never materialising Python dicts;
* ``write_csv_log`` -> a real ``time,user,track,artist`` text file so the reference-shaped loader
  (``tool/file.py:23-52``) and ``Record`` run on config C1.
"""
from dataclasses import dataclass

import numpy as np

SEED_BASE = 20260101


@dataclass
class PlayLog:
    m: int                    # users
    n: int                    # tracks
    ev_indptr: np.ndarray     # int64 [m+1]   training events per user (user-major, file order)
    ev_items: np.ndarray      # int32 [T]
    uq_indptr: np.ndarray     # int64 [m+1]   sorted unique training tracks per user
    uq_items: np.ndarray      # int32 [nnz]
    test_indptr: np.ndarray   # int64 [m+1]   sorted unique held-out tracks per user (train pairs removed)
    test_items: np.ndarray    # int32 [nt]

    @property
    def train_size(self):
        return int(self.ev_items.shape[0])

    def test_users(self):
        return np.nonzero(np.diff(self.test_indptr) > 0)[0].astype(np.int32)


def user_degrees(m, plays, alpha=0.8, min_deg=2):
    w = np.arange(1, m + 1, dtype=np.float64) ** (-alpha)
    deg = np.maximum(min_deg, np.floor(w * (plays / w.sum()))).astype(np.int64)
    # hand the rounding remainder to the heaviest users so the total is exact when possible
    rem = int(plays - deg.sum())
    if rem > 0:
        deg[:rem % m] += 1
        deg += rem // m
    return deg


def _csr_unique(users, items, m, n):
    """sorted-unique (user, item) pairs -> CSR.  users must be non-decreasing."""
    key = users.astype(np.int64) * np.int64(n) + items.astype(np.int64)
    key = np.unique(key)                         # sorts; users stay grouped
    u = key // n
    indptr = np.zeros(m + 1, dtype=np.int64)
    np.cumsum(np.bincount(u, minlength=m), out=indptr[1:])
    return indptr, (key - u * n).astype(np.int32), key


def power_law_log(m, n, plays, seed, test_ratio=0.2, alpha=0.8, beta=1.0):
    rng = np.random.Generator(np.random.Philox(seed))
    deg = user_degrees(m, plays, alpha)
    T = int(deg.sum())
    cdf = np.cumsum((np.arange(1, n + 1, dtype=np.float64)) ** (-beta))
    cdf /= cdf[-1]
    items = np.minimum(np.searchsorted(cdf, rng.random(T), side="right"), n - 1).astype(np.int32)
    users = np.repeat(np.arange(m, dtype=np.int32), deg)
    held = rng.random(T) < test_ratio if test_ratio > 0 else np.zeros(T, dtype=bool)
    tr_u, tr_i = users[~held], items[~held]
    ev_indptr = np.zeros(m + 1, dtype=np.int64)
    np.cumsum(np.bincount(tr_u, minlength=m), out=ev_indptr[1:])
    uq_indptr, uq_items, tr_key = _csr_unique(tr_u, tr_i, m, n)
    te_key = np.unique(users[held].astype(np.int64) * np.int64(n) + items[held].astype(np.int64))
    te_key = te_key[~np.isin(te_key, tr_key, assume_unique=True)]
    te_u = te_key // n
    test_indptr = np.zeros(m + 1, dtype=np.int64)
    np.cumsum(np.bincount(te_u, minlength=m), out=test_indptr[1:])
    return PlayLog(m, n, ev_indptr, np.ascontiguousarray(tr_i), uq_indptr, uq_items,
                   test_indptr, (te_key - te_u * n).astype(np.int32))


def init_factors(m, n, k, seed):
    """U[0, 0.1) float32, the distribution of base/IterativeRecommender.py:37-38."""
    rng = np.random.Generator(np.random.Philox(seed))
    P = (rng.random((m, k), dtype=np.float32) / np.float32(10)).astype(np.float32)
    Q = (rng.random((n, k), dtype=np.float32) / np.float32(10)).astype(np.float32)
    return P, Q


def mask_csr(m, n, mean_items, seed):
    """Config C4's ranking mask: ~mean_items sorted unique tracks per user."""
    rng = np.random.Generator(np.random.Philox(seed))
    deg = np.maximum(1, rng.poisson(mean_items, m)).astype(np.int64)
    users = np.repeat(np.arange(m, dtype=np.int32), deg)
    items = rng.integers(0, n, int(deg.sum()), dtype=np.int32)
    indptr, uq, _ = _csr_unique(users, items, m, n)
    return indptr, uq


def mask_csr_torch(m, n, mean_items, seed, device=None):
    """mask_csr with torch ops (GPU when present): config C4 at full size is 50 M (user, track) pairs."""
    import torch
    if device is None:
        device = "cuda" if torch.cuda.is_available() else "cpu"
    g = torch.Generator(device=device)
    g.manual_seed(int(seed))
    deg = torch.poisson(torch.full((m,), float(mean_items), device=device), generator=g).clamp_(min=1).to(torch.int64)
    users = torch.repeat_interleave(torch.arange(m, device=device, dtype=torch.int64), deg)
    items = torch.randint(0, n, (int(deg.sum().item()),), generator=g, device=device, dtype=torch.int64)
    key = torch.unique(users * n + items)
    del users, items
    u = key // n
    indptr = torch.zeros(m + 1, dtype=torch.int64, device=device)
    indptr[1:] = torch.cumsum(torch.bincount(u, minlength=m), 0)
    uq = (key - u * n).to(torch.int32)
    return indptr.cpu().numpy(), uq.cpu().numpy()


def write_csv_log(path, n_users, n_tracks, plays, seed, n_artists=500):
    """Config C1 as text: ``time,user,track,artist`` lines in shuffled (time) order so ids
    assigned by first appearance are NOT the generator's ids, like a real log."""
    rng = np.random.Generator(np.random.Philox(seed))
    deg = user_degrees(n_users, plays)
    T = int(deg.sum())
    cdf = np.cumsum((np.arange(1, n_tracks + 1, dtype=np.float64)) ** -1.0)
    cdf /= cdf[-1]
    items = np.minimum(np.searchsorted(cdf, rng.random(T), side="right"), n_tracks - 1)
    users = np.repeat(np.arange(n_users), deg)
    order = rng.permutation(T)
    artist_of = rng.integers(0, n_artists, n_tracks)
    with open(path, "w") as f:
        for t, e in enumerate(order):
            it = int(items[e])
            f.write("%d,u%d,t%d,a%d\n" % (1500000000 + t, int(users[e]), it, int(artist_of[it])))
    return T


# the five BASELINE.json configs (shape only; hyper-parameters live with the callers)
CONFIGS = {
    "C1": dict(users=4000, tracks=50000, plays=100000, d=10),
    "C2": dict(users=1_000_000, tracks=200_000, plays=50_000_000, d=64),
    "C3": dict(users=10_000_000, tracks=2_000_000, plays=1_000_000_000, d=128),
    "C4": dict(users=1_000_000, tracks=2_000_000, d=64, mask_mean=50),
    "C5": dict(users=1_000_000, tracks=200_000, plays=50_000_000, d=64),
}


def power_law_log_torch(m, n, plays, seed, test_ratio=0.0, device=None, alpha=0.8, beta=1.0):
    """Same distributions as power_law_log, generated with torch ops (on the GPU when there is
    one: ~1 s for config C2 instead of ~1 min of numpy).  Bench/data plumbing only -- the stream
    differs from the numpy generator's, both are seeded and deterministic per device type."""
    import torch
    if device is None:
        device = "cuda" if torch.cuda.is_available() else "cpu"
    g = torch.Generator(device=device)
    g.manual_seed(int(seed))
    deg_np = user_degrees(m, plays, alpha)
    deg = torch.from_numpy(deg_np).to(device)
    T = int(deg_np.sum())
    cdf = torch.cumsum(torch.arange(1, n + 1, device=device, dtype=torch.float64) ** (-beta), 0)
    cdf /= cdf[-1].clone()
    items = torch.empty(T, dtype=torch.int64, device=device)
    step = 1 << 24                                  # bounded temporaries
    for a in range(0, T, step):
        b = min(T, a + step)
        r = torch.rand(b - a, generator=g, device=device, dtype=torch.float64)
        items[a:b] = torch.searchsorted(cdf, r, right=True).clamp_(max=n - 1)
    users = torch.repeat_interleave(torch.arange(m, device=device, dtype=torch.int64), deg)
    if test_ratio > 0:
        held = torch.rand(T, generator=g, device=device) < test_ratio
    else:
        held = torch.zeros(T, dtype=torch.bool, device=device)
    tr_u, tr_i = users[~held], items[~held]
    ev_indptr = torch.zeros(m + 1, dtype=torch.int64, device=device)
    ev_indptr[1:] = torch.cumsum(torch.bincount(tr_u, minlength=m), 0)
    tr_key = torch.unique(tr_u * n + tr_i)           # sorted unique (user, item)
    uq_u = tr_key // n
    uq_indptr = torch.zeros(m + 1, dtype=torch.int64, device=device)
    uq_indptr[1:] = torch.cumsum(torch.bincount(uq_u, minlength=m), 0)
    uq_items = (tr_key - uq_u * n).to(torch.int32)
    te_key = torch.unique(users[held] * n + items[held])
    if te_key.numel():
        pos = torch.searchsorted(tr_key, te_key).clamp_(max=max(tr_key.numel() - 1, 0))
        te_key = te_key[tr_key[pos] != te_key] if tr_key.numel() else te_key
    te_u = te_key // n
    test_indptr = torch.zeros(m + 1, dtype=torch.int64, device=device)
    test_indptr[1:] = torch.cumsum(torch.bincount(te_u, minlength=m), 0)
    test_items = (te_key - te_u * n).to(torch.int32)
    host = lambda t: t.cpu().numpy()                 # noqa: E731
    return PlayLog(m, n, host(ev_indptr), host(tr_i.to(torch.int32)), host(uq_indptr), host(uq_items),
                   host(test_indptr), host(test_items))
