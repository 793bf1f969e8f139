"""Oracle (test infrastructure): the BPR per-triplet SGD epoch, restated on arrays.

Follows the commented numpy loop of ``recommender/cf/BPR.py:31-62`` (identical text in
``recommender/advanced/Song2vec.py:127-144``) operation by operation, with the dict
lookups replaced by integer arrays:

* events are visited user by user in first-appearance order and, inside a user, in
  file order with duplicates kept (``BPR.py:42-45``) -- callers pass them already in
  that order as ``ev_user`` / ``ev_item``;
* the negative ``j`` comes from ``oracle.philox`` (the reference's ``choice`` stream is
  not reproducible, see philox.py);
* ``BPR.py:50``  s = sigmoid(P[u].Q[i] - P[u].Q[j])           (tool/qmath.py:115-116)
* ``BPR.py:51``  P[u] += lr (1-s) (Q[i]-Q[j])                 -- old Q rows
* ``BPR.py:52-53``  Q[i] += lr (1-s) P[u] ; Q[j] -= ...       -- the UPDATED P[u]
* ``BPR.py:55-57``  P[u] -= lr regU P[u] ; Q[i] -= lr regI Q[i] ; Q[j] -= lr regI Q[j]
* ``BPR.py:58``  loss += -log(s)   (s from the pre-update rows)
* ``BPR.py:59``  loss += regU * sum(P*P) + regI * sum(Q*Q)

``mode='numpy32'`` keeps the reference's arithmetic: float32 rows, Python-float (float64)
scalars, exactly what numpy 2 does with ``float * float32-array``.  ``mode='float64'``
runs everything in double; it is the independent re-derivation used to bound rounding.

Pinned by ``tests/golden/sgd_*.npz``: the outputs of the reference's own loop text
executed here by ``oracle/make_golden.py``.
"""
import math

import numpy as np


def sigmoid(val):
    """tool/qmath.py:115-116."""
    return 1.0 / (1.0 + math.exp(-val))


def init_factors(m, n, k, rng):
    """base/IterativeRecommender.py:36-39: U[0,1) -> float32 -> /10 (stays float32)."""
    P = rng.random((m, k)).astype(np.float32) / 10
    Q = rng.random((n, k)).astype(np.float32) / 10
    return P, Q


def sgd_epoch(P, Q, ev_user, ev_item, ev_neg, lr, regU, regI, mode="numpy32"):
    """One pass over the given triplet stream, in place.  Returns sum of -log(s)."""
    if mode == "float64":
        assert P.dtype == np.float64 and Q.dtype == np.float64
    else:
        assert P.dtype == np.float32 and Q.dtype == np.float32
    loss = 0.0
    for e in range(len(ev_user)):
        u = int(ev_user[e]); i = int(ev_item[e]); j = int(ev_neg[e])
        # the difference of the two dots is taken in the array dtype (float32 in the
        # reference), only then does math.exp promote to a Python float
        s = sigmoid(P[u].dot(Q[i]) - P[u].dot(Q[j]))
        P[u] += lr * (1 - s) * (Q[i] - Q[j])
        Q[i] += lr * (1 - s) * P[u]
        Q[j] -= lr * (1 - s) * P[u]
        P[u] -= lr * regU * P[u]
        Q[i] -= lr * regI * Q[i]
        Q[j] -= lr * regI * Q[j]
        loss += -math.log(s)
    return loss


def reg_loss(P, Q, regU, regI):
    """BPR.py:59 -- float32 products and float32 (pairwise) sums, as numpy does it."""
    return regU * (P * P).sum() + regI * (Q * Q).sum()


class LearningRate:
    """base/IterativeRecommender.py:47-55 and 62-75, minus the print/exit side effects."""

    def __init__(self, init, max_rate):
        self.lRate = float(init)
        self.maxLRate = float(max_rate)
        self.loss = 0.0
        self.lastLoss = 0.0

    def update(self, it):
        if it > 1:
            if abs(self.lastLoss) > abs(self.loss):
                self.lRate *= 1.01
            else:
                self.lRate *= 0.5
        if self.maxLRate > 0 and self.lRate > self.maxLRate:
            self.lRate = self.maxLRate

    def is_converged(self, it, loss):
        self.loss = loss
        if math.isnan(loss):
            raise FloatingPointError("Loss = NaN or Infinity")
        delta = self.lastLoss - self.loss
        converged = abs(delta) < 1e-3
        if not converged:
            self.update(it)
        self.lastLoss = self.loss
        return converged


def train(P, Q, ev_user, ev_item, n_items, uq_indptr, uq_items, max_iter, lr, max_lr,
          regU, regI, seed, mode="numpy32", event_base=0):
    """BPR.py:40-62: epochs until max_iter or convergence.  Returns per-epoch (loss, lr)."""
    from .philox import sample_negatives
    sched = LearningRate(lr, max_lr)
    hist = []
    it = 0
    while it < max_iter:
        neg = sample_negatives(seed, it, ev_user, n_items, uq_indptr, uq_items,
                               event_base=event_base)
        loss = sgd_epoch(P, Q, ev_user, ev_item, neg, sched.lRate, regU, regI, mode)
        # python float + np.float32 -> np.float32 under numpy 2: the reference's epoch loss
        # ends up rounded to float32 (golden sgd_small.npz shows exactly that)
        loss = float(loss + reg_loss(P, Q, regU, regI)) if mode == "numpy32" \
            else loss + float(reg_loss(P, Q, regU, regI))
        it += 1
        lr_used = sched.lRate
        conv = sched.is_converged(it, loss)
        hist.append((loss, lr_used))
        if conv:
            break
    return hist


# ----------------------------------------------------------------------------------------
# Hogwild reference semantics used by the GPU throughput mode (not in the reference; this
# is the build's definition, stated here so the test can check it deterministically).
#
# The throughput kernel processes a *segment* (<= seg_len consecutive events of one user)
# per warp.  Inside a segment the update is the serial rule above applied to a private copy
# of P[u] and to Q rows read when the triplet is reached; row changes are published as
# atomic deltas.  With a single segment in flight at a time (the order below) this equals
# the serial rule exactly, which is what test_bpr_gpu checks with ``max_concurrency=1``.
# ----------------------------------------------------------------------------------------
def segments_of(ev_indptr, seg_len=32):
    """Split every user's event range into chunks of <= seg_len (host logic of the ABI)."""
    begins, users, lens = [], [], []
    for u in range(len(ev_indptr) - 1):
        b, e = int(ev_indptr[u]), int(ev_indptr[u + 1])
        while b < e:
            l = min(seg_len, e - b)
            begins.append(b); users.append(u); lens.append(l)
            b += l
    return (np.asarray(begins, np.int64), np.asarray(users, np.int32),
            np.asarray(lens, np.int32))
