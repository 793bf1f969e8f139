"""Oracle (test infrastructure, never on the product path): the training loop of CUNE's two-level BPR,
``recommender/advanced/CUNE.py:118-178`` -- groundwork for SURVEY.md 8f row 4 (the CUDA kernel for it is not built yet,
DESIGN.md section 9).

Per training event (u, i), three times (``for n in range(3)``, 129):
* if the user has implicit positives (tracks of its top-K similar users that it has not played, 111-113): draw one of
  them, k, and one unplayed track j; then, EVERY right-hand side evaluated with the rows as they are at that statement,
      P[u] += lr (1 - s(x_uik)) (Q[i] - Q[k]);   Q[i] += lr (1 - s(x_uik)) P[u];   Q[k] -= lr (1 - s(x_uik)) P[u]     134-139
      P[u] += lr/s (1 - s(x_ukj / s)) (Q[k] - Q[j]);   Q[k] += ... P[u];   Q[j] -= ... P[u]                            148-154
      P[u], Q[i], Q[j], Q[k] shrunk by (1 - lr reg)                                                                   156-159
      loss += -log s(x_uik) - log s(x_ukj / s)                                                                        161-162
* otherwise a plain BPR step on (u, i, j) without the shrink (164-171).
After EVERY USER (not once per epoch -- kept as shipped, 174) the loss gets regU |P|^2 + regI |Q|^2.

Rows are float32, the scalars Python floats (float64): numpy rounds every row statement to float32, like BPR.py.
The reference draws k and j with CPython's unseeded ``random.choice``; here both come from Philox streams
(oracle/philox.py): negatives as in BPR with slot n, the implicit positive from slot ``IMPLICIT_SLOT + n``, attempt 0,
mapped to an index of the user's list.  oracle/make_golden_cune.py feeds exactly these draws to the reference's own
loop text and pins this restatement bit for bit.
"""
from math import exp, log

import numpy as np

from . import philox

IMPLICIT_SLOT = 64


def sigmoid(x):
    return 1 / (1 + exp(-x))                  # tool/qmath.py:115-116


def sample_implicit(seed, epoch, slot, ev_user, ip_indptr, event_base=0):
    """position inside the user's implicit-positive list for every event (-1: the list is empty)"""
    ev_user = np.asarray(ev_user, dtype=np.int64)
    lens = np.diff(np.asarray(ip_indptr, dtype=np.int64))[ev_user]
    ev = np.arange(len(ev_user), dtype=np.int64) + event_base
    out = np.full(len(ev_user), -1, dtype=np.int64)
    for L in np.unique(lens[lens > 0]):
        sel = lens == L
        out[sel] = philox.draw_item(seed, epoch, ev[sel], IMPLICIT_SLOT + slot, np.zeros(int(sel.sum()), np.int64), int(L))
    return out


def epoch(P, Q, ev_indptr, ev_items, ip_indptr, ip_items, kpos, neg, lr, regU, regI, s):
    """One pass of CUNE.py:122-174.  kpos[n][e] / neg[n][e]: the draws of repeat n for event e.  Returns the loss."""
    loss = 0.0
    inv_s = 1 / s
    for u in range(len(ev_indptr) - 1):
        if ev_indptr[u + 1] == ev_indptr[u]:
            continue                                           # PositiveSet has no entry for a user without events
        has_ip = ip_indptr[u + 1] > ip_indptr[u]
        for e in range(ev_indptr[u], ev_indptr[u + 1]):
            i = ev_items[e]
            for n in range(3):
                j = neg[n][e]
                if has_ip:
                    k = ip_items[ip_indptr[u] + kpos[n][e]]
                    P[u] += lr * (1 - sigmoid(P[u].dot(Q[i]) - P[u].dot(Q[k]))) * (Q[i] - Q[k])
                    Q[i] += lr * (1 - sigmoid(P[u].dot(Q[i]) - P[u].dot(Q[k]))) * P[u]
                    Q[k] -= lr * (1 - sigmoid(P[u].dot(Q[i]) - P[u].dot(Q[k]))) * P[u]
                    P[u] += inv_s * lr * (1 - sigmoid(inv_s * (P[u].dot(Q[k]) - P[u].dot(Q[j])))) * (Q[k] - Q[j])
                    Q[k] += inv_s * lr * (1 - sigmoid(inv_s * (P[u].dot(Q[k]) - P[u].dot(Q[j])))) * P[u]
                    Q[j] -= inv_s * lr * (1 - sigmoid(inv_s * (P[u].dot(Q[k]) - P[u].dot(Q[j])))) * P[u]
                    P[u] -= lr * regU * P[u]
                    Q[i] -= lr * regI * Q[i]
                    Q[j] -= lr * regI * Q[j]
                    Q[k] -= lr * regI * Q[k]
                    loss += -log(sigmoid(P[u].dot(Q[i]) - P[u].dot(Q[k]))) - \
                        log(sigmoid(inv_s * (P[u].dot(Q[k]) - P[u].dot(Q[j]))))
                else:
                    P[u] += lr * (1 - sigmoid(P[u].dot(Q[i]) - P[u].dot(Q[j]))) * (Q[i] - Q[j])
                    Q[i] += lr * (1 - sigmoid(P[u].dot(Q[i]) - P[u].dot(Q[j]))) * P[u]
                    Q[j] -= lr * (1 - sigmoid(P[u].dot(Q[i]) - P[u].dot(Q[j]))) * P[u]
                    loss += -log(sigmoid(P[u].dot(Q[i]) - P[u].dot(Q[j])))
        loss += regU * (P * P).sum() + regI * (Q * Q).sum()
    return loss
