"""Oracle (test infrastructure): the update rule of recommender/cf/BPR.py:50-57 under the READ SCHEDULE of the blocked
throughput kernel (yue_b200/csrc/bpr_sgd_blk.cuh) when ONE warp runs it -- a deterministic restatement that pins what
"Hogwild" means inside a warp, where the statistical quality checks cannot see it.

The arithmetic per triplet is the reference's (P first from the old Q rows, Q with the updated P, then the three
multiplicative shrinks, loss from the pre-update score).  What differs from the serial loop is only WHEN a row of Q is read:
a warp works on blocks of 4 consecutive triplets of one user (segments of <= 32 events, the last block of a segment may be
shorter) and requests the rows of block k while block k-1 is being computed, i.e. before block k-1's changes have been
added -- so block k sees Q with the changes of all blocks up to k-2; a track that repeats inside a block, or in the next
block, is read at that older value, and every change is ADDED (nothing is lost).  At the start of a work item (a new
user) everything has been added.  P[u] lives in registers and is always current.

sgd_apply_blocked(P, Q, u, i, j, ...) updates the float32 tables in place like oracle/bpr_ref.py:sgd_epoch does for the
serial order, and returns the loss.  Row arithmetic in float32 with fused multiply-adds emulated in float64 (the kernel uses
fmaf), scalars in float32 like the kernel's."""
import numpy as np

BLOCK, SEGMENT = 4, 32


def _sigmoid_terms(x, lr):
    """g = lr * sigmoid(-x), loss term = -log(sigmoid(x)): float32 like bpr_grad()."""
    x = np.float32(x)
    ex = np.float32(np.exp(-np.abs(np.float64(x))))
    loss = np.float32(max(-x, np.float32(0))) + np.float32(np.log(np.float64(np.float32(1) + ex)))
    g = np.float32(lr) * (np.float32(ex if x >= 0 else 1.0) / (np.float32(1) + ex))
    return np.float32(g), np.float32(loss)


def sgd_apply_blocked(P, Q, u, i, j, lr, regU, regI):
    u, i, j = np.asarray(u), np.asarray(i), np.asarray(j)
    c_u, c_i = np.float32(lr * regU), np.float32(lr * regI)
    loss = 0.0
    T = len(u)
    t = 0
    while t < T:                                             # one work item = one run of a user
        e = t
        while e < T and u[e] == u[t]:
            e += 1
        pu = P[u[t]].astype(np.float32).copy()
        blocks = []
        for s0 in range(t, e, SEGMENT):
            s1 = min(e, s0 + SEGMENT)
            blocks += [(b0, min(s1, b0 + BLOCK)) for b0 in range(s0, s1, BLOCK)]
        pending = []                                         # changes of the previous block, not yet added when the next block's rows are read
        for (b0, b1) in blocks:
            qi = [Q[i[x]].astype(np.float32).copy() for x in range(b0, b1)]      # rows as requested: before `pending` lands
            qj = [Q[j[x]].astype(np.float32).copy() for x in range(b0, b1)]
            for row, delta in pending:
                Q[row] = (Q[row].astype(np.float32) + delta).astype(np.float32)
            pending = []
            for a, x in enumerate(range(b0, b1)):
                d = qi[a] - qj[a]
                score = np.float32(np.dot(pu.astype(np.float64), d.astype(np.float64)))
                g, l = _sigmoid_terms(score, lr)
                loss += float(l)
                pn = (pu.astype(np.float64) + np.float64(g) * d.astype(np.float64)).astype(np.float32)          # fmaf(g, d, pu)
                gp = (g * pn).astype(np.float32)
                di = ((-np.float64(c_i)) * (qi[a] + gp).astype(np.float32).astype(np.float64) + gp.astype(np.float64)).astype(np.float32)
                dj = ((-np.float64(c_i)) * (qj[a] - gp).astype(np.float32).astype(np.float64) - gp.astype(np.float64)).astype(np.float32)
                pu = ((-np.float64(c_u)) * pn.astype(np.float64) + pn.astype(np.float64)).astype(np.float32)
                pending += [(i[x], di), (j[x], dj)]
        for row, delta in pending:
            Q[row] = (Q[row].astype(np.float32) + delta).astype(np.float32)
        P[u[t]] = pu
        t = e
    return loss
