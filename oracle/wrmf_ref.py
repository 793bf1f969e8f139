"""Oracle (test infrastructure, never on the product path): WRMF, the implicit-feedback ALS of
``recommender/cf/WRMF.py`` (Hu, Koren, Volinsky), restated on arrays.

What the reference does (WRMF.py:17-88), per iteration:

* user sweep (34-57): ``YtY = Y.T.dot(Y)`` -- a **float32** product, Y is float32 -- then for EVERY user id in
  ``name2id['user']`` (test-only users included; their play set is empty and their row becomes 0)
      A = YtY + Y^T diag(10 r_ui) Y + regU I          (float64: the sparse int64 weights promote)
      b = sum over played tracks of (1 + 10 r_ui) Y[i]    (float64)
      X[u] = inv(A) . b                                  (float64, stored to the float32 table)
  with r_ui = the number of plays of track i by user u in the training log (28-33).  The loss is
  ``sum (1 - X[u].Y[i])^2`` over the played pairs with the X[u] from BEFORE its update (49-50), nothing else (82).
* track sweep (60-80): the same with the roles swapped, ``XtX`` from the X just computed, weights from
  ``listened[track][user]`` (the same counts), and -- a quirk kept here -- ``regU`` again, not ``regI`` (79).
* X = P*10, Y = Q*10 at init (19-20); ``predict = Y.dot(X[u])`` (86-88); no convergence test (84-85).

Rows are independent inside a sweep (the solve of one row reads only the OTHER table), which is what the GPU
kernel exploits; the loop order of the reference therefore does not matter.

``gram='f32'`` reproduces the reference's float32 Gram matrix (numpy's sgemm, whatever its summation order is);
``gram='f64'`` is the mathematically cleaner variant the CUDA kernel implements (float64 accumulation of the
float32 rows).  tests/test_oracle_golden.py pins ``gram='f32'`` against the output of the reference class itself
(tests/golden/wrmf_small.npz, written by oracle/make_golden_wrmf.py) and measures how far 'f64' is from it.
"""
import numpy as np

ALPHA = 10.0          # WRMF.py:46-47 / 71-72: confidence 1 + 10 r_ui


def pair_counts(ev_indptr, ev_items, uq_indptr, uq_items):
    """plays of every sorted-unique (user, track) pair: WRMF.py:28-33 (userListen[user][track] += 1)."""
    cnt = np.zeros(len(uq_items), dtype=np.int32)
    for u in range(len(ev_indptr) - 1):
        row = uq_items[uq_indptr[u]:uq_indptr[u + 1]]
        pos = np.searchsorted(row, ev_items[ev_indptr[u]:ev_indptr[u + 1]])
        np.add.at(cnt, uq_indptr[u] + pos, 1)
    return cnt


def transpose(m, n, uq_indptr, uq_items, counts):
    """track-major form of the same pairs (data/record.py:160-163, listened[track][user]): users sorted inside a track."""
    users = np.repeat(np.arange(m, dtype=np.int32), np.diff(uq_indptr))
    order = np.lexsort((users, uq_items))
    indptr = np.zeros(n + 1, dtype=np.int64)
    np.cumsum(np.bincount(uq_items, minlength=n), out=indptr[1:])
    return indptr, users[order].astype(np.int32), counts[order].astype(np.int32)


def half_sweep(out, other, indptr, idx, cnt, reg, gram="f32", want_loss=False, alpha=ALPHA):
    """out[r] = inv(G + other[idx_r]^T diag(alpha cnt_r) other[idx_r] + reg I) . sum (1 + alpha cnt) other[idx_r]
    for every row r (in place, float32 table).  Returns the loss of WRMF.py:49-50 when asked."""
    k = other.shape[1]
    if gram == "f32":
        G = other.T.dot(other).astype(np.float64)                     # float32 sgemm, like the reference
    else:
        G = other.astype(np.float64).T.dot(other.astype(np.float64))
    loss = 0.0
    eye = reg * np.eye(k)
    for r in range(out.shape[0]):
        a, b = indptr[r], indptr[r + 1]
        rows32 = other[idx[a:b]]
        if want_loss and b > a:
            err = 1.0 - rows32.dot(out[r]).astype(np.float64)         # float32 dots, float64 squares
            loss += float((err * err).sum())
        rows = rows32.astype(np.float64)
        c = alpha * cnt[a:b].astype(np.float64)
        A = G + (rows.T * c).dot(rows) + eye
        rhs = ((1.0 + c)[:, None] * rows).sum(axis=0)
        out[r] = np.dot(np.linalg.inv(A), rhs)
    return loss


def iteration(X, Y, uq_indptr, uq_items, cnt, it_indptr, it_users, it_cnt, reg, gram="f32"):
    """One pass of WRMF.py:34-83 (users, then tracks).  X, Y float32, updated in place.  Returns the loss."""
    loss = half_sweep(X, Y, uq_indptr, uq_items, cnt, reg, gram, want_loss=True)
    half_sweep(Y, X, it_indptr, it_users, it_cnt, reg, gram)
    return loss
