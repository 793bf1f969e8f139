"""Oracle (test infrastructure): Philox4x32-10 and the negative-item sampler.

The reference draws negatives with CPython's ``random.choice(itemList)`` and
redraws while the track is in the user's play set (``recommender/cf/BPR.py:46-49``;
id-based variants ``BPR.py:73-76`` and ``recommender/advanced/APR.py:104-107``).
That Mersenne-Twister stream is unseeded and order-dependent, so the build fixes
its own counter-based stream whose draw for attempt ``t``
of event ``e`` in epoch ``E`` is a pure function of ``(seed, E, e, slot, t)``:

    key     = (seed & 0xffffffff, seed >> 32)
    counter = (e & 0xffffffff, e >> 32, E, (slot << 20) | (t >> 2))
    r       = Philox4x32-10(counter, key)[t & 3]
    j       = (r * n) >> 32                       # uniform over [0, n)
    accept j unless j is in the user's sorted-unique play row, else t += 1

``slot`` numbers the negatives of one positive (0 for BPR's single negative,
0..2 for APR's three).  Philox4x32-10 follows the published Random123 algorithm
(Salmon et al., SC'11); the known-answer vectors in ``PHILOX_KAT`` are the
Random123 ``kat_vectors`` entries for philox4x32-10.  Parity unpinned by the
reference (it has no such stream); pinned by the KATs.
"""
import numpy as np

PHILOX_M0 = np.uint64(0xD2511F53)
PHILOX_M1 = np.uint64(0xCD9E8D57)
PHILOX_W0 = 0x9E3779B9
PHILOX_W1 = 0xBB67AE85
MAX_ATTEMPTS = 1 << 22      # 20 counter bits * 4 words; a full play row is an error upstream

# (counter, key, expected output) from Random123's known-answer file
PHILOX_KAT = [
    ((0x00000000, 0x00000000, 0x00000000, 0x00000000), (0x00000000, 0x00000000),
     (0x6627e8d5, 0xe169c58d, 0xbc57ac4c, 0x9b00dbd8)),
    ((0xffffffff, 0xffffffff, 0xffffffff, 0xffffffff), (0xffffffff, 0xffffffff),
     (0x408f276d, 0x41c83b0e, 0xa20bc7c6, 0x6d5451fd)),
    ((0x243f6a88, 0x85a308d3, 0x13198a2e, 0x03707344), (0xa4093822, 0x299f31d0),
     (0xd16cfe09, 0x94fdcceb, 0x5001e420, 0x24126ea1)),
]

_MASK32 = np.uint64(0xFFFFFFFF)
_SH32 = np.uint64(32)


def philox4x32_10(c0, c1, c2, c3, k0, k1):
    """Vectorised Philox4x32-10.  All arguments broadcast; returns four uint32 arrays."""
    c0, c1, c2, c3 = (np.asarray(x).astype(np.uint64) & _MASK32 for x in (c0, c1, c2, c3))
    c0, c1, c2, c3 = np.broadcast_arrays(c0, c1, c2, c3)
    k0 = int(k0) & 0xFFFFFFFF
    k1 = int(k1) & 0xFFFFFFFF
    for _ in range(10):
        p0 = PHILOX_M0 * c0          # 32x32 -> 64 bit, no overflow in uint64
        p1 = PHILOX_M1 * c2
        n0 = (p1 >> _SH32) ^ c1 ^ np.uint64(k0)
        n1 = p1 & _MASK32
        n2 = (p0 >> _SH32) ^ c3 ^ np.uint64(k1)
        n3 = p0 & _MASK32
        c0, c1, c2, c3 = n0, n1, n2, n3
        k0 = (k0 + PHILOX_W0) & 0xFFFFFFFF
        k1 = (k1 + PHILOX_W1) & 0xFFFFFFFF
    return tuple(x.astype(np.uint32) for x in (c0, c1, c2, c3))


def draw_item(seed, epoch, event, slot, attempt, n_items):
    """Candidate item id for (event, slot, attempt); vectorised over event/attempt."""
    event = np.asarray(event, dtype=np.uint64)
    attempt = np.asarray(attempt, dtype=np.uint64)
    w0, w1, w2, w3 = philox4x32_10(
        event & _MASK32, event >> _SH32, np.uint64(epoch),
        (np.uint64(slot) << np.uint64(20)) | (attempt >> np.uint64(2)),
        seed & 0xFFFFFFFF, (seed >> 32) & 0xFFFFFFFF)
    sel = np.broadcast_to(attempt & np.uint64(3), w0.shape)
    r = np.where(sel == 0, w0, np.where(sel == 1, w1, np.where(sel == 2, w2, w3)))
    return ((r.astype(np.uint64) * np.uint64(n_items)) >> _SH32).astype(np.int64)


def _in_rows(uq_indptr, uq_items, users, items):
    """Membership of items[k] in the sorted-unique row of users[k] (vectorised)."""
    # rows are sorted, so a global key (user * n_key + item) is sorted too
    n_key = np.int64(int(uq_items.max()) + 1 if uq_items.size else 1)
    n_key = max(n_key, np.int64(int(items.max()) + 1 if items.size else 1))
    row_of = np.repeat(np.arange(len(uq_indptr) - 1, dtype=np.int64), np.diff(uq_indptr))
    keys = row_of * n_key + uq_items.astype(np.int64)
    q = users.astype(np.int64) * n_key + items.astype(np.int64)
    pos = np.searchsorted(keys, q)
    pos_c = np.minimum(pos, max(len(keys) - 1, 0))
    return (pos < len(keys)) & (keys[pos_c] == q) if len(keys) else np.zeros(len(q), bool)


def sample_negatives(seed, epoch, ev_user, n_items, uq_indptr, uq_items, slot=0,
                     event_base=0, return_attempts=False):
    """One accepted negative per event: restates the redraw loop of BPR.py:46-49 on the
    Philox stream.  ``ev_user[e]`` is the user of global event ``event_base + e``."""
    ev_user = np.asarray(ev_user, dtype=np.int64)
    T = len(ev_user)
    out = np.full(T, -1, dtype=np.int32)
    tries = np.zeros(T, dtype=np.int64)
    pending = np.arange(T, dtype=np.int64)
    attempt = 0
    uq_indptr = np.asarray(uq_indptr, dtype=np.int64)
    uq_items = np.asarray(uq_items)
    while len(pending):
        if attempt >= MAX_ATTEMPTS:
            raise RuntimeError("sampler exhausted: a user has played every track")
        cand = draw_item(seed, epoch, pending + event_base, slot,
                         np.full(len(pending), attempt, dtype=np.uint64), n_items)
        hit = _in_rows(uq_indptr, uq_items, ev_user[pending], cand)
        ok = ~hit
        out[pending[ok]] = cand[ok].astype(np.int32)
        tries[pending[ok]] = attempt + 1
        pending = pending[hit]
        attempt += 1
    return (out, tries) if return_attempts else out


def attempt_stream(seed, epoch, ev_user, n_items, uq_indptr, uq_items, slot=0, event_base=0):
    """Flat list of EVERY candidate drawn (rejected ones included), event by event, in the
    order the reference's ``choice()`` would be called.  Used by make_golden.py to feed the
    reference's own loop."""
    j, tries = sample_negatives(seed, epoch, ev_user, n_items, uq_indptr, uq_items, slot,
                                event_base, return_attempts=True)
    flat = []
    for e in range(len(j)):
        for t in range(int(tries[e])):
            flat.append(int(draw_item(seed, epoch, np.array([e + event_base]), slot,
                                      np.array([t]), n_items)[0]))
    return flat, j
