"""K0: the array form of the play log built on the device (yue_ingest_events) against the oracle's
restatement of data/record.py:138-202 + BPR.py:32-45 and the reference-pinned golden record."""
import json
import os

import numpy as np
import pytest

from oracle import record_ref
from yue_b200.engine import MODE_SERIAL, YueError

pytestmark = pytest.mark.gpu


def numpy_arrays(m, n, u, it, is_test):
    """Plain restatement on id arrays: stable grouping, sorted unique rows, test minus train."""
    tr = ~is_test.astype(bool)
    tu, ti = u[tr], it[tr]
    order = np.argsort(tu, kind="stable")
    ev_indptr = np.concatenate([[0], np.cumsum(np.bincount(tu, minlength=m))]).astype(np.int64)
    key = np.unique(tu.astype(np.int64) * n + ti)
    uq_indptr = np.concatenate([[0], np.cumsum(np.bincount(key // n, minlength=m))]).astype(np.int64)
    tk = np.unique(u[~tr].astype(np.int64) * n + it[~tr])
    tk = tk[~np.isin(tk, key)]
    te_indptr = np.concatenate([[0], np.cumsum(np.bincount(tk // n, minlength=m))]).astype(np.int64)
    return ev_indptr, ti[order].astype(np.int32), uq_indptr, (key % n).astype(np.int32), te_indptr, (tk % n).astype(np.int32)


@pytest.mark.parametrize("m,n,E,seed", [(50, 30, 2000, 1), (5000, 800, 300000, 2), (7, 100000, 5, 3)])
def test_ingest_matches_restatement(engine, m, n, E, seed):
    rng = np.random.default_rng(seed)
    u = rng.integers(0, m, E).astype(np.int32)
    u[u == 3] = 4                                     # a user without any event
    it = np.minimum(rng.zipf(1.3, E) - 1, n - 1).astype(np.int32)      # heavy repeats: hot tracks get re-labelled on the device
    is_test = (rng.random(E) < 0.2).astype(np.uint8)
    engine.ingest_events(m, n, u, it, is_test)
    ref = numpy_arrays(m, n, u, it, is_test)
    got = engine.get_interactions() + engine.get_test_set()
    for a, b in zip(got, ref):
        assert a.dtype == b.dtype and np.array_equal(a, b)
    assert engine.interaction_sizes() == (m, n, len(ref[1]), len(ref[3]), len(ref[5]))


def test_ingest_equals_record_preprocess(engine, golden_dir):
    """The same log through the oracle's Record.preprocess restatement (dicts of names, pinned by the
    reference's own Record in tests/golden) and through the device path on the numbered events."""
    g = json.load(open(os.path.join(golden_dir, "record_small.json")))
    training = [e for e, h in zip(g["events"], g["held"]) if not h]
    test = [e for e, h in zip(g["events"], g["held"]) if h]
    name2id, user_record, test_set = record_ref.preprocess(training, test)
    assert name2id == g["name2id"] and {u: dict(d) for u, d in test_set.items()} == g["testSet"]     # = the reference's Record
    ev_indptr, ev_items, uq_indptr, uq_items = record_ref.interaction_arrays(name2id, user_record)
    m, n = len(name2id["user"]), len(name2id["track"])
    u = np.array([name2id["user"][e["user"]] for e in g["events"]], dtype=np.int32)          # file order, flags in place
    it = np.array([name2id["track"][e["track"]] for e in g["events"]], dtype=np.int32)
    flag = np.array(g["held"], dtype=np.uint8)
    engine.ingest_events(m, n, u, it, flag)
    got = engine.get_interactions()
    for a, b in zip(got, (ev_indptr, ev_items, uq_indptr, uq_items)):
        assert np.array_equal(a, b)
    te_indptr, te_items = engine.get_test_set()
    want = {name2id["user"][usr]: sorted(name2id["track"][t] for t in d) for usr, d in test_set.items()}
    for usr in range(m):
        assert te_items[te_indptr[usr]:te_indptr[usr + 1]].tolist() == want.get(usr, [])
    # the handle is ready to train: one serial epoch runs
    from yue_b200 import synth
    P, Q = synth.init_factors(m, n, 10, seed=1)
    engine.set_factors(P, Q)
    assert np.isfinite(engine.bpr_epoch(0.02, 0.01, 0.01, 1, 0, MODE_SERIAL))


def test_ingest_edge_cases(engine):
    engine.ingest_events(3, 5, np.zeros(0, np.int32), np.zeros(0, np.int32))            # empty log
    assert engine.interaction_sizes() == (3, 5, 0, 0, 0)
    engine.ingest_events(2, 4, np.array([0, 0, 1], np.int32), np.array([1, 1, 3], np.int32), np.array([1, 1, 1], np.uint8))
    assert engine.interaction_sizes() == (2, 4, 0, 0, 2)                                 # everything held out
    with pytest.raises(YueError):
        engine.ingest_events(2, 4, np.array([0, 2], np.int32), np.array([1, 1], np.int32))   # user id out of range
    with pytest.raises(YueError):                                                          # a user who played the whole catalog
        engine.ingest_events(1, 2, np.array([0, 0], np.int32), np.array([0, 1], np.int32))
