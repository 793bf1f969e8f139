"""Oracle (test infrastructure): id assignment, test-set semantics and the array form of
the interaction log.  Restates ``data/record.py:138-202`` (preprocess) and the iteration
order of ``recommender/cf/BPR.py:32-45``.

Events are dicts as produced by ``tool/file.py:23-52`` -- key order = the order of
``-columns`` in ``record.setup``.  Ids are handed out by first appearance while walking the
training events and, inside each event, its keys in dict order (``time`` skipped); test
events then extend the same maps (record.py:138-146, 182-188).  ``testSet[user][track]`` counts
plays, pairs already present in the user's training data are deleted and users left empty are
dropped (record.py:189-202).  Pinned by tests/golden/record_*.json (the reference's own Record).
"""
from collections import OrderedDict

import numpy as np


def preprocess(training, test, rec_type="track"):
    name2id = {}
    user_record = OrderedDict()
    listened = {}
    for ev in training:
        for key in ev:
            if key != 'time':
                m = name2id.setdefault(key, {})
                if ev[key] not in m:
                    m[ev[key]] = len(m)
        user_record.setdefault(ev['user'], []).append(ev)
        if rec_type in ev:
            listened.setdefault(ev[rec_type], {}).setdefault(ev['user'], 0)
            listened[ev[rec_type]][ev['user']] += 1
    test_set = OrderedDict()
    for ev in test:
        for key in ev:
            if key != 'time':
                m = name2id.setdefault(key, {})
                if ev[key] not in m:
                    m[ev[key]] = len(m)
        d = test_set.setdefault(ev['user'], OrderedDict())
        d[ev[rec_type]] = d.get(ev[rec_type], 0) + 1
    for item in listened:
        for user in listened[item]:
            if user in test_set:
                test_set[user].pop(item, None)
                if len(test_set[user]) == 0:
                    del test_set[user]
    return name2id, user_record, test_set


def interaction_arrays(name2id, user_record, rec_type="track"):
    """Array form consumed by the C ABI (include/yue_b200.h, yue_set_interactions):

    ev_indptr[m+1], ev_items[T]   every training event, grouped by user in user-id order
                                  (= first-appearance order, BPR.py:42), file order inside a
                                  user, duplicates kept (BPR.py:44-45)
    uq_indptr[m+1], uq_items[nnz] per-user sorted unique played tracks = userListen
                                  (BPR.py:32-35), the rejection and the ranking mask set
    Users that only occur in the test data get empty rows (getSize counts them,
    IterativeRecommender.py:37).
    """
    m = len(name2id['user'])
    rows = [[] for _ in range(m)]
    for user, evs in user_record.items():
        rows[name2id['user'][user]] = [name2id[rec_type][e[rec_type]] for e in evs]
    ev_indptr = np.zeros(m + 1, dtype=np.int64)
    uq_indptr = np.zeros(m + 1, dtype=np.int64)
    ev_items, uq_items = [], []
    for u, r in enumerate(rows):
        ev_items += r
        uq = sorted(set(r))
        uq_items += uq
        ev_indptr[u + 1] = len(ev_items)
        uq_indptr[u + 1] = len(uq_items)
    return (ev_indptr, np.asarray(ev_items, dtype=np.int32),
            uq_indptr, np.asarray(uq_items, dtype=np.int32))


def ev_users(ev_indptr):
    return np.repeat(np.arange(len(ev_indptr) - 1, dtype=np.int32), np.diff(ev_indptr))
