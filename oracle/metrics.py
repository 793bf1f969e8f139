"""Oracle (test infrastructure): ranking metrics, restating ``evaluation/measure.py:6-101``
on integer ids, plus the NDCG@N the reference lacks (definition in 8c).

``origin``   list over test users of the collection of held-out track ids
             (the keys of ``testSet[user]``, data/record.py:182-202)
``rec``      list over the same users of recommended id lists (duplicates allowed, as the
             reference's lossy selection can emit them)

The string formatting mirrors ``Measure.rankingMeasure`` (16-41) so a result can be compared
to the reference's output character for character; pinned by tests/golden/measure_*.json.
"""
import math


def hits(origin, rec):
    """measure.py:7-13 -- size of set(test) & set(predicted) per user."""
    return [len(set(o).intersection(set(r))) for o, r in zip(origin, rec)]


def precision(h, N):
    """measure.py:51-53."""
    return float(sum(h)) / (len(h) * N)


def recall(h, origin):
    """measure.py:91-94."""
    lst = [float(x) / len(o) for x, o in zip(h, origin)]
    return sum(lst) / float(len(lst))


def f1(prec, rec):
    """measure.py:97-101."""
    return 2 * prec * rec / (prec + rec) if (prec + rec) != 0 else 0


def mean_ap(origin, rec, N):
    """measure.py:56-66 (divides by min(n_test_u, N); counts duplicate hits again)."""
    total = 0
    for o, r in zip(origin, rec):
        o = set(o)
        h = 0
        p = 0
        for n, item in enumerate(r):
            if item in o:
                h += 1
                p += h / (n + 1.0)
        total += p / (min(len(o), N) + 0.0)
    return total / len(rec)


def coverage(rec, item_count):
    """measure.py:43-48."""
    seen = set()
    for r in rec:
        seen.update(r)
    return len(seen) / float(item_count)


def ranking_measure(origin, rec, tops, item_count):
    """measure.py:16-41 -- same list of strings."""
    out = []
    for n in tops:
        pred = [list(r[:n]) for r in rec]
        if len(origin) != len(pred):
            raise ValueError("The Lengths of test set and predicted set are not match!")
        h = hits(origin, pred)
        prec = precision(h, n)
        rc = recall(h, origin)
        ind = ['Precision:' + str(prec) + '\n', 'Recall:' + str(rc) + '\n',
               'F1:' + str(f1(prec, rc)) + '\n', 'MAP:' + str(mean_ap(origin, pred, n)) + '\n',
               'Coverage:' + str(coverage(pred, item_count)) + '\n']
        out.append('Top ' + str(n) + '\n')
        out += ind
    return out


def ndcg(origin, rec, N):
    """Binary-relevance NDCG@N (not in the reference):
    DCG = sum_{r=1..N} rel_r / log2(r+1), IDCG = sum_{r=1..min(n_test_u,N)} 1/log2(r+1),
    averaged over test users.  A duplicate id counts once (first occurrence)."""
    total = 0.0
    for o, r in zip(origin, rec):
        o = set(o)
        seen = set()
        dcg = 0.0
        for rank, item in enumerate(r[:N]):
            if item in o and item not in seen:
                dcg += 1.0 / math.log2(rank + 2)
            seen.add(item)
        idcg = sum(1.0 / math.log2(k + 2) for k in range(min(len(o), N)))
        total += dcg / idcg if idcg > 0 else 0.0
    return total / len(rec)
