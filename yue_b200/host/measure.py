"""Ranking metrics with the interface and output format of the reference's
evaluation/measure.py:2-101, plus the binary-relevance NDCG@N the reference lacks.

``origin`` is ``{user: {item: count}}`` (Record.testSet), ``res`` is ``{user: [item, ...]}``.
"""
import math


class Measure(object):
    @staticmethod
    def hits(origin, res):
        return {user: len(set(origin[user]).intersection(res[user])) for user in origin}

    @staticmethod
    def precision(hits, N):
        return float(sum(hits.values())) / (len(hits) * N)

    @staticmethod
    def recall(hits, origin):
        per_user = [float(hits[user]) / len(origin[user]) for user in hits]
        return sum(per_user) / float(len(per_user))

    @staticmethod
    def F1(prec, recall):
        return 2 * prec * recall / (prec + recall) if (prec + recall) != 0 else 0

    @staticmethod
    def MAP(origin, res, N):
        total = 0
        for user, items in res.items():
            found, prec = 0, 0
            for rank, item in enumerate(items):
                if item in origin[user]:
                    found += 1
                    prec += found / (rank + 1.0)
            total += prec / (min(len(origin[user]), N) + 0.0)
        return total / len(res)

    @staticmethod
    def coverage(res, itemCount):
        distinct = set()
        for items in res.values():
            distinct.update(items)
        return len(distinct) / float(itemCount)

    @staticmethod
    def NDCG(origin, res, N):
        """DCG = sum_r rel_r/log2(r+1) over the first N, IDCG over min(n_test_u, N) ones; a repeated
        id counts once.  Not part of rankingMeasure's output (the reference has no NDCG)."""
        total = 0.0
        for user, items in res.items():
            seen, dcg = set(), 0.0
            for rank, item in enumerate(items[:N]):
                if item in origin[user] and item not in seen:
                    dcg += 1.0 / math.log2(rank + 2)
                seen.add(item)
            ideal = sum(1.0 / math.log2(r + 2) for r in range(min(len(origin[user]), N)))
            total += dcg / ideal if ideal > 0 else 0.0
        return total / len(res)

    @staticmethod
    def rankingMeasure(origin, res, N, itemCount):
        print('rank measure...')
        measure = []
        for n in N:
            predicted = {user: res[user][:n] for user in res}
            if len(origin) != len(predicted):
                print('The Lengths of test set and predicted set are not match!')
                exit(-1)
            hits = Measure.hits(origin, predicted)
            prec = Measure.precision(hits, n)
            recall = Measure.recall(hits, origin)
            measure.append('Top ' + str(n) + '\n')
            measure.append('Precision:' + str(prec) + '\n')
            measure.append('Recall:' + str(recall) + '\n')
            measure.append('F1:' + str(Measure.F1(prec, recall)) + '\n')
            measure.append('MAP:' + str(Measure.MAP(origin, predicted, n)) + '\n')
            measure.append('Coverage:' + str(Measure.coverage(predicted, itemCount)) + '\n')
        return measure
