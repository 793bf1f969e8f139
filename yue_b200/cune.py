"""CUNE-BPR -- collaborative user network embedding + two-level BPR (Zhang et al. 2017) behind the reference's class API
(recommender/advanced/CUNE.py:11-183).  SURVEY.md 8f row 4.

What runs on the GPU is the training loop (CUNE.py:118-178 -> yue_cune_epoch, csrc/cune_sgd.cuh) and, as for every model
here, scoring / ranking / metrics.  The stage before it -- the collaborative user network, random walks and word2vec
user embedding that yield each user's top-K similar users (CUNE.py:34-97) -- is host-side graph / gensim work outside
the hot path: it is reached through `_top_k_similar()`, which follows the reference (gensim's Word2Vec, and like the
reference it fails at import when gensim is missing) unless `self.topKSim` was supplied by the caller
({user: [(similar user, similarity), ...]}).

Config keys as in config/CUNE.conf: `CUNE=-T .. -L .. -l .. -w .. -k .. -s .. -ep ..` plus the IterativeRecommender keys.
Kept from the reference: three repeats per event (129); users without implicit positives take a plain BPR step (164-171);
duplicates in a user's implicit-positive list are kept, so a track liked by two similar users is drawn twice as often
(111-113); the per-user regulariser term in the loss (174); isConverged / the learning-rate schedule on the host (176).
Fixed here because the reference leaves it to set iteration order: inside one similar user's contribution the tracks are
listed by ascending id.
"""
import random

import numpy as np

from .bpr import GpuBPRMixin
from .engine import MODE_HOGWILD, MODE_SERIAL
from .host.config import LineConfig
from .host.recommender import IterativeRecommender


def implicit_positive_lists(m, uq_indptr, uq_items, top_k_ids):
    """CSR form of IPositiveSet (CUNE.py:111-113).  top_k_ids: {user id: [similar user ids in rank order]}.
    For every similar user, its played tracks minus the user's own, ascending; contributions concatenated."""
    uq_indptr = np.asarray(uq_indptr, dtype=np.int64)
    uq_items = np.asarray(uq_items, dtype=np.int32)
    indptr = np.zeros(m + 1, dtype=np.int64)
    parts = []
    for u in range(m):
        own = uq_items[uq_indptr[u]:uq_indptr[u + 1]]
        cnt = 0
        for f in top_k_ids.get(u, ()):
            d = np.setdiff1d(uq_items[uq_indptr[f]:uq_indptr[f + 1]], own, assume_unique=True)
            parts.append(d)
            cnt += len(d)
        indptr[u + 1] = indptr[u] + cnt
    items = np.concatenate(parts).astype(np.int32) if parts else np.zeros(0, np.int32)
    return indptr, items


class GpuCUNEMixin(GpuBPRMixin):
    topKSim = None

    def readConfiguration(self):
        super(GpuCUNEMixin, self).readConfiguration()
        options = LineConfig(self.config['CUNE'])                 # CUNE.py:15-24
        self.walkCount = int(options['-T'])
        self.walkLength = int(options['-L'])
        self.walkDim = int(options['-l'])
        self.winSize = int(options['-w'])
        self.topK = int(options['-k'])
        self.s = float(options['-s'])
        self.epoch = int(options['-ep'])

    def _top_k_similar(self):
        """{user: [(similar user, cosine), ...]} -- CUNE.py:34-97 (network, walks, word2vec, cosine top-K)."""
        if self.topKSim is not None:
            return self.topKSim
        import gensim.models.word2vec as w2v                      # CUNE.py:9: the reference needs it too
        from collections import defaultdict
        listen = {u: set(e[self.recType] for e in evs) for u, evs in self.data.userRecord.items()}
        net = defaultdict(list)
        for u1, s1 in listen.items():                             # CUNE.py:45-52
            for u2, s2 in listen.items():
                if u1 != u2:
                    w = len(s1 & s2)
                    if w > 0:
                        net[u1] += [u2] * w
        walks, visited = [], defaultdict(dict)
        for user in net:                                          # CUNE.py:54-72
            for _ in range(self.walkCount):
                path, last = [user], user
                for _ in range(1, self.walkLength):
                    nxt, count = random.choice(net[last]), 0
                    while nxt in visited[last]:
                        nxt = random.choice(net[last])
                        count += 1
                        if count == 10:
                            break
                    path.append(nxt)
                    visited[user][nxt] = 1
                    last = nxt
                walks.append(path)
        random.shuffle(walks)
        try:
            model = w2v.Word2Vec(walks, vector_size=self.walkDim, window=self.winSize, min_count=0, epochs=self.epoch)
        except TypeError:                                         # gensim < 4 spells them size / iter (CUNE.py:77)
            model = w2v.Word2Vec(walks, size=self.walkDim, window=self.winSize, min_count=0, iter=self.epoch)
        users = list(net)
        W = np.stack([model.wv[u] for u in users]).astype(np.float64)
        W /= np.maximum(np.linalg.norm(W, axis=1, keepdims=True), 1e-30)
        sim = W @ W.T
        np.fill_diagonal(sim, -np.inf)
        top = {}
        for a, u in enumerate(users):                             # CUNE.py:87-94
            order = np.argsort(-sim[a], kind='stable')[:self.topK]
            top[u] = [(users[b], float(sim[a, b])) for b in order if np.isfinite(sim[a, b])]
        self.topKSim = top
        return top

    def _implicit_positives(self, eng):
        top = self._top_k_similar()
        uid = self.data.name2id['user']
        ids = {uid[u]: [uid[f[0]] for f in friends] for u, friends in top.items()}
        _, _, uq_indptr, uq_items = eng.get_interactions()
        return implicit_positive_lists(self.m, uq_indptr, uq_items, ids)

    def buildModel(self):
        print('Kind Note: This method will probably take much time.')
        eng = self._push_factors()
        print('Preparing item sets...')
        eng.cune_set_implicit(*self._implicit_positives(eng))
        print('Training...')
        mode = MODE_SERIAL if self._opt('yue.sgd', 'hogwild') == 'serial' else MODE_HOGWILD
        seed = int(self._opt('yue.seed', random.getrandbits(63)))
        iteration = 0
        while iteration < self.maxIter:                           # CUNE.py:119-178
            self.loss = eng.cune_epoch(self.lRate, self.regU, self.regI, self.s, seed, iteration, mode)
            iteration += 1
            if self.isConverged(iteration):
                break
        self._pull_factors()


class CUNE(GpuCUNEMixin, IterativeRecommender):
    def __init__(self, conf, trainingSet=None, testSet=None, fold='[1]'):
        super(CUNE, self).__init__(conf, trainingSet, testSet, fold)
