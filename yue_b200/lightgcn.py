"""LightGCN behind the reference's class API (recommender/advanced/LightGCN.py:10-105 on base/DeepRecommender:7-59).
SURVEY.md 8f row 4.

The reference builds a TF-1 graph: three sparse products of the un-normalised play graph with [U; V], the row-normalised
layers summed, a BPR loss on the propagated rows of a batch of 128 consecutive training events, AdamOptimizer on U and V
(every row moves every step).  Here the whole pass over the batches is ONE cooperative kernel launch
(csrc/lightgcn.cuh, yue_gcn_epoch); scoring / ranking / metrics are the BPR path's, on the propagated tables.

Kept from the reference: truncated-normal(0.005) initialisation (DeepRecommender:30-31); the adjacency weighs a pair
played c times with c * c (one SparseTensor entry of value c per event, LightGCN.py:29-33); three layers (35); batches are
slices of the training events in file order and one negative per event is kept (the fifth of five draws, 56-79);
regU alone regularises (86-87); exactly num.max.iter passes, no learning-rate schedule (94-98); predict ranks with the
propagated tables (101-105).  Not kept: under `-byTime` the reference's trainingData is the unsplit log, so its graph
and batches leak the held-out events (data/record.py:37 against 47-48); here the training split only.  The reference prints every batch's loss; here the first and last of every pass are
printed (`yue.verbose=on` prints all).
"""
import random

import numpy as np

from .bpr import GpuBPRMixin
from .host.recommender import IterativeRecommender


def truncated_normal(shape, stddev, rng=np.random):
    """tf.truncated_normal: values beyond two standard deviations are drawn again."""
    x = rng.standard_normal(shape)
    bad = np.abs(x) > 2.0
    while bad.any():
        x[bad] = rng.standard_normal(int(bad.sum()))
        bad = np.abs(x) > 2.0
    return (stddev * x).astype(np.float32)


class GpuLightGCNMixin(GpuBPRMixin):
    n_layers = 3                                               # LightGCN.py:35
    negativeCount = 5                                          # 19; only the last draw is used (76-78)
    #: the tables evalRanking / predict read are the propagated ones (the reference's tensors of the same names)
    _user_table, _item_table = 'multi_user_embeddings', 'multi_item_embeddings'

    def readConfiguration(self):
        super(GpuLightGCNMixin, self).readConfiguration()
        self.batch_size = int(self.config['batch_size'])       # DeepRecommender:11-17

    def initModel(self):
        super(GpuLightGCNMixin, self).initModel()              # m, n, train_size
        self.U = truncated_normal((self.m, self.k), 0.005)     # DeepRecommender:30-31
        self.V = truncated_normal((self.n, self.k), 0.005)
        self.multi_user_embeddings, self.multi_item_embeddings = self.U, self.V

    def _file_order_events(self):
        """(user ids, track ids) of trainingData in file order (LightGCN.py:59-60)."""
        if hasattr(self.data, 'log'):                          # yue.ingest=arrays: the numbered events are already arrays
            log = self.data.log
            keep = np.flatnonzero(log.is_test == 0)
            if log.file_pos is not None:                       # -byTime grouped the events by user: back to the file's order
                keep = keep[np.argsort(log.file_pos[keep], kind='stable')]
            return log.ev_user[keep], log.ev_item[keep]
        uid, tid = self.data.name2id['user'], self.data.name2id[self.recType]
        events = self.data.trainingData
        n_train = sum(len(evs) for evs in self.data.userRecord.values())
        if len(events) != n_train:
            # -byTime (config/LightGCN.conf): Record keeps the UNSPLIT log in trainingData (data/record.py:37 runs before the
            # split at 47-48), so the reference's graph and batches include the held-out events -- a leak that is not
            # reproduced: the events of the training split (the entries userRecord holds), in the file's order
            kept = set(id(e) for evs in self.data.userRecord.values() for e in evs)
            events = [e for e in events if id(e) in kept]
            if len(events) != n_train:                         # a Record that copies its entries: the users' own order
                events = [e for evs in self.data.userRecord.values() for e in evs]
        ev_user = np.fromiter((uid[e['user']] for e in events), dtype=np.int32, count=len(events))
        ev_item = np.fromiter((tid[e[self.recType]] for e in events), dtype=np.int32, count=len(events))
        return ev_user, ev_item

    def buildModel(self):
        print('training...')
        eng = self._get_engine()
        eng.set_factors(self.U, self.V)
        eng.gcn_set_events(*self._file_order_events())
        seed = int(self._opt('yue.seed', random.getrandbits(63)))
        verbose = self._opt('yue.verbose', 'off') == 'on'
        for iteration in range(self.maxIter):                  # LightGCN.py:94-98
            losses = eng.gcn_epoch(self.batch_size, self.lRate, self.regU, seed, iteration, self.n_layers)
            shown = range(len(losses)) if verbose else sorted({0, len(losses) - 1})
            for n in shown:
                print('training:', iteration + 1, 'batch', n, 'loss:', losses[n])
            self.loss = float(losses[-1]) if len(losses) else 0.0
        self.U, self.V = eng.get_factors()
        eng.gcn_finalize(self.n_layers)
        self._pull_factors()                                   # multi_*_embeddings <- F; ranking and predict read them


class LightGCN(GpuLightGCNMixin, IterativeRecommender):
    def __init__(self, conf, trainingSet=None, testSet=None, fold='[1]'):
        super(LightGCN, self).__init__(conf, trainingSet, testSet, fold)
