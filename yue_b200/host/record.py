"""Record: the data container the recommenders read, with the semantics of the reference's
data/record.py:11-233, plus the array form the CUDA library consumes.

Ids are handed out by first appearance -- training events first, walking each event's keys in
``-columns`` order and skipping ``time`` -- and the test events then extend the same maps
(record.py:138-146, 182-188).  ``userRecord[user]`` keeps every training event in file order
(repeat plays included); ``testSet[user][track]`` counts held-out plays, minus every pair the
user already has in training, minus users left empty (record.py:189-202).  ``-byTime r`` splits
each user's time-sorted events into the first (1-r) share for training and the rest for test
(record.py:108-123).
"""
from collections import defaultdict

import numpy as np

from .config import LineConfig


class Record(object):
    'data access control'

    def __init__(self, config, trainingSet, testSet):
        self.config = config
        self.recordConfig = LineConfig(config['record.setup'])
        self.evalConfig = LineConfig(config['evaluation.setup'])
        self.name2id = defaultdict(dict)
        self.id2name = defaultdict(dict)
        self.listened = {'artist': defaultdict(dict), 'track': defaultdict(dict), 'album': defaultdict(dict)}
        self.userRecord = defaultdict(list)
        self.trackRecord = defaultdict(list)
        self.testSet = defaultdict(dict)
        self.recordCount = 0
        self.globalMean = 0
        self.userMeans = {}
        self.PopTrack = {}
        self.columns = {}
        self.trainingData = trainingSet
        for col in self.recordConfig['-columns'].split(','):
            name, pos = col.split(':')
            self.columns[name] = int(pos)
        if self.evalConfig.contains('-byTime'):
            trainingSet, testSet = self.splitDataByTime(trainingSet)
        self.preprocess(trainingSet, testSet)
        self._arrays = None

    # -- splitting -------------------------------------------------------------------------
    def splitDataByTime(self, dataset):
        ratio = float(self.evalConfig['-byTime'])
        per_user = defaultdict(list)
        for event in dataset:
            per_user[event['user']].append(event)
        trainingSet, testSet = [], []
        for user, events in per_user.items():
            ordered = sorted(events, key=lambda d: d['time'])
            cut = int(len(ordered) * (1 - ratio))
            trainingSet += ordered[:cut]
            testSet += ordered[cut:]
        return trainingSet, testSet

    # -- id maps and containers ------------------------------------------------------------
    def _register(self, entry):
        for key, value in entry.items():
            if key == 'time':
                continue
            ids = self.name2id[key]
            if value not in ids:
                self.id2name[key][len(ids)] = value
                ids[value] = len(ids)

    def preprocess(self, trainingSet, testSet):
        recType = self.evalConfig['-target']
        for entry in trainingSet:
            self.recordCount += 1
            self._register(entry)
            user = entry['user']
            self.userRecord[user].append(entry)
            for kind in ('artist', 'album', 'track'):
                if kind in entry:
                    seen = self.listened[kind][entry[kind]]
                    seen[user] = seen.get(user, 0) + 1
            if 'track' in entry:
                self.trackRecord[entry['track']].append(entry)
        for entry in testSet:
            self._register(entry)
            held = self.testSet[entry['user']]
            held[entry[recType]] = held.get(entry[recType], 0) + 1
        # a held-out pair the user already has in training is not a test item
        for item, users in self.listened[recType].items():
            for user in users:
                if user in self.testSet:
                    self.testSet[user].pop(item, None)
                    if not self.testSet[user]:
                        del self.testSet[user]

    # -- accessors of the reference --------------------------------------------------------
    def printTrainingSize(self):
        for kind in ('user', 'artist', 'album', 'track'):
            if kind in self.name2id:
                print(kind + ' count:', len(self.name2id[kind]))
        print('Training set size:', self.recordCount)

    def getId(self, obj, t):
        if obj in self.name2id[t]:
            return self.name2id[t][obj]
        print('No ' + t + ' ' + obj + ' exists!')
        exit(-1)

    def getSize(self, t):
        return len(self.name2id[t])

    def contains(self, obj, t):
        return obj in self.name2id[t]

    # -- array form for the CUDA library ---------------------------------------------------
    def interaction_arrays(self, recType='track'):
        """(ev_indptr, ev_items, uq_indptr, uq_items) as described in include/yue_b200.h.
        Event order = BPR.py:42-45 (users in id order = first appearance, file order inside)."""
        return interaction_arrays(self.name2id, self.userRecord, recType)


def interaction_arrays(name2id, userRecord, recType='track'):
    m = len(name2id['user'])
    uid, tid = name2id['user'], name2id[recType]
    deg = np.zeros(m, dtype=np.int64)
    for user, events in userRecord.items():
        deg[uid[user]] = len(events)
    ev_indptr = np.zeros(m + 1, dtype=np.int64)
    np.cumsum(deg, out=ev_indptr[1:])
    ev_items = np.empty(int(ev_indptr[-1]), dtype=np.int32)
    for user, events in userRecord.items():
        b = ev_indptr[uid[user]]
        ev_items[b:b + len(events)] = [tid[e[recType]] for e in events]
    ev_user = np.repeat(np.arange(m, dtype=np.int64), deg)
    n = max(len(tid), 1)
    key = np.unique(ev_user * n + ev_items)
    uq_u = key // n
    uq_indptr = np.zeros(m + 1, dtype=np.int64)
    np.cumsum(np.bincount(uq_u, minlength=m), out=uq_indptr[1:])
    return ev_indptr, ev_items, uq_indptr, (key - uq_u * n).astype(np.int32)
