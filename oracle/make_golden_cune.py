#!/usr/bin/env python
"""Generate tests/golden/cune_small.npz: the training loop of the REFERENCE's CUNE (recommender/advanced/CUNE.py:119-178),
read from the reference file at run time, exec'd and attached to a subclass of the reference's own IterativeRecommender
(the module itself cannot be imported: it needs gensim for the embedding stage that precedes the loop), with
``random.choice`` replaced by the Philox streams of oracle/cune_ref.py.  Test infrastructure; build container only.

The implicit-positive sets, which the reference derives from the embedding stage (CUNE.py:76-113), are constructed here:
for most users the tracks of two other users that the user has not played (duplicates kept, as `+=` of the per-friend
lists does), for every fifth user none -- so both branches of the loop run.

    python oracle/make_golden_cune.py
"""
import io
import json
import os
import sys
import tempfile
import textwrap
from collections import defaultdict
from contextlib import redirect_stdout
from math import log

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
REF = os.environ.get("YUE_REFERENCE", "/root/reference")
sys.path.insert(0, ROOT)
sys.path.insert(1, REF)
OUT = os.path.join(ROOT, "tests", "golden")

from oracle import cune_ref, philox, record_ref  # noqa: E402

SEED = 20260105


class IPList(list):
    """marks the lists `choice` is asked to draw an implicit positive from"""


def reference_loop():
    src = open(os.path.join(REF, "recommender", "advanced", "CUNE.py"), encoding="utf8").read()
    a = src.index("        iteration = 0\n        while iteration < self.maxIter:", src.index("print ('Training...')"))
    b = src.index("    def predict(self, u):")
    return "def train(self):\n" + textwrap.indent(textwrap.dedent(src[a:b]), "    ")


def main():
    from tool.config import Config
    from tool.qmath import sigmoid
    from base.IterativeRecommender import IterativeRecommender

    g = json.load(open(os.path.join(OUT, "record_small.json")))
    keep = set('u%d' % x for x in range(60))                   # a small slice: the loop is three triplets per event
    train = [e for e, h in zip(g["events"], g["held"]) if not h and e['user'] in keep]
    test = [e for e, h in zip(g["events"], g["held"]) if h and e['user'] in keep]
    tmp = tempfile.mkdtemp(prefix="yue_golden_cune_")
    K, ITERS, S = 12, 2, 2.0
    cpath = os.path.join(tmp, "cune.conf")
    with open(cpath, "w") as f:
        f.write("record=%s\nrecord.setup=-columns user:1,track:2,artist:3,time:0 -delim ,\nrecommender=CUNE\n"
                "evaluation.setup=-target track -ap 0.2\nitem.ranking=-topN 5,10\nnum.factors=%d\nnum.max.iter=%d\n"
                "CUNE=-T 20 -L 10 -l 20 -w 5 -k 50 -s %g -ep 10\nlearnRate=-init 0.02 -max 0.1\n"
                "reg.lambda=-u 0.01 -i 0.01 -b 0.01 -s 0.2\noutput.setup=on -dir %s/res/\n"
                % (os.path.join(tmp, "log.txt"), K, ITERS, S, tmp))
    body = reference_loop()
    assert "IPositiveSet" in body and "sigmoid(" in body and "isConverged" in body
    snaps = []

    class RefCUNE(IterativeRecommender):
        def isConverged(self, it):
            snaps.append((self.P.copy(), self.Q.copy(), float(self.loss)))
            return False                                        # run exactly maxIter iterations (the lr schedule is BPR's test)

    with redirect_stdout(io.StringIO()):
        model = RefCUNE(Config(cpath), train, test)
        model.readConfiguration()
        np.random.seed(99)
        model.initModel()
    model.s = S
    name2id = {k: dict(v) for k, v in model.data.name2id.items()}
    u2i, t2i = name2id['user'], name2id['track']
    i2t = {v: k for k, v in t2i.items()}
    ev_indptr, ev_items, uq_indptr, uq_items = record_ref.interaction_arrays(name2id, model.data.userRecord)
    ev_user = record_ref.ev_users(ev_indptr)
    m, n = len(u2i), len(t2i)
    # PositiveSet as the reference builds it (CUNE.py:107-109), implicit positives constructed (see the docstring)
    model.PositiveSet = defaultdict(list)
    for user in model.data.userRecord:
        for event in model.data.userRecord[user]:
            model.PositiveSet[user].append(event[model.recType])
    users_by_id = [None] * m
    for name, uid in u2i.items():
        users_by_id[uid] = name
    ip_rows = []
    model.IPositiveSet = defaultdict(IPList)
    for uid in range(m):
        row = []
        name = users_by_id[uid]
        if name in model.PositiveSet and uid % 5 != 0:
            own = set(model.PositiveSet[name])
            for f in ((uid * 7 + 3) % m, (uid * 11 + 5) % m):
                fr = users_by_id[f]
                if fr in model.PositiveSet and fr != name:
                    row += sorted(t2i[t] for t in set(model.PositiveSet[fr]).difference(own))
            model.IPositiveSet[name] = IPList(i2t[t] for t in row)
        ip_rows.append(row)
    ip_indptr = np.zeros(m + 1, dtype=np.int64)
    np.cumsum([len(r) for r in ip_rows], out=ip_indptr[1:])
    ip_items = np.array([t for r in ip_rows for t in r], dtype=np.int32)

    # the draws, in the order the loop asks for them: per event, per repeat n: [k] then the attempts for j
    item_names = list(t2i.keys())
    streams = []
    kpos_all, neg_all = [], []
    for ep in range(ITERS):
        kp = [cune_ref.sample_implicit(SEED, ep, nn, ev_user, ip_indptr) for nn in range(3)]
        att = [philox.attempt_stream(SEED, ep, ev_user, n, uq_indptr, uq_items, slot=nn) for nn in range(3)]
        tries = [philox.sample_negatives(SEED, ep, ev_user, n, uq_indptr, uq_items, slot=nn, return_attempts=True)[1] for nn in range(3)]
        cur = [0, 0, 0]
        for e in range(len(ev_user)):
            for nn in range(3):
                if kp[nn][e] >= 0:
                    streams.append(('k', int(kp[nn][e])))
                for _ in range(int(tries[nn][e])):
                    streams.append(('j', att[nn][0][cur[nn]]))
                    cur[nn] += 1
        kpos_all.append(np.stack(kp))
        neg_all.append(np.stack([a[1] for a in att]))
    cursor = [0]

    def fake_choice(lst):
        kind, v = streams[cursor[0]]
        cursor[0] += 1
        assert (kind == 'k') == isinstance(lst, IPList), (kind, type(lst))
        return lst[v]

    ns = {"choice": fake_choice, "sigmoid": sigmoid, "log": log}
    exec(body, ns)
    RefCUNE.train = ns["train"]
    P0, Q0 = model.P.copy(), model.Q.copy()
    with redirect_stdout(io.StringIO()):
        model.train()
    assert cursor[0] == len(streams), (cursor[0], len(streams))
    assert len(snaps) == ITERS and model.P.dtype == np.float32
    # the event order of the loop is PositiveSet's = userRecord's = the event CSR's (record_ref); item_names[v] is track v
    assert item_names == [i2t[v] for v in range(n)]
    np.savez_compressed(os.path.join(OUT, "cune_small.npz"), seed=np.int64(SEED), P0=P0, Q0=Q0, s=np.float64(S),
                        lr=np.float64(model.lRate), regU=np.float64(model.regU), regI=np.float64(model.regI),
                        ev_indptr=ev_indptr, ev_items=ev_items, uq_indptr=uq_indptr, uq_items=uq_items,
                        ip_indptr=ip_indptr, ip_items=ip_items, kpos=np.stack(kpos_all), neg=np.stack(neg_all),
                        P=np.stack([x[0] for x in snaps]), Q=np.stack([x[1] for x in snaps]),
                        loss=np.array([x[2] for x in snaps]))
    print("wrote cune_small.npz:", m, "users", n, "tracks", len(ev_items), "events,", int((np.diff(ip_indptr) == 0).sum()),
          "users without implicit positives,", os.path.getsize(os.path.join(OUT, "cune_small.npz")), "B")


if __name__ == "__main__":
    main()
