#!/usr/bin/env python
"""Generate tests/golden/* by running the REFERENCE'S OWN CODE in this container.

Test infrastructure.  Needs /root/reference (read-only) on sys.path, so it only runs in the
build container; the GPU box consumes the committed outputs.  Nothing from the reference is
copied into the repo: its modules are imported, and the SGD loop -- which the reference ships
as a commented-out string, recommender/cf/BPR.py:30-63 -- is read from that file at run time,
exec'd, and attached to a subclass of the reference's own IterativeRecommender.

    python oracle/make_golden.py            # rewrites tests/golden/

Outputs
  config_cases.json   tool/config.py  Config / LineConfig parses
  record_small.json   data/record.py  Record (explicit train/test and -byTime modes)
  sgd_small.npz       BPR.py:31-62 loop, 3 epochs incl. the lr schedule
                      (IterativeRecommender.py:47-75), negatives fed from the Philox stream
  eval_small.npz      IterativeRecommender.evalRanking (the shipped lossy selection) + Measure
  measure_small.json  evaluation/measure.py on exact top-N lists
  sigmoid.json        tool/qmath.py sigmoid
"""
import io
import json
import os
import sys
import tempfile
import textwrap
from contextlib import redirect_stdout

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
REF = os.environ.get("YUE_REFERENCE", "/root/reference")
sys.path.insert(0, ROOT)
sys.path.insert(1, REF)
OUT = os.path.join(ROOT, "tests", "golden")

from oracle import philox, record_ref, topn as otopn  # noqa: E402

SEED = 20260101


def synth_events(n_users, n_tracks, plays, seed):
    """Small time-ordered log as list-of-dict events in -columns order user,track,artist,time."""
    rng = np.random.Generator(np.random.Philox(seed))
    w = np.arange(1, n_users + 1) ** -0.8
    deg = np.maximum(2, np.floor(w * plays / w.sum())).astype(int)
    users = np.repeat(np.arange(n_users), deg)
    cdf = np.cumsum(np.arange(1, n_tracks + 1) ** -1.0)
    cdf /= cdf[-1]
    items = np.minimum(np.searchsorted(cdf, rng.random(len(users)), side="right"), n_tracks - 1)
    order = rng.permutation(len(users))
    evs = []
    for t, e in enumerate(order):
        evs.append({'user': 'u%d' % users[e], 'track': 't%d' % items[e],
                    'artist': 'a%d' % (items[e] % 17), 'time': str(1500000000 + t)})
    return evs


def write_conf(path, record_path, out_dir, eval_setup, k=10, iters=3, topn="5,10"):
    with open(path, "w") as f:
        f.write("record=%s\n" % record_path)
        f.write("record.setup=-columns user:1,track:2,artist:3,time:0 -delim ,\n")
        f.write("recommender=BPR\n")
        f.write("evaluation.setup=%s\n" % eval_setup)
        f.write("item.ranking=-topN %s\n" % topn)
        f.write("num.factors=%d\n" % k)
        f.write("num.max.iter=%d\n" % iters)
        f.write("learnRate=-init 0.02 -max 1\n")
        f.write("reg.lambda=-u 0.01 -i 0.01 -b 0.2 -s 0.2\n")
        f.write("output.setup=on -dir %s/\n" % out_dir)


def reference_sgd_method():
    """The text of BPR.py between the two ''' markers, as a function object."""
    src = open(os.path.join(REF, "recommender", "cf", "BPR.py"), encoding="utf8").read()
    a = src.index("'''")
    b = src.index("'''", a + 3)
    body = textwrap.dedent(src[a + 3:b])
    return body


def main():
    os.makedirs(OUT, exist_ok=True)
    from tool.config import Config, LineConfig
    from tool.qmath import sigmoid
    from data.record import Record
    from base.IterativeRecommender import IterativeRecommender
    from evaluation.measure import Measure

    tmp = tempfile.mkdtemp(prefix="yue_golden_")

    # ---------------------------------------------------------------- config cases
    line_cases = [
        "-columns user:1,track:2,artist:3,time:0 -delim ,",
        "-target track -byTime 0.2",
        "-init 0.02 -max 1",
        "-u 0.01 -i 0.01 -b 0.2 -s 0.2",
        "on -dir ./results/",
        "off -dir ./results/APR/",
        "-topN 5,10",
        "-regA 2 -eps 0.5 -advEpoch 100",
        "-target track -cv 5 -p",
        "-ap 0.2 -b 1 -cold 5 -sample",
        "-k -5 -x -0.5 3",
        "-a  b   -c",
        "-testSet ./dataset/test file.txt -target artist",
        "",
    ]
    cases = []
    for s in line_cases:
        lc = LineConfig(s)
        cases.append({"line": s, "options": {k: v for k, v in lc.options.items()},
                      "main": lc.isMainOn()})
    conf_text = ("record=./dataset/log.txt\nrecommender=BPR\n\nnum.factors=10\n"
                 "bad line without equals\nlearnRate=-init 0.02 -max 1\n")
    cpath = os.path.join(tmp, "case.conf")
    open(cpath, "w").write(conf_text)
    with redirect_stdout(io.StringIO()):
        cfg = Config(cpath)
    json.dump({"line_cases": cases, "conf_text": conf_text, "conf": cfg.config},
              open(os.path.join(OUT, "config_cases.json"), "w"), indent=1)

    # ---------------------------------------------------------------- sigmoid
    xs = [-30.0, -5.5, -1.0, -1e-3, 0.0, 1e-3, 0.25, 1.0, 7.0, 30.0]
    json.dump({"x": xs, "y": [sigmoid(x) for x in xs]},
              open(os.path.join(OUT, "sigmoid.json"), "w"))

    # ---------------------------------------------------------------- record
    events = synth_events(200, 400, 6000, SEED)
    rng = np.random.Generator(np.random.Philox(SEED + 1))
    held = rng.random(len(events)) < 0.2
    train = [e for e, h in zip(events, held) if not h]
    test = [e for e, h in zip(events, held) if h]
    conf_path = os.path.join(tmp, "bpr.conf")
    write_conf(conf_path, os.path.join(tmp, "log.txt"), os.path.join(tmp, "res"),
               "-target track -ap 0.2")
    with redirect_stdout(io.StringIO()):
        conf = Config(conf_path)
        rec = Record(conf, train, test)
    # -byTime mode: Record splits internally (record.py:108-123)
    conf_bt_path = os.path.join(tmp, "bpr_bt.conf")
    write_conf(conf_bt_path, os.path.join(tmp, "log.txt"), os.path.join(tmp, "res"),
               "-target track -byTime 0.2")
    with redirect_stdout(io.StringIO()):
        conf_bt = Config(conf_bt_path)
        rec_bt = Record(conf_bt, events[:1500], [])
    json.dump({
        "events": events, "held": [bool(h) for h in held],
        "name2id": {k: dict(v) for k, v in rec.name2id.items()},
        "userRecord_order": list(rec.userRecord.keys()),
        "userRecord_len": [len(v) for v in rec.userRecord.values()],
        "testSet": {u: dict(d) for u, d in rec.testSet.items()},
        "testSet_order": list(rec.testSet.keys()),
        "recordCount": rec.recordCount,
        "byTime": {
            "n_events": 1500,
            "name2id": {k: dict(v) for k, v in rec_bt.name2id.items()},
            "userRecord": {u: [e['track'] for e in v] for u, v in rec_bt.userRecord.items()},
            "testSet": {u: dict(d) for u, d in rec_bt.testSet.items()},
            "testSet_order": list(rec_bt.testSet.keys()),
            "recordCount": rec_bt.recordCount,
        },
    }, open(os.path.join(OUT, "record_small.json"), "w"))

    # ---------------------------------------------------------------- SGD: the reference's loop text
    body = reference_sgd_method()
    assert "def buildModel(self):" in body and "sigmoid(" in body
    name2id = {k: dict(v) for k, v in rec.name2id.items()}
    ev_indptr, ev_items, uq_indptr, uq_items = record_ref.interaction_arrays(name2id, rec.userRecord)
    ev_user = record_ref.ev_users(ev_indptr)
    n_items = len(name2id['track'])
    MAXIT = 3
    streams, negs = [], []
    for ep in range(MAXIT):
        flat, j = philox.attempt_stream(SEED, ep, ev_user, n_items, uq_indptr, uq_items)
        streams += flat
        negs.append(j)
    cursor = [0]

    def fake_choice(lst):
        v = streams[cursor[0]]
        cursor[0] += 1
        return lst[v]

    from collections import defaultdict
    from math import log
    ns = {"defaultdict": defaultdict, "choice": fake_choice, "sigmoid": sigmoid, "log": log}
    exec(body, ns)

    snaps = []

    class RefBPR(IterativeRecommender):
        def initModel(self):
            super(RefBPR, self).initModel()
            self.m = self.data.getSize('user')
            self.n = self.data.getSize(self.recType)
            self.train_size = len(self.data.trainingData)

        def isConverged(self, it):
            snaps.append((self.P.copy(), self.Q.copy(), float(self.loss), float(self.lRate)))
            return super(RefBPR, self).isConverged(it)

    RefBPR.buildModel = ns["buildModel"]

    with redirect_stdout(io.StringIO()):
        model = RefBPR(conf, train, test)
        model.readConfiguration()
        np.random.seed(1234)
        model.initModel()
        P0, Q0 = model.P.copy(), model.Q.copy()
        model.buildModel()
    assert cursor[0] == len(streams), (cursor[0], len(streams))
    assert model.P.dtype == np.float32
    np.savez_compressed(
        os.path.join(OUT, "sgd_small.npz"),
        seed=np.int64(SEED), P0=P0, Q0=Q0,
        ev_indptr=ev_indptr, ev_items=ev_items, uq_indptr=uq_indptr, uq_items=uq_items,
        neg=np.stack(negs),
        P=np.stack([s[0] for s in snaps]), Q=np.stack([s[1] for s in snaps]),
        loss=np.array([s[2] for s in snaps]), lr_used=np.array([s[3] for s in snaps]),
        lr_final=np.float64(model.lRate), lr_init=np.float64(0.02), max_lr=np.float64(1.0),
        regU=np.float64(model.regU), regI=np.float64(model.regI))

    # ---------------------------------------------------------------- evalRanking (shipped selection)
    with redirect_stdout(io.StringIO()):
        rec_lists = {}
        orig_rm = Measure.rankingMeasure

        def spy(origin, res, N, itemCount):
            rec_lists.update({u: list(v) for u, v in res.items()})
            return orig_rm(origin, res, N, itemCount)
        Measure.rankingMeasure = staticmethod(spy)
        try:
            model.evalRanking()
        finally:
            Measure.rankingMeasure = staticmethod(orig_rm)
    t2i, u2i = name2id['track'], name2id['user']
    test_users = list(model.data.testSet.keys())
    quirk_ids = np.array([[t2i[x] for x in rec_lists[u]] for u in test_users], dtype=np.int32)
    blas_scores = np.stack([model.predict(u) for u in test_users[:40]])
    test_indptr = np.zeros(len(test_users) + 1, dtype=np.int64)
    test_items = []
    for k, u in enumerate(test_users):
        test_items += [t2i[x] for x in model.data.testSet[u]]
        test_indptr[k + 1] = len(test_items)
    np.savez_compressed(
        os.path.join(OUT, "eval_small.npz"),
        P=model.P, Q=model.Q, test_users=np.array([u2i[u] for u in test_users], dtype=np.int32),
        test_indptr=test_indptr, test_items=np.array(test_items, dtype=np.int32),
        quirk_ids=quirk_ids, blas_scores=blas_scores, topN=np.array([5, 10]),
        uq_indptr=uq_indptr, uq_items=uq_items)

    # ---------------------------------------------------------------- Measure on exact top-N
    uid = np.array([u2i[u] for u in test_users])
    ids, _ = otopn.topn_exact(model.P, model.Q, uid, 10, uq_indptr, uq_items)
    i2t = {v: k for k, v in t2i.items()}
    exact_lists = {u: [i2t[int(x)] for x in ids[k]] for k, u in enumerate(test_users)}
    with redirect_stdout(io.StringIO()):
        m_exact = Measure.rankingMeasure(model.data.testSet, exact_lists, [5, 10], n_items)
        m_quirk = list(model.measure)
    json.dump({"measure_exact": m_exact, "measure_quirk": m_quirk,
               "exact_ids": ids.tolist(), "item_count": n_items},
              open(os.path.join(OUT, "measure_small.json"), "w"))
    print("golden written to", OUT)
    for f in sorted(os.listdir(OUT)):
        print("  %-22s %8d B" % (f, os.path.getsize(os.path.join(OUT, f))))


if __name__ == "__main__":
    main()
