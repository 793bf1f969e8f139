#!/usr/bin/env python
"""Generate tests/golden/lightgcn_small.npz by running the REFERENCE's own LightGCN -- recommender/advanced/LightGCN.py and
base/DeepRecommender, both unmodified, imported from the reference tree -- over oracle/tf1_shim.py (a stand-in for the
TensorFlow 1.x calls those two files make, on torch autograd in float64) with `random.randint` replaced by the Philox
streams of oracle/philox.py.  Test infrastructure; build container only (the reference tree is read at run time).

What this pins (see the shim's header for what it cannot): the graph the reference BUILDS -- the per-event adjacency entries,
which tensor feeds which layer, the sum of the normalised layers, the loss and its regulariser, Adam over U and V -- its
batch loop (slices of trainingData in file order) and its sampler (five draws per event with rejection against the user's
plays, the LAST one kept).  The generator asserts that the reference consumed exactly the flattened Philox attempt stream
(slots 0..4 of every event, every rejected candidate included), so its rejection logic is the oracle's.

    python oracle/make_golden_lightgcn.py
"""
import importlib.machinery
import importlib.util
import io
import json
import os
import sys
import tempfile
from contextlib import redirect_stdout

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
REF = os.environ.get("YUE_REFERENCE", "/root/reference")
sys.path.insert(0, ROOT)
sys.path.insert(1, REF)
OUT = os.path.join(ROOT, "tests", "golden")

from oracle import philox, tf1_shim  # noqa: E402

SEED, K, ITERS, BATCH, LR, REG = 20260109, 10, 2, 128, 0.002, 0.001


class PhiloxRandom(object):
    """stands in for the `random` module inside LightGCN.py: randint pops the next candidate of the attempt stream"""

    def __init__(self, n):
        self.n, self.stream, self.pos = n, [], 0

    def randint(self, a, b):
        assert (a, b) == (0, self.n - 1), (a, b)
        v = self.stream[self.pos]
        self.pos += 1
        return v


def attempt_streams(seed, epoch, ev_user, n, uq_indptr, uq_items):
    """per event, the candidates the reference's five draw-and-redraw loops see, in call order; and the kept fifth draw"""
    per_slot = []
    for slot in range(5):
        j, tries = philox.sample_negatives(seed, epoch, ev_user, n, uq_indptr, uq_items, slot=slot, return_attempts=True)
        cands = [[int(philox.draw_item(seed, epoch, np.array([e]), slot, np.array([t]), n)[0]) for t in range(int(tries[e]))]
                 for e in range(len(ev_user))]
        assert all(c[-1] == int(j[e]) for e, c in enumerate(cands))
        per_slot.append((cands, j))
    flat = [c for e in range(len(ev_user)) for slot in range(5) for c in per_slot[slot][0][e]]
    return flat, per_slot[4][1]


def main():
    sys.modules["tensorflow"] = tf1_shim
    loader = importlib.machinery.SourceFileLoader("base.DeepRecommender", os.path.join(REF, "base", "DeepRecommender"))
    spec = importlib.util.spec_from_loader("base.DeepRecommender", loader)
    deep = importlib.util.module_from_spec(spec)
    sys.modules["base.DeepRecommender"] = deep
    loader.exec_module(deep)                                    # the file has no .py extension; its text is unmodified
    import recommender.advanced.LightGCN as ref_mod             # the reference's module, unmodified
    from tool.config import Config
    from yue_b200.host.record import interaction_arrays

    g = json.load(open(os.path.join(OUT, "record_small.json")))
    keep = set("u%d" % x for x in range(90))
    train = [e for e, h in zip(g["events"], g["held"]) if not h and e["user"] in keep]
    test = [e for e, h in zip(g["events"], g["held"]) if h and e["user"] in keep]
    tmp = tempfile.mkdtemp(prefix="yue_golden_lightgcn_")
    cpath = os.path.join(tmp, "lightgcn.conf")
    with open(cpath, "w") as f:
        f.write("record=%s\nrecord.setup=-columns user:1,track:2,artist:3,time:0 -delim ,\nrecommender=LightGCN\n"
                "evaluation.setup=-target track -ap 0.2\nitem.ranking=-topN 5,10\nnum.factors=%d\nnum.max.iter=%d\nbatch_size=%d\n"
                "learnRate=-init %g -max 1\nreg.lambda=-u %g -i 0.001 -b 0.2 -s 0.2\noutput.setup=on -dir %s/res/\n"
                % (os.path.join(tmp, "log.txt"), K, ITERS, BATCH, LR, REG, tmp))
    tf1_shim.reset()
    out = io.StringIO()
    with redirect_stdout(out):
        model = ref_mod.LightGCN(Config(cpath), train, test)
        model.readConfiguration()
        np.random.seed(99)
        model.initModel()                                       # DeepRecommender's variables + LightGCN's graph
    m, n = model.m, model.n
    uid, tid = model.data.name2id["user"], model.data.name2id["track"]
    ev_user = np.array([uid[e["user"]] for e in model.data.trainingData], dtype=np.int32)
    ev_item = np.array([tid[e["track"]] for e in model.data.trainingData], dtype=np.int32)
    ev_indptr, ev_items, uq_indptr, uq_items = interaction_arrays(model.data.name2id, model.data.userRecord, "track")
    U0, V0 = model.U.tensor.detach().numpy().copy(), model.V.tensor.detach().numpy().copy()
    assert U0.shape == (m, K) and V0.shape == (n, K) and np.abs(U0).max() <= 0.01
    assert model.n_layers == 3 and model.negativeCount == 5 and model.batch_size == BATCH

    fake = PhiloxRandom(n)
    kept = []
    for ep in range(ITERS):
        flat, j = attempt_streams(SEED, ep, ev_user, n, uq_indptr, uq_items)
        fake.stream += flat
        kept.append(j)
    ref_mod.random = fake                                       # LightGCN.py: `import random` ... random.randint(0, self.n-1)
    with redirect_stdout(out):
        model.buildModel()
    assert fake.pos == len(fake.stream), "the reference's sampler did not consume exactly the Philox attempt stream"
    losses = [float(l.split("loss:")[1]) for l in out.getvalue().splitlines() if l.startswith("training:")]
    steps = (len(ev_user) + BATCH - 1) // BATCH
    assert len(losses) == ITERS * steps
    U, V = model.U.tensor.detach().numpy().copy(), model.V.tensor.detach().numpy().copy()
    FU = model.sess.run(model.multi_user_embeddings)
    FV = model.sess.run(model.multi_item_embeddings)
    some = [u for u in list(uid)[:5]]
    preds = np.stack([model.predict(u) for u in some])
    assert np.allclose(preds, FU[[uid[u] for u in some]] @ FV.T)
    np.savez_compressed(os.path.join(OUT, "lightgcn_small.npz"), m=m, n=n, k=K, iters=ITERS, batch=BATCH, lr=LR, reg=REG, seed=SEED,
                        ev_user=ev_user, ev_item=ev_item, ev_indptr=ev_indptr, ev_items=ev_items, uq_indptr=uq_indptr, uq_items=uq_items,
                        U0=U0, V0=V0, losses=np.array(losses), U=U, V=V, FU=FU, FV=FV, neg=np.stack(kept),
                        pred_users=np.array([uid[u] for u in some]), preds=preds)
    print("lightgcn_small.npz: %d users x %d tracks, %d training events, %d steps, loss %.4f -> %.4f; sampler consumed %d draws"
          % (m, n, len(ev_user), len(losses), losses[0], losses[-1], fake.pos))


if __name__ == "__main__":
    main()
