#!/usr/bin/env python
"""Generate tests/golden/wrmf_small.npz by running the REFERENCE'S OWN WRMF class, unmodified, in this container.

Test infrastructure.  Needs /root/reference on sys.path (build container only); the GPU box consumes the committed
output.  recommender/cf/WRMF.py imports cleanly here (numpy + scipy.sparse), so -- unlike BPR, whose loop is a
comment -- the golden is the output of the class itself: `initModel` under a seeded global numpy stream, then
`buildModel` for two iterations on the log of tests/golden/record_small.json, with X/Y captured after each
iteration through the `print('iteration:', ...)` the loop ends with (WRMF.py:83).

    python oracle/make_golden_wrmf.py
"""
import builtins
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

from oracle import record_ref, wrmf_ref  # noqa: E402


def main():
    from tool.config import Config
    from recommender.cf.WRMF import WRMF

    g = json.load(open(os.path.join(OUT, "record_small.json")))
    train = [e for e, h in zip(g["events"], g["held"]) if not h]
    test = [e for e, h in zip(g["events"], g["held"]) if h]
    tmp = tempfile.mkdtemp(prefix="yue_golden_wrmf_")
    K, ITERS, REG = 20, 2, 1.0
    cpath = os.path.join(tmp, "wrmf.conf")
    with open(cpath, "w") as f:                       # the keys of config/WRMF.conf
        f.write("record=%s\nrecord.setup=-columns user:1,track:2,artist:3,time:0 -delim ,\nrecommender=WRMF\n"
                "evaluation.setup=-target track -ap 0.2\nitem.ranking=-topN 5,10\nnum.factors=%d\nnum.max.iter=%d\n"
                "learnRate=-init 0.02 -max 1\nreg.lambda=-u %g -i 0.1 -b 0.2 -s 0.2\noutput.setup=on -dir %s/res/\n"
                % (os.path.join(tmp, "log.txt"), K, ITERS, REG, tmp))
    snaps = []
    real_print = builtins.print

    with redirect_stdout(io.StringIO()):
        model = WRMF(Config(cpath), train, test)
        model.readConfiguration()
        np.random.seed(4321)
        model.initModel()
        X0, Y0 = model.X.copy(), model.Y.copy()

        def spy(*args, **kw):                         # WRMF.py:83 prints once per iteration
            if args and args[0] == 'iteration:':
                snaps.append((model.X.copy(), model.Y.copy(), float(args[3])))
            return real_print(*args, **kw)
        builtins.print = spy
        try:
            model.buildModel()
        finally:
            builtins.print = real_print
    assert len(snaps) == ITERS and model.X.dtype == np.float32
    name2id = {k: dict(v) for k, v in model.data.name2id.items()}
    ev_indptr, ev_items, uq_indptr, uq_items = record_ref.interaction_arrays(name2id, model.data.userRecord)
    cnt = wrmf_ref.pair_counts(ev_indptr, ev_items, uq_indptr, uq_items)
    # the restated counts against the reference's own containers (record.py:160-163)
    t2i, u2i = name2id['track'], name2id['user']
    for t, users in model.data.listened['track'].items():
        for u, c in users.items():
            row = uq_items[uq_indptr[u2i[u]]:uq_indptr[u2i[u] + 1]]
            assert cnt[uq_indptr[u2i[u]] + np.searchsorted(row, t2i[t])] == c
    scores = np.stack([model.predict(u) for u in list(model.data.testSet)[:20]])
    np.savez_compressed(os.path.join(OUT, "wrmf_small.npz"), X0=X0, Y0=Y0, reg=np.float64(REG),
                        X=np.stack([s[0] for s in snaps]), Y=np.stack([s[1] for s in snaps]),
                        loss=np.array([s[2] for s in snaps]), ev_indptr=ev_indptr, ev_items=ev_items,
                        uq_indptr=uq_indptr, uq_items=uq_items, counts=cnt,
                        score_users=np.array([u2i[u] for u in list(model.data.testSet)[:20]], dtype=np.int32),
                        scores=scores)
    print("wrote", os.path.join(OUT, "wrmf_small.npz"), os.path.getsize(os.path.join(OUT, "wrmf_small.npz")), "B")


if __name__ == "__main__":
    main()
