#!/usr/bin/env python
"""Generate tests/golden/apr_graph.npz by evaluating the REFERENCE's own APR graph -- recommender/advanced/APR.py and
base/DeepRecommender, both unmodified, over oracle/tf1_shim.py (their TensorFlow-1 calls on torch autograd, float64) -- on
single triplets.  Test infrastructure; build container only.

What it pins.  north_star asks for APR "reusing the triplet SGD kernel with fused perturbation": oracle/apr_ref.py takes the
reference's losses and its perturbation rule PER TRIPLET (the reference aggregates the perturbation gradient per row over a
mini-batch and trains with Adam; that optimiser and that aggregation are NOT reproduced -- DESIGN.md).  For a batch that is
one triplet the two coincide, so the reference's own graph gives, for that triplet and the current rows:
  * the perturbation `adv_U.assign(l2_normalize(grad) * eps)` (APR.py:51-60) -> the oracle's closed form  D_u = -eps d^, ...
  * `loss_adv` = softplus(-y) + regA softplus(-y_adv) with that perturbation (62-72) -> the oracle's loss and adv_score()
  * d loss_adv / d U, d V with the perturbation held constant (what `minimize(self.loss_adv)` differentiates, 76)
    -> the oracle's step: one apr_epoch() with lr and no shrink moves the rows by -lr times these gradients.

    python oracle/make_golden_apr.py
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

from oracle import tf1_shim as tf  # noqa: E402

K, EPS, REGA = 8, 0.5, 2.0


def main():
    import torch
    sys.modules["tensorflow"] = tf
    loader = importlib.machinery.SourceFileLoader("base.DeepRecommender", os.path.join(REF, "base", "DeepRecommender"))
    spec = importlib.util.spec_from_loader("base.DeepRecommender", loader)
    deep = importlib.util.module_from_spec(spec)
    sys.modules["base.DeepRecommender"] = deep
    loader.exec_module(deep)
    import recommender.advanced.APR as ref_mod                  # the reference's module, unmodified
    from tool.config import Config

    g = json.load(open(os.path.join(OUT, "record_small.json")))
    keep = set("u%d" % x for x in range(40))
    train = [e for e, h in zip(g["events"], g["held"]) if not h and e["user"] in keep]
    test = [e for e, h in zip(g["events"], g["held"]) if h and e["user"] in keep]
    tmp = tempfile.mkdtemp(prefix="yue_golden_apr_")
    cpath = os.path.join(tmp, "apr.conf")
    with open(cpath, "w") as f:
        f.write("record=%s/log.txt\nrecord.setup=-columns user:1,track:2,artist:3,time:0 -delim ,\nrecommender=APR\n"
                "evaluation.setup=-target track -ap 0.2\nitem.ranking=-topN 5,10\nnum.factors=%d\nnum.max.iter=2\nbatch_size=16\n"
                "APR=-regA %g -eps %g -advEpoch 2\nlearnRate=-init 0.003 -max 1\nreg.lambda=-u 0.002 -i 0.01 -b 0.2 -s 0.2\n"
                "output.setup=on -dir %s/res/\n" % (tmp, K, REGA, EPS, tmp))
    tf.reset()
    with redirect_stdout(io.StringIO()):
        model = ref_mod.APR(Config(cpath), train, test)
        model.readConfiguration()
        np.random.seed(1)
        model.initModel()                                       # variables, losses, adversarial update, optimisers: APR.py:25-93
    assert model.negativeCount == 3 and model.eps == EPS and model.regAdv == REGA
    rng = np.random.default_rng(7)
    m, n = model.m, model.n
    with torch.no_grad():                                       # rows of ordinary size (the truncated-normal init is 0.005)
        model.U.tensor.copy_(torch.tensor(rng.normal(0, 0.4, (m, K))))
        model.V.tensor.copy_(torch.tensor(rng.normal(0, 0.4, (n, K))))
    sess = tf.Session()
    y_node, ya_node = model._create_inference(), model._create_adv_inference()
    gU_node, gV_node = tf.gradients(model.loss_adv, [model.U, model.V])
    rows = []
    for _ in range(24):
        u, i = int(rng.integers(0, m)), int(rng.integers(0, n))
        j = int(rng.integers(0, n - 1))
        j += j >= i
        feed = {model.u_idx: [u], model.v_idx: [i], model.neg_idx: [j]}
        with torch.no_grad():
            model.adv_U.tensor.zero_()
            model.adv_V.tensor.zero_()
        sess.run([model.update_U, model.update_V], feed_dict=feed)      # APR.py:128: the perturbation of this batch
        dU, dV = model.adv_U.tensor.detach().numpy(), model.adv_V.tensor.detach().numpy()
        assert np.count_nonzero(np.abs(dU).sum(1)) == 1 and np.count_nonzero(np.abs(dV).sum(1)) == 2
        y, ya, loss_adv, gU, gV = sess.run([y_node, ya_node, model.loss_adv, gU_node, gV_node], feed_dict=feed)
        rows.append(dict(u=u, i=i, j=j, p=model.U.tensor[u].detach().numpy().copy(), qi=model.V.tensor[i].detach().numpy().copy(),
                         qj=model.V.tensor[j].detach().numpy().copy(), du=dU[u].copy(), di=dV[i].copy(), dj=dV[j].copy(),
                         y=float(y[0]), ya=float(ya[0]), loss_adv=float(loss_adv), gp=gU[u].copy(), gi=gV[i].copy(), gj=gV[j].copy()))
    np.savez_compressed(os.path.join(OUT, "apr_graph.npz"), eps=EPS, regA=REGA,
                        **{k: np.array([r[k] for r in rows]) for k in rows[0]})
    print("apr_graph.npz: %d single-triplet evaluations of the reference's APR graph (k = %d, eps %g, regA %g)" % (len(rows), K, EPS, REGA))


if __name__ == "__main__":
    main()
