"""The oracle against the golden vectors produced by the reference's own code
(oracle/make_golden.py).  CPU only."""
import json
import os

import numpy as np
import pytest

from oracle import bpr_ref, metrics, philox, record_ref, topn


def test_philox_kat():
    for ctr, key, exp in philox.PHILOX_KAT:
        out = philox.philox4x32_10(*[np.array([c]) for c in ctr], key[0], key[1])
        assert tuple(int(x[0]) for x in out) == exp


def test_sigmoid_golden(golden_dir):
    g = json.load(open(os.path.join(golden_dir, "sigmoid.json")))
    for x, y in zip(g["x"], g["y"]):
        assert bpr_ref.sigmoid(x) == y


def test_sampler_never_returns_played_and_is_uniform():
    n = 50
    uq_indptr = np.array([0, 10, 10])
    uq_items = np.arange(0, 20, 2).astype(np.int32)          # user 0 played the even ids < 20
    ev_user = np.zeros(40000, dtype=np.int64)
    j = philox.sample_negatives(99, 3, ev_user, n, uq_indptr, uq_items)
    assert not np.isin(j, uq_items).any()
    counts = np.bincount(j, minlength=n)[np.setdiff1d(np.arange(n), uq_items)]
    expected = len(j) / 40.0
    chi2 = ((counts - expected) ** 2 / expected).sum()
    assert chi2 < 80.0                                       # 39 dof, p ~ 1e-4
    # pure function of (seed, epoch, event): a shard sees the same draws
    j2 = philox.sample_negatives(99, 3, ev_user[100:200], n, uq_indptr, uq_items, event_base=100)
    assert (j2 == j[100:200]).all()


def test_sgd_epochs_match_reference_loop(golden_dir):
    g = np.load(os.path.join(golden_dir, "sgd_small.npz"))
    P, Q = g["P0"].copy(), g["Q0"].copy()
    ev_user = record_ref.ev_users(g["ev_indptr"])
    for ep in range(g["neg"].shape[0]):
        neg = philox.sample_negatives(int(g["seed"]), ep, ev_user, Q.shape[0], g["uq_indptr"], g["uq_items"])
        assert (neg == g["neg"][ep]).all()
    P, Q = g["P0"].copy(), g["Q0"].copy()
    hist = bpr_ref.train(P, Q, ev_user, g["ev_items"], Q.shape[0], g["uq_indptr"], g["uq_items"],
                         3, float(g["lr_init"]), float(g["max_lr"]), float(g["regU"]), float(g["regI"]),
                         int(g["seed"]))
    assert np.array_equal(P, g["P"][-1]) and np.array_equal(Q, g["Q"][-1])     # bit for bit
    assert [h[0] for h in hist] == list(g["loss"])
    assert [h[1] for h in hist] == list(g["lr_used"])


def test_sgd_float64_rederivation_agrees(golden_dir):
    g = np.load(os.path.join(golden_dir, "sgd_small.npz"))
    ev_user = record_ref.ev_users(g["ev_indptr"])
    P, Q = g["P0"].astype(np.float64), g["Q0"].astype(np.float64)
    bpr_ref.sgd_epoch(P, Q, ev_user, g["ev_items"], g["neg"][0], 0.02, 0.01, 0.01, mode="float64")
    assert np.allclose(P, g["P"][0], rtol=1e-5, atol=1e-7)
    assert np.allclose(Q, g["Q"][0], rtol=1e-5, atol=1e-7)


def test_ref_quirk_selection_matches_reference(golden_dir):
    e = np.load(os.path.join(golden_dir, "eval_small.npz"))
    for b, u in enumerate(e["test_users"][:40]):
        m = e["uq_items"][e["uq_indptr"][u]:e["uq_indptr"][u + 1]]
        assert topn.topn_ref_quirk(e["blas_scores"][b], m, 10) == list(e["quirk_ids"][b])


def test_measure_matches_reference(golden_dir):
    e = np.load(os.path.join(golden_dir, "eval_small.npz"))
    mj = json.load(open(os.path.join(golden_dir, "measure_small.json")))
    origin = [e["test_items"][e["test_indptr"][k]:e["test_indptr"][k + 1]].tolist()
              for k in range(len(e["test_users"]))]
    assert metrics.ranking_measure(origin, mj["exact_ids"], [5, 10], mj["item_count"]) == mj["measure_exact"]
    assert metrics.ranking_measure(origin, e["quirk_ids"].tolist(), [5, 10], mj["item_count"]) == mj["measure_quirk"]


def test_exact_topn_golden_and_fma_scores(golden_dir):
    e = np.load(os.path.join(golden_dir, "eval_small.npz"))
    mj = json.load(open(os.path.join(golden_dir, "measure_small.json")))
    ids, sc = topn.topn_exact(e["P"], e["Q"], e["test_users"], 10, e["uq_indptr"], e["uq_items"])
    assert ids.tolist() == mj["exact_ids"]
    assert (np.diff(sc, axis=1) <= 0).all()
    # canonical FMA-chain scores stay within 1e-3 relative of numpy's dot (north_star tolerance)
    s = topn.scores_fma32(e["P"][e["test_users"][:40]], e["Q"])
    assert np.allclose(s, e["blas_scores"], rtol=1e-3, atol=1e-6)
    # masked ids never appear
    for b, u in enumerate(e["test_users"]):
        assert not np.isin(ids[b], e["uq_items"][e["uq_indptr"][u]:e["uq_indptr"][u + 1]]).any()


def test_record_restatement_matches_reference(golden_dir):
    g = json.load(open(os.path.join(golden_dir, "record_small.json")))
    train = [e for e, h in zip(g["events"], g["held"]) if not h]
    test = [e for e, h in zip(g["events"], g["held"]) if h]
    name2id, user_record, test_set = record_ref.preprocess(train, test)
    assert name2id == g["name2id"]
    assert list(user_record.keys()) == g["userRecord_order"]
    assert [len(v) for v in user_record.values()] == g["userRecord_len"]
    assert list(test_set.keys()) == g["testSet_order"]
    assert {u: dict(d) for u, d in test_set.items()} == g["testSet"]


def test_ndcg_definition():
    origin = [[1, 2, 3], [9]]
    rec = [[1, 7, 3, 8], [4, 9]]
    import math
    d0 = (1 / math.log2(2) + 1 / math.log2(4)) / (1 / math.log2(2) + 1 / math.log2(3) + 1 / math.log2(4))
    d1 = (1 / math.log2(3)) / 1.0
    assert metrics.ndcg(origin, rec, 4) == pytest.approx((d0 + d1) / 2)


def test_c_oracle_agrees_with_numpy_emulation():
    """oracle/csrc/oracle.c scores with a true fmaf(); oracle/topn.py emulates the chain in float64.  They must give the
    same scores and the same masked top-N (ids and order, ties by id) -- including on factors with exact ties."""
    import shutil
    if shutil.which("gcc") is None:
        pytest.skip("no C compiler")
    from oracle import c_oracle, topn
    from yue_b200 import synth
    rng = np.random.default_rng(4)
    for d, n, m in ((64, 3000, 40), (10, 500, 25), (37, 1200, 16)):
        P = (rng.normal(size=(m, d)) * rng.lognormal(0, 0.5, (m, 1))).astype(np.float32)
        Q = (rng.normal(size=(n, d)) * 0.3).astype(np.float32)
        Q[100:130] = Q[99]                       # a run of identical tracks: equal scores, order by id
        indptr, uq = synth.mask_csr(m, n, 20, seed=d)
        users = np.arange(m, dtype=np.int32)
        assert np.array_equal(c_oracle.scores_fma32(P, Q), topn.scores_fma32(P, Q))
        for N in (1, 10, 50):
            ci, cs = c_oracle.topn_exact(P, Q, users, N, indptr, uq)
            ni, ns = topn.topn_exact(P, Q, users, N, indptr, uq)
            assert np.array_equal(ci, ni) and np.array_equal(cs, ns)


def test_wrmf_oracle_reproduces_the_reference_class(golden_dir):
    """tests/golden/wrmf_small.npz is the output of recommender/cf/WRMF.py itself (oracle/make_golden_wrmf.py).  With the
    float32 Gram matrix the restatement gives the same float32 factors bit for bit; with the float64 Gram matrix (what
    the CUDA kernel computes) it stays within 1e-5 per row."""
    from oracle import wrmf_ref
    g = np.load(os.path.join(golden_dir, "wrmf_small.npz"))
    m, n = g["X0"].shape[0], g["Y0"].shape[0]
    cnt = wrmf_ref.pair_counts(g["ev_indptr"], g["ev_items"], g["uq_indptr"], g["uq_items"])
    assert np.array_equal(cnt, g["counts"]) and cnt.sum() == len(g["ev_items"])
    itp, itu, itc = wrmf_ref.transpose(m, n, g["uq_indptr"], g["uq_items"], cnt)
    assert itp[-1] == len(cnt) and all(np.all(np.diff(itu[itp[t]:itp[t + 1]]) > 0) for t in range(n))
    X, Y = g["X0"].copy(), g["Y0"].copy()
    X64, Y64 = g["X0"].copy(), g["Y0"].copy()
    for it in range(len(g["loss"])):
        loss = wrmf_ref.iteration(X, Y, g["uq_indptr"], g["uq_items"], cnt, itp, itu, itc, float(g["reg"]), gram="f32")
        assert np.array_equal(X, g["X"][it]) and np.array_equal(Y, g["Y"][it])
        assert loss == pytest.approx(float(g["loss"][it]), rel=1e-7)
        wrmf_ref.iteration(X64, Y64, g["uq_indptr"], g["uq_items"], cnt, itp, itu, itc, float(g["reg"]), gram="f64")
        for a, b in ((X64, X), (Y64, Y)):
            d = np.linalg.norm(a.astype(np.float64) - b, axis=1) / np.maximum(np.linalg.norm(b, axis=1), 1e-12)
            assert d.max() < 1e-5


def test_cune_oracle_reproduces_the_reference_loop(golden_dir):
    """tests/golden/cune_small.npz is the output of the loop text of recommender/advanced/CUNE.py:119-178 itself, exec'd
    on the reference's own IterativeRecommender with the Philox draws in place of random.choice
    (oracle/make_golden_cune.py).  The restatement (oracle/cune_ref.py: groundwork for the next model of SURVEY 8f row 4)
    gives the same float32 tables bit for bit over two iterations, both branches of the loop included, and the draws
    are reproducible from the seed."""
    from oracle import cune_ref
    g = np.load(os.path.join(golden_dir, "cune_small.npz"))
    ev_user = record_ref.ev_users(g["ev_indptr"])
    n = g["Q0"].shape[0]
    deg_ip = np.diff(g["ip_indptr"])
    assert (deg_ip == 0).any() and (deg_ip > 0).any()
    P, Q = g["P0"].copy(), g["Q0"].copy()
    for it in range(len(g["loss"])):
        for nn in range(3):
            kp = cune_ref.sample_implicit(int(g["seed"]), it, nn, ev_user, g["ip_indptr"])
            assert np.array_equal(kp, g["kpos"][it][nn]) and (kp[deg_ip[ev_user] > 0] < deg_ip[ev_user][deg_ip[ev_user] > 0]).all()
            j = philox.sample_negatives(int(g["seed"]), it, ev_user, n, g["uq_indptr"], g["uq_items"], slot=nn)
            assert np.array_equal(j, g["neg"][it][nn])
        loss = cune_ref.epoch(P, Q, g["ev_indptr"], g["ev_items"], g["ip_indptr"], g["ip_items"], g["kpos"][it], g["neg"][it],
                              float(g["lr"]), float(g["regU"]), float(g["regI"]), float(g["s"]))
        assert np.array_equal(P, g["P"][it]) and np.array_equal(Q, g["Q"][it])
        assert loss == pytest.approx(float(g["loss"][it]), rel=1e-9)
