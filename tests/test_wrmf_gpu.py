"""K7 (WRMF half-sweeps) against the oracle and against the output of the reference's own WRMF
class (tests/golden/wrmf_small.npz, written by oracle/make_golden_wrmf.py from recommender/cf/WRMF.py unmodified).

Tolerances.  The reference accumulates YtY / XtX with a float32 sgemm and everything else in float64; the kernel
accumulates the Gram matrix in float64 too.  oracle/wrmf_ref.py restates both: gram='f32' is bit-identical to the
reference class on the golden log, gram='f64' is what the kernel must match (1e-5 per row, relative L2), and the two
differ by ~5e-6 per row -- so against the reference itself the bound is 1e-4 per row."""
import io
import os
from contextlib import redirect_stdout

import numpy as np
import pytest

from oracle import wrmf_ref
from yue_b200 import synth

pytestmark = pytest.mark.gpu


def row_rel(a, b):
    """largest relative L2 distance between corresponding rows (rows that are zero in both count as 0)"""
    num = np.linalg.norm(a.astype(np.float64) - b.astype(np.float64), axis=1)
    den = np.maximum(np.linalg.norm(b.astype(np.float64), axis=1), 1e-12)
    return float(np.max(np.where(num == 0, 0.0, num / den)))


def csr_from_events(m, n, users, items):
    """event CSR (user-major, given order kept inside a user) + sorted-unique CSR from event pairs"""
    order = np.argsort(users, kind="stable")
    users, items = users[order], items[order]
    ev_indptr = np.zeros(m + 1, np.int64)
    np.cumsum(np.bincount(users, minlength=m), out=ev_indptr[1:])
    key = np.unique(users.astype(np.int64) * n + items)
    uq_indptr = np.zeros(m + 1, np.int64)
    np.cumsum(np.bincount(key // n, minlength=m), out=uq_indptr[1:])
    return ev_indptr, items.astype(np.int32), uq_indptr, (key % n).astype(np.int32)


def test_pair_counts_and_track_major_form(engine):
    log = synth.power_law_log(700, 900, 30000, seed=5)
    engine.set_interactions(log.m, log.n, log.ev_indptr, log.ev_items, log.uq_indptr, log.uq_items)
    cnt, itp, itu, itc = engine.wrmf_pair_counts()
    rc = wrmf_ref.pair_counts(log.ev_indptr, log.ev_items, log.uq_indptr, log.uq_items)
    rp, ru, rcn = wrmf_ref.transpose(log.m, log.n, log.uq_indptr, log.uq_items, rc)
    assert int(cnt.sum()) == log.train_size
    assert np.array_equal(cnt, rc) and np.array_equal(itp, rp) and np.array_equal(itu, ru) and np.array_equal(itc, rcn)


def test_matches_the_reference_class_on_the_golden_log(engine, golden_dir):
    g = np.load(os.path.join(golden_dir, "wrmf_small.npz"))
    m, n = g["X0"].shape[0], g["Y0"].shape[0]
    reg = float(g["reg"])
    engine.set_interactions(m, n, g["ev_indptr"], g["ev_items"], g["uq_indptr"], g["uq_items"])
    engine.set_factors(g["X0"], g["Y0"])
    itp, itu, itc = wrmf_ref.transpose(m, n, g["uq_indptr"], g["uq_items"], g["counts"])
    Xo, Yo = g["X0"].copy(), g["Y0"].copy()
    for it in range(len(g["loss"])):
        loss = engine.wrmf_sweep(0, reg, 10.0, want_loss=True)
        engine.wrmf_sweep(1, reg, 10.0)
        X, Y = engine.get_factors()
        lo = wrmf_ref.iteration(Xo, Yo, g["uq_indptr"], g["uq_items"], g["counts"], itp, itu, itc, reg, gram="f64")
        assert row_rel(X, Xo) < 1e-5 and row_rel(Y, Yo) < 1e-5, (it, row_rel(X, Xo), row_rel(Y, Yo))
        assert row_rel(X, g["X"][it]) < 1e-4 and row_rel(Y, g["Y"][it]) < 1e-4          # the reference class itself
        assert loss == pytest.approx(lo, rel=1e-6) and loss == pytest.approx(float(g["loss"][it]), rel=1e-5)
    # predict = Y.dot(X[u]) (WRMF.py:86-88) through the scoring entry point
    for u, ref_scores in zip(g["score_users"][:5], g["scores"][:5]):
        assert np.allclose(engine.predict(int(u)), ref_scores, rtol=1e-3, atol=1e-5)


@pytest.mark.parametrize("k", [8, 16, 20, 32, 50, 64, 100, 128])
def test_every_width_heavy_rows_and_empty_rows(engine, k):
    """All four tile widths; a user with > 4096 tracks and a track with > 4096 listeners (the chunked path); users and
    tracks without plays (zero rows); repeat plays (r_ui > 1)."""
    m, n = 5200, 6100
    rng = np.random.default_rng(k)
    base = synth.power_law_log(m - 100, n - 100, 60000, seed=k)                  # users/tracks >= m-100 / n-100 stay empty
    bu = np.repeat(np.arange(base.m), np.diff(base.ev_indptr))
    users = np.concatenate([bu, np.zeros(5000, np.int64), np.arange(4500), rng.integers(0, 50, 3000)])
    items = np.concatenate([base.ev_items, np.arange(5000), np.full(4500, 7), rng.integers(0, 20, 3000)])
    ev_indptr, ev_items, uq_indptr, uq_items = csr_from_events(m, n, users.astype(np.int64), items.astype(np.int64))
    cnt = wrmf_ref.pair_counts(ev_indptr, ev_items, uq_indptr, uq_items)
    itp, itu, itc = wrmf_ref.transpose(m, n, uq_indptr, uq_items, cnt)
    assert np.diff(uq_indptr).max() > 4096 and np.diff(itp).max() > 4096 and cnt.max() > 1
    X0, Y0 = synth.init_factors(m, n, k, seed=3)
    X0, Y0 = X0 * 10, Y0 * 10
    engine.set_interactions(m, n, ev_indptr, ev_items, uq_indptr, uq_items)
    engine.set_factors(X0, Y0)
    loss = engine.wrmf_sweep(0, 0.7, 10.0, want_loss=True)
    X1, _ = engine.get_factors()
    engine.wrmf_sweep(1, 0.7, 10.0)
    X, Y = engine.get_factors()
    Xo, Yo = X0.copy(), Y0.copy()
    lo = wrmf_ref.iteration(Xo, Yo, uq_indptr, uq_items, cnt, itp, itu, itc, 0.7, gram="f64")
    assert np.array_equal(X1, X)                                   # the track sweep leaves the user table alone
    assert row_rel(X, Xo) < 1e-5 and row_rel(Y, Yo) < 1e-5, (row_rel(X, Xo), row_rel(Y, Yo))
    assert loss == pytest.approx(lo, rel=1e-6)
    assert not X[m - 100:].any() and not Y[n - 100:].any()        # nobody there: b = 0 -> zero rows, like the reference
    # a sweep is a pure function of the other table: same bits when repeated
    engine.wrmf_sweep(0, 0.7, 10.0)
    Xa, _ = engine.get_factors()
    engine.wrmf_sweep(0, 0.7, 10.0)
    Xb, _ = engine.get_factors()
    assert np.array_equal(Xa, Xb)


def test_normal_equations_hold_at_scale(engine):
    """Size-independent property on a log the oracle would take minutes for: after a user sweep every sampled row
    satisfies (YtY + Y_u^T C_u Y_u + reg I) x = b to float32 storage precision."""
    m, n, k, reg = 60000, 20000, 64, 1.0
    log = synth.power_law_log(m, n, 1500000, seed=11)
    X0, Y0 = synth.init_factors(m, n, k, seed=12)
    engine.set_interactions(m, n, log.ev_indptr, log.ev_items, log.uq_indptr, log.uq_items)
    engine.set_factors(X0 * 10, Y0 * 10)
    engine.wrmf_sweep(0, reg, 10.0)
    X, Y = engine.get_factors()
    # this log is large enough to have hot tracks, which the SGD planner stores re-labelled in the device copy of the
    # events (mark_hot_kernel): the play counts must see through that
    cnt = wrmf_ref.pair_counts(log.ev_indptr, log.ev_items, log.uq_indptr, log.uq_items)
    assert np.array_equal(engine.wrmf_pair_counts()[0], cnt)
    assert not np.isnan(X).any()
    Y64 = Y.astype(np.float64)
    G = Y64.T @ Y64
    for u in np.r_[0:4, np.random.default_rng(0).integers(0, m, 40)]:
        a, b = log.uq_indptr[u], log.uq_indptr[u + 1]
        rows, c = Y64[log.uq_items[a:b]], 10.0 * cnt[a:b]
        A = G + (rows.T * c) @ rows + reg * np.eye(k)
        rhs = ((1 + c)[:, None] * rows).sum(0)
        x = np.linalg.solve(A, rhs)
        assert np.linalg.norm(X[u] - x) <= 2e-6 * max(np.linalg.norm(x), 1e-12), u


def test_wrmf_class_trains_and_ranks(tmp_path):
    """config/WRMF.conf-shaped run through Yue -> WRMF.execute(): factors follow the oracle from the same init stream,
    the measure list has the reference's layout."""
    import random
    from yue_b200.host.config import Config
    from yue_b200.host.driver import Yue
    from yue_b200.wrmf import WRMF
    log_path = tmp_path / "log.txt"
    synth.write_csv_log(str(log_path), 600, 2000, 15000, seed=8)
    vals = {"record": str(log_path), "record.setup": "-columns user:1,track:2,artist:3,time:0 -delim ,",
            "recommender": "WRMF", "evaluation.setup": "-target track -ap 0.2", "item.ranking": "-topN 5,10",
            "num.factors": "20", "num.max.iter": "3", "learnRate": "-init 0.02 -max 1",
            "reg.lambda": "-u 1 -i 0.1 -b 0.2 -s 0.2", "output.setup": "on -dir %s/" % (tmp_path / "res")}
    random.seed(2)
    np.random.seed(6)
    with redirect_stdout(io.StringIO()):
        y = Yue(Config(values=vals))
        model = WRMF(y.config, y.trainingData, y.testData)
        measure = model.execute()
    ev_indptr, ev_items, uq_indptr, uq_items = model.data.interaction_arrays()
    np.random.seed(6)
    X = np.random.rand(model.m, 20).astype(np.float32) / 10 * 10
    Y = np.random.rand(model.n, 20).astype(np.float32) / 10 * 10
    cnt = wrmf_ref.pair_counts(ev_indptr, ev_items, uq_indptr, uq_items)
    itp, itu, itc = wrmf_ref.transpose(model.m, model.n, uq_indptr, uq_items, cnt)
    for _ in range(3):
        loss = wrmf_ref.iteration(X, Y, uq_indptr, uq_items, cnt, itp, itu, itc, 1.0, gram="f64")
    assert row_rel(model.X, X) < 1e-4 and row_rel(model.Y, Y) < 1e-4
    assert model.loss == pytest.approx(loss, rel=1e-5)
    assert measure[0] == "Top 5\n" and measure[1].startswith("Precision:") and measure[6] == "Top 10\n"
    assert float(measure[8].split(":")[1]) > 0.02                 # Recall@10 well above chance (10/2000)


def test_row_ranges_compose_to_the_sweep_two_handles_emulate_two_ranks():
    """yue_wrmf_sweep_rows: two handles with the same log and tables each solve a range of rows (balanced by entries, a
    heavy row on each side of the cut), the solved rows are exchanged through the host in place of the broadcasts of
    WrmfShardedTrainer: tables bit-identical to the single-handle sweep, losses add up."""
    from yue_b200 import sharding
    from yue_b200.engine import Engine
    m, n, k = 5200, 6100, 40
    rng = np.random.default_rng(1)
    base = synth.power_law_log(m, n, 60000, seed=4)
    bu = np.repeat(np.arange(base.m), np.diff(base.ev_indptr))
    users = np.concatenate([bu, np.zeros(5000, np.int64), np.full(4800, m - 1), np.arange(4500)])
    items = np.concatenate([base.ev_items, np.arange(5000), np.arange(1000, 5800), np.full(4500, 7)])
    ev_indptr, ev_items, uq_indptr, uq_items = csr_from_events(m, n, users.astype(np.int64), items.astype(np.int64))
    X0, Y0 = synth.init_factors(m, n, k, seed=3)
    X0, Y0 = X0 * 10, Y0 * 10
    engs = [Engine(0), Engine(0), Engine(0)]
    try:
        for e in engs:
            e.set_interactions(m, n, ev_indptr, ev_items, uq_indptr, uq_items)
            e.set_factors(X0, Y0)
        itp = engs[0].wrmf_pair_counts()[1]
        ub, tb = sharding.shard_users_by_events(uq_indptr, 2), sharding.shard_users_by_events(itp, 2)
        assert 0 < ub[1] < m and 0 < tb[1] < n
        ref_loss = engs[2].wrmf_sweep(0, 0.7, 10.0, want_loss=True)
        engs[2].wrmf_sweep(1, 0.7, 10.0)
        Xr, Yr = engs[2].get_factors()
        parts = [engs[r].wrmf_sweep_rows(0, ub[r], ub[r + 1], 0.7, 10.0, want_loss=True) for r in range(2)]
        X = np.concatenate([engs[r].get_factors()[0][ub[r]:ub[r + 1]] for r in range(2)])
        assert np.array_equal(X, Xr) and sum(parts) == pytest.approx(ref_loss, rel=1e-12)
        for e in engs[:2]:
            e.set_factors(X, Y0)                                   # the exchange
        for r in range(2):
            engs[r].wrmf_sweep_rows(1, tb[r], tb[r + 1], 0.7, 10.0)
        Y = np.concatenate([engs[r].get_factors()[1][tb[r]:tb[r + 1]] for r in range(2)])
        assert np.array_equal(Y, Yr)
    finally:
        for e in engs:
            e.close()


@pytest.mark.parametrize("k", [12, 20, 64])
def test_every_solver_variant_agrees(monkeypatch, k):
    """Three ways through a half-sweep: the k x k LDL^T on a 16 x 16 grid of blocks (160 threads per row), for
    32 < k <= 64 the same on an 8 x 8 grid of 8 x 8 blocks (64 threads per row, YUE_WRMF_FAT=1), and --
    for rows with 1..32 entries -- the d x d Woodbury system on one warp (YUE_WRMF_LIGHT=1).  Same answer to float32
    storage precision, every one within 1e-5 of the oracle."""
    from yue_b200.engine import Engine
    log = synth.power_law_log(3000, 1500, 60000, seed=31)
    deg = np.diff(log.uq_indptr)
    assert (deg <= 32).sum() > 1000 and (deg > 32).sum() > 50
    X0, Y0 = synth.init_factors(log.m, log.n, k, seed=32)
    X0, Y0 = X0 * 10, Y0 * 10
    res = {}
    for light, fat in (("0", "1"), ("0", "0"), ("1", "0")):
        monkeypatch.setenv("YUE_WRMF_LIGHT", light)
        monkeypatch.setenv("YUE_WRMF_FAT", fat)
        e = Engine(0)
        try:
            e.set_interactions(log.m, log.n, log.ev_indptr, log.ev_items, log.uq_indptr, log.uq_items)
            e.set_factors(X0, Y0)
            loss = e.wrmf_sweep(0, 0.3, 10.0, want_loss=True)
            e.wrmf_sweep(1, 0.3, 10.0)
            res[light + fat] = e.get_factors() + (loss,)
        finally:
            e.close()
    cnt = wrmf_ref.pair_counts(log.ev_indptr, log.ev_items, log.uq_indptr, log.uq_items)
    itp, itu, itc = wrmf_ref.transpose(log.m, log.n, log.uq_indptr, log.uq_items, cnt)
    Xo, Yo = X0.copy(), Y0.copy()
    lo = wrmf_ref.iteration(Xo, Yo, log.uq_indptr, log.uq_items, cnt, itp, itu, itc, 0.3, gram="f64")
    for key, (X, Y, loss) in res.items():
        assert row_rel(X, res["01"][0]) < 2e-6 and row_rel(Y, res["01"][1]) < 2e-6, key
        assert row_rel(X, Xo) < 1e-5 and row_rel(Y, Yo) < 1e-5, key
        assert loss == pytest.approx(lo, rel=1e-6) and loss == pytest.approx(res["01"][2], rel=1e-12), key


def test_edge_cases(engine):
    """Empty log, tiny k (padded leading dimension), alpha = 0, empty row range, and the refusals."""
    from yue_b200.engine import Engine, YueError
    from yue_b200 import _lib
    # a log without events: every row is 0 (b = 0), loss 0
    m, n = 7, 5
    z = np.zeros(m + 1, np.int64)
    engine.set_interactions(m, n, z, np.zeros(0, np.int32), z, np.zeros(0, np.int32))
    X0, Y0 = synth.init_factors(m, n, 3, seed=1)
    engine.set_factors(X0, Y0)
    assert engine.wrmf_sweep(0, 0.5, 10.0, want_loss=True) == 0.0
    engine.wrmf_sweep(1, 0.5, 10.0)
    X, Y = engine.get_factors()
    assert not X.any() and not Y.any()
    # k = 1 and k = 3 (ld = 4), alpha = 0 (x = (G + reg I)^-1 sum of the played rows), repeat plays
    log = synth.power_law_log(40, 30, 600, seed=2)
    cnt = wrmf_ref.pair_counts(log.ev_indptr, log.ev_items, log.uq_indptr, log.uq_items)
    itp, itu, itc = wrmf_ref.transpose(log.m, log.n, log.uq_indptr, log.uq_items, cnt)
    for k, alpha in ((1, 10.0), (3, 10.0), (3, 0.0), (17, 2.5)):
        X0, Y0 = synth.init_factors(log.m, log.n, k, seed=k)
        X0, Y0 = X0 * 10, Y0 * 10
        engine.set_interactions(log.m, log.n, log.ev_indptr, log.ev_items, log.uq_indptr, log.uq_items)
        engine.set_factors(X0, Y0)
        loss = engine.wrmf_sweep(0, 0.25, alpha, want_loss=True)
        engine.wrmf_sweep_rows(1, 3, 3, 0.25, alpha)                 # empty range: nothing happens
        Xh, Yh = engine.get_factors()
        assert np.array_equal(Yh, Y0)
        engine.wrmf_sweep(1, 0.25, alpha)
        X, Y = engine.get_factors()
        Xo, Yo = X0.copy(), Y0.copy()
        lo = wrmf_ref.half_sweep(Xo, Yo, log.uq_indptr, log.uq_items, cnt, 0.25, "f64", True, alpha)
        wrmf_ref.half_sweep(Yo, Xo, itp, itu, itc, 0.25, "f64", False, alpha)
        assert row_rel(X, Xo) < 1e-5 and row_rel(Y, Yo) < 1e-5, (k, alpha)
        assert loss == pytest.approx(lo, rel=1e-6)
    # refusals
    with pytest.raises(YueError) as ei:
        engine.wrmf_sweep(2, 1.0)
    assert ei.value.code == _lib.E_ARG
    with pytest.raises(YueError) as ei:
        engine.wrmf_sweep_rows(0, 5, log.m + 1, 1.0)
    assert ei.value.code == _lib.E_ARG
    with pytest.raises(YueError) as ei:
        engine.wrmf_sweep(0, -1.0)
    assert ei.value.code == _lib.E_ARG
    X0, Y0 = synth.init_factors(log.m, log.n, 130, seed=1)
    engine.set_factors(X0, Y0)
    with pytest.raises(YueError) as ei:
        engine.wrmf_sweep(0, 1.0)
    assert ei.value.code == _lib.E_UNSUPPORTED
    e2 = Engine(0)                                                   # a user shard cannot run the track sweep
    try:
        e2.set_interactions(log.m, log.n, log.ev_indptr, log.ev_items, log.uq_indptr, log.uq_items, user_begin=10, event_base=100)
        e2.set_factors(*synth.init_factors(log.m, log.n, 8, seed=1))
        with pytest.raises(YueError) as ei:
            e2.wrmf_sweep(0, 1.0)
        assert ei.value.code == _lib.E_UNSUPPORTED
    finally:
        e2.close()
