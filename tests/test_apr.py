"""APR (adversarial BPR): the oracle's closed forms against the REFERENCE'S OWN GRAPH (recommender/advanced/APR.py and
base/DeepRecommender, unmodified, evaluated over oracle/tf1_shim.py on single triplets: tests/golden/apr_graph.npz, written
by oracle/make_golden_apr.py) and against first principles (CPU), and the fused GPU kernel against the oracle.  The
optimiser (per-triplet SGD where the reference runs Adam on mini-batches) and the per-triplet perturbation (the reference
aggregates its gradient per row over the batch) are north_star's and stay this build's."""
import numpy as np
import pytest

from oracle import apr_ref


def test_closed_form_adversarial_score_and_step_match_first_principles():
    rng = np.random.default_rng(0)
    d, eps, regA, lr = 16, 0.5, 2.0, 1e-3
    for _ in range(5):
        p, qi, qj = (rng.normal(size=d) for _ in range(3))
        dv = qi - qj
        du = -eps * dv / np.linalg.norm(dv)
        di, dj = -eps * p / np.linalg.norm(p), eps * p / np.linalg.norm(p)
        # the perturbation is eps * normalised gradient of the adversarial loss at delta = 0
        f = lambda a, b, c: apr_ref.triplet_loss_with_fixed_delta(p, qi, qj, a, b, c, regA)
        h = 1e-6
        z = np.zeros(d)
        g_u = np.array([(f(z + h * np.eye(d)[k], z, z) - f(z - h * np.eye(d)[k], z, z)) / (2 * h) for k in range(d)])
        g_i = np.array([(f(z, z + h * np.eye(d)[k], z) - f(z, z - h * np.eye(d)[k], z)) / (2 * h) for k in range(d)])
        assert np.allclose(eps * g_u / np.linalg.norm(g_u), du, atol=1e-5)
        assert np.allclose(eps * g_i / np.linalg.norm(g_i), di, atol=1e-5)
        # closed-form adversarial score
        ya = (p + du).dot((qi + di) - (qj + dj))
        assert apr_ref.adv_score(p.dot(dv), np.linalg.norm(p), np.linalg.norm(dv), eps) == pytest.approx(ya, rel=1e-12)
        # one oracle step (no shrink) equals -lr * numerical gradient with delta held constant
        P, Q = p[None, :].copy(), np.stack([qi, qj])
        apr_ref.apr_epoch(P, Q, [0], [0], [1], lr, 0.0, 0.0, eps, regA)
        L = lambda pp, a, b: apr_ref.triplet_loss_with_fixed_delta(pp, a, b, du, di, dj, regA)
        gp = np.array([(L(p + h * np.eye(d)[k], qi, qj) - L(p - h * np.eye(d)[k], qi, qj)) / (2 * h) for k in range(d)])
        gi = np.array([(L(p, qi + h * np.eye(d)[k], qj) - L(p, qi - h * np.eye(d)[k], qj)) / (2 * h) for k in range(d)])
        gj = np.array([(L(p, qi, qj + h * np.eye(d)[k]) - L(p, qi, qj - h * np.eye(d)[k])) / (2 * h) for k in range(d)])
        assert np.allclose(P[0] - p, -lr * gp, atol=1e-8)
        assert np.allclose(Q[0] - qi, -lr * gi, atol=1e-8) and np.allclose(Q[1] - qj, -lr * gj, atol=1e-8)


def test_oracle_step_is_the_reference_graph_on_a_single_triplet(golden_dir):
    """For a batch that is ONE triplet the reference's graph and the per-triplet form coincide.  From the reference's own
    text: the perturbation `assign(l2_normalize(grad) * eps)` (APR.py:51-60) is the closed form -eps d^ / -eps P^ / +eps P^;
    `loss_adv` (62-72) is softplus(-y) + regA softplus(-y_adv) with adv_score()'s y_adv; and the gradients of `loss_adv`
    with respect to U and V with the perturbation held constant (what `minimize(self.loss_adv)` differentiates) are the
    step of apr_epoch(): one call with lr and no shrink moves the rows by -lr times them."""
    import os
    g = np.load(os.path.join(golden_dir, "apr_graph.npz"))
    eps, regA, lr = float(g["eps"]), float(g["regA"]), 0.01
    assert len(g["u"]) == 24
    for t in range(len(g["u"])):
        p, qi, qj = g["p"][t], g["qi"][t], g["qj"][t]
        d = qi - qj
        n_p, n_d = np.linalg.norm(p), np.linalg.norm(d)
        assert np.allclose(g["du"][t], -eps * d / n_d, atol=1e-13)
        assert np.allclose(g["di"][t], -eps * p / n_p, atol=1e-13) and np.allclose(g["dj"][t], eps * p / n_p, atol=1e-13)
        assert g["y"][t] == pytest.approx(p.dot(d), rel=1e-12)
        assert apr_ref.adv_score(float(p.dot(d)), n_p, n_d, eps) == pytest.approx(g["ya"][t], rel=1e-12, abs=1e-13)
        P, Q = p[None, :].copy(), np.stack([qi, qj])
        loss = apr_ref.apr_epoch(P, Q, [0], [0], [1], lr, 0.0, 0.0, eps, regA)
        assert loss == pytest.approx(g["loss_adv"][t], rel=1e-12)
        assert np.allclose(P[0] - p, -lr * g["gp"][t], atol=1e-14)
        assert np.allclose(Q[0] - qi, -lr * g["gi"][t], atol=1e-14) and np.allclose(Q[1] - qj, -lr * g["gj"][t], atol=1e-14)


def test_apr_reduces_to_simultaneous_bpr_when_adversary_is_off():
    rng = np.random.default_rng(1)
    P, Q = rng.normal(size=(3, 8)), rng.normal(size=(5, 8))
    P2, Q2 = P.copy(), Q.copy()
    apr_ref.apr_epoch(P, Q, [0, 1], [2, 3], [4, 0], 0.05, 0.0, 0.0, 0.0, 0.0)
    for u, i, j in ((0, 2, 4), (1, 3, 0)):
        p, dv = P2[u].copy(), Q2[i] - Q2[j]
        g = 0.05 / (1.0 + np.exp(p.dot(dv)))
        P2[u] += g * dv; Q2[i] += g * p; Q2[j] -= g * p
    assert np.allclose(P, P2) and np.allclose(Q, Q2)


# ---- GPU: the fused kernel (K2a) against the oracle --------------------------------------------
@pytest.mark.gpu
@pytest.mark.parametrize("d", [10, 64, 128])
def test_apr_apply_serial_matches_oracle(engine, d):
    from yue_b200 import synth
    from yue_b200.engine import MODE_SERIAL
    rng = np.random.default_rng(d)
    m, n, T = 40, 70, 800
    log = synth.power_law_log(m, n, 1500, seed=3)
    engine.set_interactions(log.m, log.n, log.ev_indptr, log.ev_items, log.uq_indptr, log.uq_items)
    P = rng.normal(0, 0.3, (m, d)).astype(np.float32)       # norms well away from 0: the step divides by |P|, |d|
    Q = rng.normal(0, 0.3, (n, d)).astype(np.float32)
    engine.set_factors(P, Q)
    u = np.sort(rng.integers(0, m, T)).astype(np.int32)
    i = rng.integers(0, n, T).astype(np.int32)
    j = ((i + 1 + rng.integers(0, n - 1, T)) % n).astype(np.int32)
    loss = engine.apr_apply(u, i, j, 0.01, 0.002, 0.01, 0.5, 2.0, MODE_SERIAL)
    Pr, Qr = P.copy(), Q.copy()
    ref = apr_ref.apr_epoch(Pr, Qr, u, i, j, 0.01, 0.002, 0.01, 0.5, 2.0)
    Pg, Qg = engine.get_factors()
    P64, Q64 = P.astype(np.float64), Q.astype(np.float64)
    apr_ref.apr_epoch(P64, Q64, u, i, j, 0.01, 0.002, 0.01, 0.5, 2.0)
    err = lambda a, b: float(np.max(np.abs(a - b) / (np.abs(b) + 1e-2)))
    # the float64 run is the yardstick: the kernel may be no further from it than a few times the
    # float32 oracle is (both are fp32 chains of 800 steps with two normalisations each)
    tol = max(3 * max(err(Pr, P64), err(Qr, Q64)), 2e-5)
    assert err(Pg, P64) < tol and err(Qg, Q64) < tol and tol < 1e-3
    assert loss == pytest.approx(ref, rel=1e-5)


@pytest.mark.gpu
def test_apr_hogwild_conflict_free_and_epoch_slots(engine):
    from oracle import philox, record_ref
    from yue_b200 import synth
    from yue_b200.engine import MODE_HOGWILD, MODE_SERIAL
    m, n, d = 3000, 6200, 64
    log = synth.power_law_log(m, n, 40000, seed=4)
    engine.set_interactions(log.m, log.n, log.ev_indptr, log.ev_items, log.uq_indptr, log.uq_items)
    P, Q = synth.init_factors(m, n, d, seed=1)
    u = np.arange(m, dtype=np.int32)
    i, j = (2 * np.arange(m)).astype(np.int32), (2 * np.arange(m) + 1).astype(np.int32)
    Pr, Qr = P.copy(), Q.copy()
    ref = apr_ref.apr_epoch(Pr, Qr, u, i, j, 0.003, 0.002, 0.01, 0.5, 2.0)
    engine.set_factors(P, Q)
    loss = engine.apr_apply(u, i, j, 0.003, 0.002, 0.01, 0.5, 2.0, MODE_HOGWILD)
    Pg, Qg = engine.get_factors()
    assert np.allclose(Pg, Pr, rtol=1e-5, atol=1e-7) and np.allclose(Qg, Qr, rtol=1e-5, atol=1e-7)
    assert loss == pytest.approx(ref, rel=1e-5)
    # an APR epoch with slot s uses the slot-s negatives of the shared sampler
    ev_user = record_ref.ev_users(log.ev_indptr)
    for slot in (0, 2):
        engine.set_factors(P, Q)
        le = engine.apr_epoch(0.003, 0.002, 0.01, 0.5, 2.0, 9, 1, slot, MODE_SERIAL)
        neg = philox.sample_negatives(9, 1, ev_user, log.n, log.uq_indptr, log.uq_items, slot=slot)
        assert np.array_equal(engine.sample_negatives(9, 1, slot), neg)
        Pr, Qr = P.copy(), Q.copy()
        ref = apr_ref.apr_epoch(Pr, Qr, ev_user, log.ev_items, neg, 0.003, 0.002, 0.01, 0.5, 2.0)
        Pg, Qg = engine.get_factors()
        assert np.allclose(Pg, Pr, rtol=1e-4, atol=1e-6) and np.allclose(Qg, Qr, rtol=1e-4, atol=1e-6)
        assert le == pytest.approx(ref, rel=1e-5)


@pytest.mark.gpu
def test_apr_class_trains_and_ranks(tmp_path):
    import io
    from contextlib import redirect_stdout
    from yue_b200 import synth
    from yue_b200.apr import APR
    from yue_b200.host.config import Config
    from yue_b200.host.driver import Yue
    log_path = tmp_path / "log.txt"
    synth.write_csv_log(str(log_path), 1500, 800, 60000, seed=3)
    vals = {"record": str(log_path), "record.setup": "-columns user:1,track:2,artist:3,time:0 -delim ,",
            "recommender": "APR", "evaluation.setup": "-target track -ap 0.2", "item.ranking": "-topN 5,10",
            "num.factors": "64", "num.max.iter": "6", "batch_size": "512", "APR": "-regA 2 -eps 0.5 -advEpoch 4",
            "learnRate": "-init 0.02 -max 1", "reg.lambda": "-u 0.002 -i 0.01 -b 0.2 -s 0.2",
            "output.setup": "off -dir %s/" % tmp_path, "yue.seed": "5"}
    with redirect_stdout(io.StringIO()):
        measure = Yue(Config(values=vals)).execute()
    assert measure[6] == "Top 10\n" and float(measure[8].split(":")[1]) > 0.05       # Recall@10 of a trained model


@pytest.mark.gpu
@pytest.mark.parametrize("d", [32, 64, 128])
def test_apr_blocked_kernel_recurrence_all_widths(engine, d):
    """The blocked APR kernel (16 dots per block of 4 triplets, |P|^2 and the scores from the scalar
    recurrence) against the per-triplet oracle: several triplets per user on disjoint rows, so blocks
    of 1-4 triplets, the |P| recurrence over a block and segment boundaries are all exercised."""
    from yue_b200 import synth
    from yue_b200.engine import MODE_HOGWILD
    m = 200
    per = 1 + np.arange(m) % 37
    u = np.repeat(np.arange(m, dtype=np.int32), per)
    T = len(u)
    n = 2 * T + 10
    log = synth.power_law_log(m, n, 6000, seed=3)
    engine.set_interactions(log.m, log.n, log.ev_indptr, log.ev_items, log.uq_indptr, log.uq_items)
    P, Q = synth.init_factors(m, n, d, seed=1)
    i, j = (2 * np.arange(T)).astype(np.int32), (2 * np.arange(T) + 1).astype(np.int32)
    Pr, Qr = P.astype(np.float64), Q.astype(np.float64)
    ref = apr_ref.apr_epoch(Pr, Qr, u, i, j, 0.01, 0.002, 0.01, 0.5, 2.0)
    engine.set_factors(P, Q)
    loss = engine.apr_apply(u, i, j, 0.01, 0.002, 0.01, 0.5, 2.0, MODE_HOGWILD)
    Pg, Qg = engine.get_factors()
    assert np.allclose(Pg, Pr, rtol=2e-5, atol=2e-7) and np.allclose(Qg, Qr, rtol=2e-5, atol=2e-7)
    assert loss == pytest.approx(ref, rel=1e-5)
