"""K8 (CUNE) on hardware: print -- not assert -- how far each mode lands from the oracle loop.
Run on the GPU box:  python tools/cune_diag.py  (writes gpurun_out/cune_diag.log when that directory exists)."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from oracle import cune_ref, philox, record_ref  # noqa: E402
from yue_b200 import synth  # noqa: E402
from yue_b200.cune import implicit_positive_lists  # noqa: E402
from yue_b200.engine import MODE_HOGWILD, MODE_SERIAL, Engine  # noqa: E402


def rel(a, b):
    d = np.linalg.norm(a.astype(np.float64) - b, axis=1) / np.maximum(np.linalg.norm(b.astype(np.float64), axis=1), 1e-12)
    return float(d.max()), int(d.argmax())


def scale(say):
    """Serial order against the Hogwild schedule the C-ABI picks, at a size where both are meaningful: the -log part of
    the loss per epoch, and Recall@10 / NDCG@10 of the trained tables (K3 exact ranking + K6 metrics)."""
    import time
    from yue_b200.engine import RANK_EXACT
    users, tracks, plays, d, epochs = 20000, 5000, 500000, 64, 3
    log = synth.power_law_log(users, tracks, plays, seed=41)
    m, n = log.m, log.n
    top = {u: [f for f in ((u * 7 + 3) % m, (u * 11 + 5) % m) if f != u] for u in range(m) if u % 5}
    t0 = time.time()
    ip_indptr, ip_items = implicit_positive_lists(m, log.uq_indptr, log.uq_items, top)
    P0, Q0 = synth.init_factors(m, n, d, seed=6)
    say("scale log: %d users x %d tracks, %d train events, %d implicit positives (%.1f s to build), d=%d"
        % (m, n, log.train_size, len(ip_items), time.time() - t0, d))
    os.environ["YUE_CUNE_EVENT_LOSS"] = "1"
    tu = log.test_users()
    base = None
    for name, mode, env in (("serial", MODE_SERIAL, {}), ("hogwild auto", MODE_HOGWILD, {}), ("hogwild auto (run 2)", MODE_HOGWILD, {}),
                            ("hogwild 4096 ev/warp", MODE_HOGWILD, {"YUE_SGD_MIN_EVENTS_PER_WARP": "4096"}),
                            ("hogwild 1024 ev/warp", MODE_HOGWILD, {"YUE_SGD_MIN_EVENTS_PER_WARP": "1024"}),
                            ("hogwild 2368 warps", MODE_HOGWILD, {"YUE_CUNE_WARPS": "2368"})):
        os.environ.update(env)
        eng = Engine(0)
        for k in env:
            os.environ.pop(k, None)
        os.environ.update({k: v for k, v in env.items() if k == "YUE_CUNE_WARPS"})
        eng.set_interactions(m, n, log.ev_indptr, log.ev_items, log.uq_indptr, log.uq_items)
        eng.set_factors(P0.copy(), Q0.copy())
        eng.cune_set_implicit(ip_indptr, ip_items)
        losses, t0 = [], time.time()
        for ep in range(epochs):
            losses.append(eng.cune_epoch(0.02, 0.01, 0.01, 2.0, 17, ep, mode))
        dt = time.time() - t0
        os.environ.pop("YUE_CUNE_WARPS", None)
        eng.rank_topn(tu, 10, RANK_EXACT)
        eng.set_test_set(log.test_indptr, log.test_items)
        sums, _ = eng.rank_metrics([10])
        r, nd = sums[0, 1] / len(tu), sums[0, 3] / len(tu)
        if base is None:
            base = (r, nd, losses)
        say("scale %-22s %6.2f s  loss %s  rel-to-serial %s | recall@10 %.4f (%+.4f) ndcg@10 %.4f (%+.4f)"
            % (name, dt, " ".join("%.1f" % x for x in losses), " ".join("%+.4f" % (x / y - 1) for x, y in zip(losses, base[2])),
               r, r - base[0], nd, nd - base[1]))
        eng.close()
    os.environ.pop("YUE_CUNE_EVENT_LOSS", None)


def main():
    out = []
    say = lambda s: (out.append(s), print(s, flush=True))
    eng = Engine(0)
    g = np.load(os.path.join(ROOT, "tests", "golden", "cune_small.npz"))
    m, n = g["P0"].shape[0], g["Q0"].shape[0]
    lr, regU, regI, s, seed = float(g["lr"]), float(g["regU"]), float(g["regI"]), float(g["s"]), int(g["seed"])
    say("golden log: m=%d n=%d T=%d k=%d max events/user=%d" % (m, n, len(g["ev_items"]), g["P0"].shape[1], int(np.diff(g["ev_indptr"]).max())))
    for mode, name, ru, ri in ((MODE_SERIAL, "serial", regU, regI), (MODE_HOGWILD, "hogwild", 0.0, 0.0), (MODE_HOGWILD, "hogwild+reg", regU, regI)):
        for chunk in ((None,) if mode == MODE_SERIAL else (None, "32")):
            if chunk:
                os.environ["YUE_CUNE_CHUNK"] = chunk
            else:
                os.environ.pop("YUE_CUNE_CHUNK", None)
            eng.set_interactions(m, n, g["ev_indptr"], g["ev_items"], g["uq_indptr"], g["uq_items"])
            eng.set_factors(g["P0"].copy(), g["Q0"].copy())
            eng.cune_set_implicit(g["ip_indptr"], g["ip_items"])
            try:
                loss = eng.cune_epoch(lr, ru, ri, s, seed, 0, mode)
            except Exception as exc:  # noqa: BLE001
                say("golden %s chunk=%s: EXCEPTION %r" % (name, chunk, exc))
                continue
            P, Q = eng.get_factors()
            Pr, Qr = g["P0"].copy(), g["Q0"].copy()
            ref = cune_ref.epoch(Pr, Qr, g["ev_indptr"], g["ev_items"], g["ip_indptr"], g["ip_items"], g["kpos"][0], g["neg"][0], lr, ru, ri, s)
            say("golden %-12s chunk=%-4s rel P %.3e@%d Q %.3e@%d | abs P %.3e Q %.3e | loss %.8g ref %.8g | finite %s"
                % (name, chunk, *rel(P, Pr), *rel(Q, Qr), np.abs(P - Pr).max(), np.abs(Q - Qr).max(), loss, float(ref),
                   bool(np.isfinite(P).all() and np.isfinite(Q).all())))
    os.environ.pop("YUE_CUNE_CHUNK", None)
    for d in (12, 16, 20, 32, 64, 128, 200):
        log = synth.power_law_log(80, 150, 1500, seed=11, test_ratio=0.0)
        m, n = log.m, log.n
        top = {u: [f for f in ((u * 7 + 3) % m, (u * 11 + 5) % m) if f != u] for u in range(m) if u % 5}
        ip_indptr, ip_items = implicit_positive_lists(m, log.uq_indptr, log.uq_items, top)
        rng = np.random.default_rng(d)
        P0 = rng.uniform(0, 0.1, (m, d)).astype(np.float32)
        Q0 = rng.uniform(0, 0.1, (n, d)).astype(np.float32)
        ev_user = record_ref.ev_users(log.ev_indptr)
        kp = [cune_ref.sample_implicit(5, 0, nn, ev_user, ip_indptr) for nn in range(3)]
        neg = [philox.sample_negatives(5, 0, ev_user, n, log.uq_indptr, log.uq_items, slot=nn) for nn in range(3)]
        for mode, name, reg, chunk in ((MODE_SERIAL, "serial", 0.01, None), (MODE_SERIAL, "serial-noreg", 0.0, None),
                                       (MODE_HOGWILD, "hogwild", 0.0, None), (MODE_HOGWILD, "hogwild", 0.0, "32")):
            if chunk:
                os.environ["YUE_CUNE_CHUNK"] = chunk
            else:
                os.environ.pop("YUE_CUNE_CHUNK", None)
            eng.set_interactions(m, n, log.ev_indptr, log.ev_items, log.uq_indptr, log.uq_items)
            eng.set_factors(P0.copy(), Q0.copy())
            eng.cune_set_implicit(ip_indptr, ip_items)
            try:
                loss = eng.cune_epoch(0.02, reg, reg, 2.0, 5, 0, mode)
            except Exception as exc:  # noqa: BLE001
                say("d=%d %s chunk=%s: EXCEPTION %r" % (d, name, chunk, exc))
                continue
            P, Q = eng.get_factors()
            Pr, Qr = P0.copy(), Q0.copy()
            ref = cune_ref.epoch(Pr, Qr, log.ev_indptr, log.ev_items, ip_indptr, ip_items, kp, neg, 0.02, reg, reg, 2.0)
            say("d=%-3d %-12s chunk=%-4s rel P %.3e@%d Q %.3e@%d | abs P %.3e Q %.3e | loss %.8g ref %.8g"
                % (d, name, chunk, *rel(P, Pr), *rel(Q, Qr), np.abs(P - Pr).max(), np.abs(Q - Qr).max(), loss, float(ref)))
    os.environ.pop("YUE_CUNE_CHUNK", None)
    eng.close()
    if "--scale" in sys.argv:
        scale(say)
    dst = os.path.join(ROOT, "gpurun_out")
    if os.path.isdir(dst):
        open(os.path.join(dst, "cune_diag.log"), "w").write("\n".join(out) + "\n")


if __name__ == "__main__":
    main()
