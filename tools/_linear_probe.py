import os, sys, numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from yue_b200 import synth
from yue_b200.engine import Engine, MODE_SERIAL, MODE_HOGWILD
log = synth.power_law_log(1500, 400, 150000, seed=4)
P, Q = synth.init_factors(log.m, log.n, 64, seed=6)
def run(mode, env):
    os.environ.update({k: str(v) for k, v in env.items()})
    e = Engine(0)
    e.set_interactions(log.m, log.n, log.ev_indptr, log.ev_items, log.uq_indptr, log.uq_items)
    e.set_factors(P, Q)
    loss = e.bpr_epoch(1e-5, 0.0, 0.0, 3, 0, mode)
    r = e.get_factors(); e.close(); return loss, r
ls, (Ps, Qs) = run(MODE_SERIAL, {})
for name, env in [("direct 1 warp", dict(YUE_SGD_MIN_EVENTS_PER_WARP=10**9, YUE_SGD_HOT_MAX=0)),
                  ("direct 7 warps", dict(YUE_SGD_MIN_EVENTS_PER_WARP=16384, YUE_SGD_HOT_MAX=0)),
                  ("direct 58 warps", dict(YUE_SGD_MIN_EVENTS_PER_WARP=2048, YUE_SGD_HOT_MAX=0)),
                  ("hot 7 warps", dict(YUE_SGD_MIN_EVENTS_PER_WARP=16384, YUE_SGD_HOT_MAX=64, YUE_SGD_HOT_MIN_COUNT=1, YUE_SGD_HOT_FLUSH=8)),
                  ("hot 58 warps", dict(YUE_SGD_MIN_EVENTS_PER_WARP=2048, YUE_SGD_HOT_MAX=64, YUE_SGD_HOT_MIN_COUNT=1, YUE_SGD_HOT_FLUSH=8))]:
    l, (Pg, Qg) = run(MODE_HOGWILD, env)
    mq = np.linalg.norm(Qg - Q) / np.linalg.norm(Qs - Q); mp = np.linalg.norm(Pg - P) / np.linalg.norm(Ps - P)
    eq = np.abs((Qg - Q) - (Qs - Q)).max() / np.abs(Qs - Q).max(); ep = np.abs((Pg - P) - (Ps - P)).max() / np.abs(Ps - P).max()
    worst_u = int(np.abs((Pg - P) - (Ps - P)).max(axis=1).argmax())
    print("%-16s loss %.3f (serial %.3f) movement ratio Q %.4f P %.4f  max err Q %.3e P %.3e (user %d, deg %d)" % (name, l, ls, mq, mp, eq, ep, worst_u, log.ev_indptr[worst_u+1]-log.ev_indptr[worst_u]))
