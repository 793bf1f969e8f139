import os, sys, numpy as np
sys.path.insert(0, '/root/repo')
from yue_b200 import synth
from yue_b200.engine import Engine, MODE_SERIAL, MODE_HOGWILD, MODE_HOGWILD_STORE
log = synth.power_law_log(2000, 500, 100000, seed=7)
P, Q = synth.init_factors(log.m, log.n, 64, seed=1)
def run(mode, env):
    os.environ.update({k: str(v) for k, v in env.items()})
    e = Engine(0)
    e.set_interactions(log.m, log.n, log.ev_indptr, log.ev_items, log.uq_indptr, log.uq_items)
    e.set_factors(P, Q)
    for ep in range(3):
        loss = e.bpr_epoch(0.02, 0.01, 0.01, 5, ep, mode)
    r = e.get_factors(); e.close(); return loss, r
ls, (Ps, Qs) = run(MODE_SERIAL, {})
for name, mode, env in [("atomic 1 warp", MODE_HOGWILD, dict(YUE_SGD_MIN_EVENTS_PER_WARP=10**9, YUE_SGD_HOT_MAX=0)),
                        ("store 1 warp", MODE_HOGWILD_STORE, dict(YUE_SGD_MIN_EVENTS_PER_WARP=10**9, YUE_SGD_HOT_MAX=0)),
                        ("atomic 4 warps", MODE_HOGWILD, dict(YUE_SGD_MIN_EVENTS_PER_WARP=20000, YUE_SGD_HOT_MAX=0)),
                        ("atomic 40 warps", MODE_HOGWILD, dict(YUE_SGD_MIN_EVENTS_PER_WARP=2000, YUE_SGD_HOT_MAX=0))]:
    l, (Pg, Qg) = run(mode, env)
    dq = np.abs(Qg - Qs).max() / np.abs(Qs).max(); dp = np.abs(Pg - Ps).max() / np.abs(Ps).max()
    mv = np.linalg.norm(Qg - Q) / np.linalg.norm(Qs - Q)
    print("%-16s loss %.2f (serial %.2f)  max|dQ|/max|Q| %.3e  max|dP|/max|P| %.3e  movement ratio %.4f  |Q0| %.4f vs %.4f" % (name, l, ls, dq, dp, mv, np.linalg.norm(Qg[0]), np.linalg.norm(Qs[0])))
