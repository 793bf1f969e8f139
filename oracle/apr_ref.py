"""Oracle (test infrastructure): APR -- adversarial BPR -- as a per-triplet SGD step.

The reference's APR (recommender/advanced/APR.py, He et al. 2018) is a TF-1 graph that cannot be
imported here (base/DeepRecommender has no .py suffix) and is trained with Adam on
mini-batches of 512 events x 3 negatives.  What it defines (APR.py:25-76) and what is restated:

* y     = U_u . (V_i - V_j)                                     (_create_inference, 38-41)
* y_adv = (U_u + D_u) . ((V_i + D_i) - (V_j + D_j))             (_create_adv_inference, 43-49)
* D     = eps * l2_normalize(d L_adv / d D) per row, held constant in the step  (_create_adversarial, 51-60)
* loss  = softplus(-y) + regA * softplus(-y_adv)                (_create_loss, 62-72)

north_star asks for "the triplet SGD kernel with fused perturbation": the perturbation is taken
per TRIPLET at D = 0 (the reference aggregates the gradient per row over the batch first), which
gives it in closed form.  With d = V_i - V_j, P = U_u, hats = unit vectors:
    D_u = -eps d^,  D_i = -eps P^,  D_j = +eps P^
    y_adv = y - 2 eps |P| - eps |d| + 2 eps^2 y / (|P| |d|)
and the SGD step on the loss above with D constant (all gradients from the OLD rows):
    s0 = sigmoid(-y), s1 = sigmoid(-y_adv), a = lr (s0 + regA s1), b = lr regA s1 eps
    P  += a d - 2 b P^      V_i += a P - b d^      V_j -= a P - b d^
followed by the multiplicative L2 shrink of BPR.py:55-57.  The optimiser (SGD instead of Adam), the
epoch structure and the per-triplet perturbation are this build's (north_star's).  The FORMULAS are pinned by the
reference's own graph: APR.py and base/DeepRecommender, unmodified, are evaluated over oracle/tf1_shim.py (their
TensorFlow-1 calls on torch autograd) on single triplets -- where the batch form and the per-triplet form coincide --
and the perturbation, y_adv, loss_adv and its gradients with the perturbation held constant agree with this file to
1e-14 (oracle/make_golden_apr.py -> tests/golden/apr_graph.npz, tests/test_apr.py); a numerical-gradient check besides.
"""
import math

import numpy as np


def softplus(z):
    return max(z, 0.0) + math.log1p(math.exp(-abs(z)))


def adv_score(x, n_p, n_d, eps):
    y = x - 2.0 * eps * n_p - eps * n_d
    if n_p > 0 and n_d > 0:
        y += 2.0 * eps * eps * x / (n_p * n_d)
    return y


def apr_epoch(P, Q, ev_user, ev_item, ev_neg, lr, regU, regI, eps, regA):
    """One pass over the triplet stream, in place (float32 or float64 tables).  Returns the loss."""
    loss = 0.0
    dt = P.dtype.type
    for e in range(len(ev_user)):
        u, i, j = int(ev_user[e]), int(ev_item[e]), int(ev_neg[e])
        p, d = P[u].copy(), Q[i] - Q[j]
        x = float(p.dot(d))
        n_p, n_d = float(np.sqrt(p.dot(p))), float(np.sqrt(d.dot(d)))
        xa = adv_score(x, n_p, n_d, eps)
        s0 = 1.0 / (1.0 + math.exp(x)) if x < 700 else 0.0
        s1 = 1.0 / (1.0 + math.exp(xa)) if xa < 700 else 0.0
        a, b = lr * (s0 + regA * s1), lr * regA * s1 * eps
        ph = p / n_p if n_p > 0 else p * 0
        dh = d / n_d if n_d > 0 else d * 0
        P[u] = p + dt(a) * d - dt(2.0 * b) * ph
        step = dt(a) * p - dt(b) * dh
        Q[i] += step
        Q[j] -= step
        P[u] -= dt(lr * regU) * P[u]
        Q[i] -= dt(lr * regI) * Q[i]
        Q[j] -= dt(lr * regI) * Q[j]
        loss += softplus(-x) + regA * softplus(-xa)
    return loss


def triplet_loss_with_fixed_delta(p, qi, qj, du, di, dj, regA):
    """softplus(-y) + regA softplus(-y_adv) with the perturbation given (for the gradient check)."""
    y = p.dot(qi - qj)
    ya = (p + du).dot((qi + di) - (qj + dj))
    return softplus(-y) + regA * softplus(-ya)
