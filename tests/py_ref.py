"""Second, independent restatement of CollapsedEMOptimizer::optimize in pure Python
(small cases only) -- used to cross-check oracle/em_oracle.c.  Written from the
reference source (src/inference/CollapsedEMOptimizer.cpp:732-1035), not from the C oracle.
digamma comes from mpmath (arbitrary precision), so it is independent of the C digamma too.
"""
import math
import sys

import mpmath

DBL_MIN = sys.float_info.min


def digamma(x):
    return float(mpmath.digamma(mpmath.mpf(x)))


def optimize(classes, M, projected, eff_len, unique, *, use_vbem=True, per_txp_prior=True,
             init_uniform=False, eq_class_mode=False, no_rich_eq=False, alt_init=False,
             vb_prior=1e-2, tol=0.01, num_required_frags=5e7, min_iter=100, max_iter=10000):
    """classes: list of (tids, weights, count)."""
    alphas = [float(x) for x in projected]                       # :780
    total_weight = 0.0
    for a in alphas:
        total_weight += a
    eff = [float(x) for x in eff_len]
    alphas_prime = []
    for i in range(M):
        uniq = float(unique[i]) + 0.5
        alphas_prime.append(100.0 if init_uniform else uniq * 1e-3 * eff[i])   # :790-792
    prior = [vb_prior if per_txp_prior else vb_prior * eff[i] for i in range(M)]
    uniform_prior = total_weight / M
    frac = min(0.999, total_weight / num_required_frags)
    for i in range(M):
        if init_uniform:
            alphas[i] = alphas_prime[i]
        else:
            uni = alphas_prime[i] if alt_init else uniform_prior
            alphas[i] = alphas[i] * frac + uni * (1.0 - frac)
        alphas_prime[i] = 1.0
    comb = []
    for tids, ws, count in classes:                              # :830-873
        cw = []
        wsum = 0.0
        for t, w in zip(tids, ws):
            el = max(eff[t], 1.0) if eff[t] > 1.0 else 1.0
            w = 1.0 if no_rich_eq else w
            wt = w if eq_class_mode else count * w * (1.0 / el)
            cw.append(wt)
            wsum += wt
        wn = (1.0 / wsum) if wsum != 0.0 else math.inf   # C++: 1.0/0.0 == inf
        comb.append([x * wn for x in cw])
    valid = []
    for (tids, ws, count), cw in zip(classes, comb):             # :330-394
        denom = 0.0
        for t, a in zip(tids, cw):
            v = alphas[t] * a
            if not math.isnan(v):
                denom += v
        valid.append(denom > DBL_MIN)
    it = 0
    converged = False
    max_rel = -sys.float_info.max
    while it < min_iter or (it < max_iter and not converged):    # :900
        if use_vbem:                                             # :241-328
            s = 0.0
            for i in range(M):
                s += alphas[i] + prior[i]
            log_norm = digamma(s)
            theta = []
            for i in range(M):
                ap = alphas[i] + prior[i]
                theta.append(math.exp(digamma(ap) - log_norm) if ap > 1e-10 else 0.0)
                alphas_prime[i] = 0.0
        else:
            theta = alphas
        for (tids, ws, count), cw, ok in zip(classes, comb, valid):
            if not ok:
                continue
            if len(tids) > 1:
                denom = 0.0
                for t, a in zip(tids, cw):
                    if (not use_vbem) or theta[t] > 0.0:
                        denom += theta[t] * a
                if denom <= DBL_MIN:
                    continue
                inv = count / denom
                for t, a in zip(tids, cw):
                    v = theta[t] * a
                    if (theta[t] > 0.0) if use_vbem else (not math.isnan(v)):
                        alphas_prime[t] += v * inv
            else:
                alphas_prime[tids[0]] += count
        converged = True
        max_rel = -sys.float_info.max
        for i in range(M):                                       # :945-957
            if alphas_prime[i] > 1e-2:
                rel = abs(alphas[i] - alphas_prime[i]) / alphas_prime[i]
                max_rel = max(max_rel, rel)
                if rel > tol:
                    converged = False
            alphas[i] = alphas_prime[i]
            alphas_prime[i] = 0.0
        it += 1
    alpha_sum = 0.0
    for i in range(M):                                           # :1004-1014
        if alphas[i] <= 1e-8:
            alphas[i] = 0.0
        alpha_sum += alphas[i]
    return alphas, it, converged, max_rel, comb, valid
