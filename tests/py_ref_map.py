"""Independent pure-Python restatement of the salmon-owned per-fragment arithmetic of Stage A, written from the
reference source (not from oracle/map_oracle.c), for the regime of the first batch (no fragments assigned yet: not
burned in, auxiliary parameters off):

  * filterAndCollectAlignments (include/salmon/internal/quant/SalmonMappingUtils.hpp:283-405): kept alignments in
    transcript order, estAlnProb = exp(-scoreExp * (bestScore - score)), dropped below minAlnProb;
  * processMiniBatch (src/quant/SalmonQuantify.cpp:599-857): auxProb = logFragProb + log(estAlnProb) +
    logAlignCompatProb, where an orphan of a paired-end library takes LogCMFCache::getAmbigFragLengthProb
    (src/util/DistributionUtils.cpp:145-177) on the pre-burn-in cache = running logAdd of a constant LOG_EPSILON pmf
    (:104-118, FragmentLengthDistribution.cpp:192-203); weights = exp(auxProb - logAdd of all); range-factorisation bins
    int(weight * (int(sqrt(n)) + rangeFactorizationBins)) appended to the label (:845-853).

Checks an oracle result read by read; libm exp / log here, the oracle's deterministic forms there: tolerances 1e-12."""
import math

LOG_0 = float("inf")
LOG_EPSILON = math.log(0.375e-10)          # SalmonMath.hpp:44-45


def log_add(x, y):                          # SalmonMath.hpp:55-67
    if abs(x) == LOG_0:
        return y
    if abs(y) == LOG_0:
        return x
    if y > x:
        x, y = y, x
    return x + math.log(1 + math.exp(y - x))


def quirk_cmf(max_val):
    """evaluateLogCMF before burn-in: cmf of a pmf that is LOG_EPSILON everywhere on 0..max_val"""
    out, cum = [], LOG_0
    for _ in range(max_val + 1):
        cum = log_add(cum, LOG_EPSILON)
        out.append(cum)
    return out


def cmf_value(cmf, length):                 # LogCMFCache::cmfValue_
    return cmf[length] if length < len(cmf) else cmf[-1]


def ambig_frag_len_prob(cmf, fwd, pos, rlen, tlen):   # LogCMFCache::getAmbigFragLengthProb
    if fwd:
        p1 = 0 if pos < 0 else pos
        p1 = tlen if p1 > tlen else p1
        max_frag_len = tlen - p1
    else:
        p1 = pos + rlen
        p1 = 0 if p1 < 0 else p1
        p1 = tlen if p1 > tlen else p1
        max_frag_len = p1
    ref_cm = cmf_value(cmf, tlen)
    return (cmf_value(cmf, max_frag_len) - ref_cm) if ref_cm != LOG_0 else LOG_EPSILON


def check_read(n_aln, tid, score, prob, pos, flags, label, weight, read_len, tx_len, cmf, *, score_exp=1.0,
               min_aln_prob=1e-5, range_bins=4, max_read_occ=200):
    """Raises AssertionError when the alignments of one read contradict the reference's arithmetic.
    flags: bit 0 = this mate forward, bit 1 = mate forward, bits 2-3 = 0 paired / 1 left orphan / 2 right orphan."""
    n = int(n_aln)
    assert 0 <= n <= max_read_occ
    if n == 0:
        return 0
    t = [int(x) for x in tid[:n]]
    assert t == sorted(t), "alignments must be in transcript order (keptPerm sorted by tid)"
    best = max(int(s) for s in score[:n])        # the best-scoring hit always survives the soft filter
    aux = []
    for a in range(n):
        est = math.exp(-score_exp * (best - int(score[a])))
        assert est >= min_aln_prob
        assert abs(est - prob[a]) <= 1e-12 * est, (est, prob[a])
        status = (int(flags[a]) >> 2) & 3
        log_frag_prob = 0.0
        if status != 0:                          # isUnexpectedOrphan in a paired-end library, modelSingleFragProb on
            log_frag_prob = ambig_frag_len_prob(cmf, bool(int(flags[a]) & 1), int(pos[a]), read_len, int(tx_len[t[a]]))
        aux.append(log_frag_prob + (math.log(est) if est > 0 else 0.0) + 0.0)      # compatible: LOG_1
    denom = LOG_0
    for x in aux:
        denom = log_add(denom, x)
    range_count = int(math.sqrt(n)) + range_bins
    for a in range(n):
        w = math.exp(aux[a] - denom)
        assert abs(w - weight[a]) <= 1e-12, (w, weight[a])
        assert int(label[a]) == t[a]
        if range_bins > 0:
            x = w * range_count
            if abs(x - round(x)) > 1e-9:          # away from a bin edge the bin is unambiguous
                assert int(label[n + a]) == int(x), (int(label[n + a]), x)
    return n


# ---- normalizeAlphas (src/util/SalmonUtils.cpp:461-529) + TranscriptCluster::projectToPolytope
# (include/salmon/internal/quant/TranscriptCluster.hpp:46-101), restated from the reference ------------------------------
def normalize_alphas(log_mass, classes):
    """log_mass[t]: Transcript::mass(false) (log scale, +inf = none).  classes: iterable of (transcript ids, count).
    Clusters = connected components of the transcripts that share a fragment (ClusterForest::mergeClusters), a cluster's
    numHits = its fragments.  -> (projectedCounts, uniqueCounts, totalCounts)"""
    M = len(log_mass)
    parent = list(range(M))

    def find(x):
        while parent[x] != x:
            parent[x] = parent[parent[x]]
            x = parent[x]
        return x
    uniq = [0] * M
    total = [0] * M
    for tids, cnt in classes:
        for t in tids:
            total[t] += cnt
        if len(tids) == 1:
            uniq[tids[0]] += cnt
        for t in tids[1:]:
            a, b = find(tids[0]), find(t)
            if a != b:
                parent[max(a, b)] = min(a, b)
    hits = {}
    for tids, cnt in classes:
        r = find(tids[0])
        hits[r] = hits.get(r, 0) + cnt
    members = {}
    for t in range(M):
        members.setdefault(find(t), []).append(t)
    proj = [0.0] * M
    for root, mem in members.items():
        if root not in hits:
            continue
        log_cluster_mass = LOG_0
        for t in mem:
            log_cluster_mass = log_add(log_cluster_mass, log_mass[t])
        log_cluster_count = math.log(float(hits[root]))
        needs = False
        for t in mem:
            if log_mass[t] == LOG_0:
                proj[t] = 0.0
            else:
                proj[t] = math.exp(log_mass[t] - log_cluster_mass + log_cluster_count)
                needs |= proj[t] > float(total[t]) or proj[t] < float(uniq[t])
        if len(mem) > 1 and needs:
            project_to_polytope(mem, proj, uniq, total, float(hits[root]))
    return proj, uniq, total


def project_to_polytope(mem, proj, uniq, total, cluster_counts):
    bound = [False] * len(mem)
    rounds = 0
    while True:
        unbound_counts = bound_counts = 0.0
        for i, t in enumerate(mem):
            if proj[t] > total[t]:
                proj[t] = float(total[t]); bound[i] = True
            elif proj[t] < uniq[t]:
                proj[t] = float(uniq[t]); bound[i] = True
            if bound[i]:
                bound_counts += proj[t]
            else:
                unbound_counts += proj[t]
        if abs(unbound_counts + bound_counts - cluster_counts) <= 0.375e-10:      # approxEqual
            return
        if unbound_counts == 0:
            bound = [False] * len(mem)
            unbound_counts, bound_counts = bound_counts, 0.0
        normalizer = (cluster_counts - bound_counts) / unbound_counts
        for i, t in enumerate(mem):
            if not bound[i]:
                proj[t] *= normalizer
        rounds += 1
        if rounds > 5000:
            return


# ---- FragmentLengthDistribution prior (src/model/FragmentLengthDistribution.cpp:22-78, pmf :114-125, cmf :141-156) -----
def fld_prior_tables(mean, sd, max_val, alpha=1.0):
    """log pmf / log cmf over 0..max_val of the prior N(mean, sd) discretised per unit bin (bin_size 1)."""
    def ncdf(x):
        return 0.5 * math.erfc(-(x - mean) / (sd * math.sqrt(2.0)))
    tot = math.log(alpha)
    hist, tot_mass = [], LOG_0
    for i in range(max_val + 1):
        norm_mass = ncdf(i + 0.5) - ncdf(i - 0.5)
        mass = LOG_EPSILON if norm_mass == 0 else tot + math.log(norm_mass)
        hist.append(mass)
        tot_mass = log_add(tot_mass, mass)
    pmf = [h - tot_mass for h in hist]
    cmf, cum = [], LOG_0
    for h in hist:
        cum = log_add(cum, h)
        cmf.append(cum - tot_mass)
    return pmf, cmf
