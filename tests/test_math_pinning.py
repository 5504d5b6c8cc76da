"""Pinning of the Stage A floating-point arithmetic to the PLATFORM libm (VERDICT r1, weak #1 / next #3a).

The reference evaluates the per-alignment model with std::exp / std::log (SalmonMath.hpp:40-81, SalmonQuantify.cpp:
599-857) and truncates normalised weights into range-factorisation bins (:846-851).  The product's device code cannot
call the host libm; it implements the published fdlibm algorithms (include/sb_detmath.h).  The oracle's default mode IS
the libm (oracle/orc_math.h includes nothing of the product); these tests show that
  * the oracle in libm mode and in its fdlibm-restatement mode assign identical labels and bins, weights to a few ulp,
  * the product's per-read logic (map_core.h compiled for the host) gives the libm oracle's labels and bins,
and account for every read that differs: such a read must sit on a bin boundary (a normalised weight times the number
of bins within 1e-9 of an integer) and the two bin vectors may differ by one at that position only."""
import ctypes as C

import numpy as np

import oracle_lib as O
from salmon_b200._capi import Index, map_default_params
from salmon_b200.synth import synth_reads, synth_reads_fast, synth_txome


def label_flips(a, b, cap, binned=True):
    """reads whose labels differ between two result sets; asserts everything else is identical / close"""
    assert np.array_equal(a["n_aln"], b["n_aln"])
    na = a["n_aln"]
    m = np.arange(cap)[None, :] < na[:, None]
    for k in ("tid", "score", "pos", "mate_pos", "flags", "flen"):
        assert np.array_equal(a[k][m], b[k][m]), k
    np.testing.assert_allclose(a["weight"][m], b["weight"][m], rtol=1e-13, atol=0)
    np.testing.assert_allclose(a["prob"][m], b["prob"][m], rtol=1e-13, atol=0)
    m2 = np.arange(2 * cap)[None, :] < 2 * na[:, None]
    diff = np.where(((a["label"] != b["label"]) & m2).any(axis=1))[0]
    for r in diff:            # every flip is a boundary case, off by one at that position
        k = int(na[r]); nb = int(np.sqrt(k)) + 4
        assert np.array_equal(a["label"][r, :k], b["label"][r, :k])
        ba, bb = a["label"][r, k:2 * k].astype(np.int64), b["label"][r, k:2 * k].astype(np.int64)
        pos = np.where(ba != bb)[0]
        x = b["weight"][r, :k] * nb
        assert np.all(np.abs(ba[pos] - bb[pos]) == 1) and np.all(np.abs(x[pos] - np.round(x[pos])) < 1e-9), (r, ba, bb, x)
    return diff


def test_fdlibm_restatement_vs_libm_scalars():
    """the oracle's two math modes on random arguments: never more than 1 ulp apart, identical in the vast majority"""
    lib = O.load()
    lib.orc_math_probe.restype = None
    rng = np.random.default_rng(3)
    x = np.concatenate([rng.uniform(-60, 5, 400000), -rng.exponential(3.0, 400000), rng.uniform(-1e-6, 1e-6, 1000)])
    y = np.concatenate([10.0 ** rng.uniform(-12, 9, 400000), rng.uniform(0.5, 2.0, 400000)])
    out = {}
    for mode in (0, 1):
        e = np.empty_like(x); l = np.empty_like(y)
        lib.orc_math_probe(C.c_int(mode), C.c_uint64(len(x)), x.ctypes.data_as(C.c_void_p), e.ctypes.data_as(C.c_void_p),
                           C.c_uint64(len(y)), y.ctypes.data_as(C.c_void_p), l.ctypes.data_as(C.c_void_p))
        out[mode] = (e, l)
    for i, name in ((0, "exp"), (1, "log")):
        a, b = out[0][i], out[1][i]
        ulp = np.abs(a.view(np.int64) - b.view(np.int64))
        assert ulp.max() <= 1, (name, int(ulp.max()))
        assert (ulp == 0).mean() > 0.85, (name, float((ulp == 0).mean()))    # fdlibm and glibc are both < 1 ulp, not identical


def test_oracle_libm_vs_fdlibm_labels():
    txps, _ = synth_txome(seed=44, n_genes=500)
    left, right, _ = synth_reads_fast(txps, seed=7, n=80000)
    idx = O.MapIndex(txps); p = O.map_params()
    for fc in (0, 6000, 6_000_000):                 # the three regimes of the auxiliary model
        fd = O.map_reads(idx, p, left[:40000] if fc else left, right[:40000] if fc else right, fc)
        with O.math_mode("libm"):
            lm = O.map_reads(idx, p, left[:40000] if fc else left, right[:40000] if fc else right, fc)
        flips = label_flips(fd, lm, p.max_read_occ)          # (asserts that every flip is an exact-boundary case)
        assert (fd["n_aln"] > 1).sum() > 20000
        # Measured here: 0 flips before the fragment-length model is in use, ~1.2 % of the reads once it is -- reads whose
        # alignments all have the same auxiliary probability, so that a normalised weight times the number of bins is an
        # integer in exact arithmetic (e.g. 6 x 1/6) and the truncation depends on the last bit of exp / log.  On those
        # reads the reference's own label depends on the libm it was linked with; everywhere else the labels are equal.
        assert len(flips) <= 0.02 * len(fd["n_aln"]), len(flips)


def test_product_host_logic_vs_libm_oracle():
    """map_core.h (the code the kernels execute, compiled for the host with sb_detmath.h) against the oracle on libm"""
    import hostmap_lib
    txps, _ = synth_txome(seed=45, n_genes=300)
    left, right, _ = synth_reads(txps, seed=9, n=30000)
    idx = Index(txps); p = map_default_params()
    got = hostmap_lib.map_reads(idx, p, left, right, 0)
    with O.math_mode("libm"):
        ref = O.map_reads(O.MapIndex(txps), O.map_params(), left, right, 0)
    flips = label_flips(got, ref, p.max_read_occ)
    assert len(flips) <= 0.02 * len(got["n_aln"]), len(flips)
    for k in ("lookups", "postings", "seeds", "kept", "label_entries", "mapped"):
        assert got["counters"][k] == ref["counters"][k], k


def test_oracle_digamma_vs_reference_tree_eigen():
    """the oracle's digamma (checker of the fused VBEM transform) against the digamma the reference tree itself vendors
    (Eigen's Cephes-derived implementation, compiled by oracle/build_ref.sh from the headers under /root/reference through
    a five-line shim; boost::math::digamma, which salmon calls, is not in the tree).  Skipped when oracle/_ref is absent."""
    import ctypes as C
    import os
    import numpy as np
    import pytest
    import oracle_lib as O
    so = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "oracle", "_ref", "libeigen_digamma_ref.so")
    if not os.path.exists(so):
        pytest.skip("oracle/_ref/libeigen_digamma_ref.so not built (needs /root/reference)")
    lib = C.CDLL(so)
    rng = np.random.default_rng(11)
    x = np.concatenate([10.0 ** rng.uniform(-10, 9, 40000), np.linspace(0.01, 40.0, 20000), rng.uniform(1.3, 1.6, 5000)])
    ref = np.empty_like(x)
    lib.ref_eigen_digamma(C.c_ulong(len(x)), x.ctypes.data_as(C.c_void_p), ref.ctypes.data_as(C.c_void_p))
    mine = np.array([O.digamma(v) for v in x])
    err = np.abs(mine - ref)
    # relative where digamma is not near its root (x0 = 1.4616...), absolute there
    far = np.abs(ref) > 1e-2
    assert np.all(err[far] <= 2e-13 * np.abs(ref[far])), float((err[far] / np.abs(ref[far])).max())
    assert np.all(err[~far] <= 1e-14), float(err[~far].max())
