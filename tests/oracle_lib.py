"""ctypes binding of oracle/liboracle.so -- TEST INFRASTRUCTURE ONLY.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference
legs may import this module.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
ORACLE_DIR = os.path.join(ROOT, "oracle")
LIB = os.path.join(ORACLE_DIR, "liboracle.so")


class orc_em_params(C.Structure):
    _fields_ = [
        ("use_vbem", C.c_int32), ("per_txp_prior", C.c_int32), ("init_uniform", C.c_int32),
        ("eq_class_mode", C.c_int32), ("no_rich_eq", C.c_int32), ("no_length_correction", C.c_int32),
        ("alt_init", C.c_int32), ("n_threads", C.c_int32),
        ("vb_prior", C.c_double), ("tol", C.c_double), ("num_required_frags", C.c_double),
        ("min_iter", C.c_uint32), ("max_iter", C.c_uint32),
    ]


class orc_em_stats(C.Structure):
    _fields_ = [
        ("iters", C.c_uint32), ("converged", C.c_uint32), ("max_rel_diff", C.c_double),
        ("alpha_sum", C.c_double), ("n_degenerate", C.c_uint64), ("ok", C.c_int32),
    ]


_lib = None


def build():
    subprocess.check_call(["make", "-C", ORACLE_DIR, "-s"])


def load():
    global _lib
    if _lib is None:
        if not os.path.exists(LIB):
            build()
        lib = C.CDLL(LIB)
        lib.orc_digamma.restype = C.c_double
        lib.orc_digamma.argtypes = [C.c_double]
        _lib = lib
    return _lib


def params_from(p, n_threads=0) -> orc_em_params:
    """Copy the shared fields of an sb_em_params (or dict) into orc_em_params."""
    o = orc_em_params()
    for k in ("use_vbem", "per_txp_prior", "init_uniform", "eq_class_mode", "no_rich_eq",
              "no_length_correction", "alt_init", "vb_prior", "tol", "num_required_frags",
              "min_iter", "max_iter"):
        setattr(o, k, p[k] if isinstance(p, dict) else getattr(p, k))
    o.n_threads = n_threads
    return o


def _p(a):
    return a.ctypes.data_as(C.c_void_p) if a is not None else None


def em_optimize(eq, projected, eff_len, unique, p, mt=False, n_threads=0, want_combined=False):
    lib = load()
    M = eq.n_txps
    op = params_from(p, n_threads)
    alpha = np.empty(M, dtype=np.float64)
    st = orc_em_stats()
    projected = np.ascontiguousarray(projected, dtype=np.float64)
    eff_len = np.ascontiguousarray(eff_len, dtype=np.float64)
    unique = np.ascontiguousarray(unique, dtype=np.uint64)
    if mt:
        rc = lib.orc_em_optimize_mt(C.c_uint64(eq.n_classes), C.c_uint32(M), _p(eq.off), _p(eq.tids),
                                    _p(eq.weights), _p(eq.counts), _p(projected), _p(eff_len),
                                    _p(unique), C.byref(op), _p(alpha), C.byref(st))
        assert rc == 0
        return alpha, st
    cw = np.empty(eq.nnz, dtype=np.float64) if want_combined else None
    valid = np.empty(eq.n_classes, dtype=np.uint8) if want_combined else None
    rc = lib.orc_em_optimize(C.c_uint64(eq.n_classes), C.c_uint32(M), _p(eq.off), _p(eq.tids),
                             _p(eq.weights), _p(eq.counts), _p(projected), _p(eff_len), _p(unique),
                             C.byref(op), _p(alpha), _p(cw), _p(valid), None, C.byref(st))
    assert rc == 0
    if want_combined:
        return alpha, st, cw, valid
    return alpha, st


def em_step(eq, cw, valid, prior, alpha_in, vbem, serial=False, alpha_out_init=None):
    lib = load()
    M = eq.n_txps
    out = np.zeros(M) if alpha_out_init is None else np.array(alpha_out_init, dtype=np.float64)
    th = np.zeros(M)
    fn = lib.orc_em_step_serial if serial else lib.orc_em_step
    fn(C.c_uint64(eq.n_classes), C.c_uint32(M), _p(eq.off), _p(eq.tids), _p(np.ascontiguousarray(cw)),
       _p(eq.counts), _p(valid), _p(np.ascontiguousarray(prior, dtype=np.float64)),
       _p(np.ascontiguousarray(alpha_in, dtype=np.float64)), _p(out), _p(th), C.c_int(vbem))
    return out, th


def tpm(alpha, eff_len):
    lib = load()
    out = np.empty_like(alpha)
    lib.orc_tpm(C.c_uint32(alpha.shape[0]), _p(np.ascontiguousarray(alpha)),
                _p(np.ascontiguousarray(eff_len, dtype=np.float64)), _p(out))
    return out


def digamma(x):
    return load().orc_digamma(float(x))


def philox(c, k):
    out = (C.c_uint32 * 4)()
    load().orc_philox4x32(C.c_uint32(c[0]), C.c_uint32(c[1]), C.c_uint32(c[2]), C.c_uint32(c[3]),
                          C.c_uint32(k[0]), C.c_uint32(k[1]), out)
    return list(out)


def bootstrap(eq, cw, valid, prior, active, p, n_boot, seed):
    """orc_bootstrap -> (alphas [n_boot, M], sampled counts [n_boot, C], rc)."""
    lib = load()
    M, Cn = eq.n_txps, eq.n_classes
    op = params_from(p)
    alphas = np.zeros((n_boot, M))
    samp = np.zeros((n_boot, Cn), dtype=np.uint64)
    rc = lib.orc_bootstrap(C.c_uint64(Cn), C.c_uint32(M), _p(eq.off), _p(eq.tids), _p(np.ascontiguousarray(cw)),
                           _p(eq.counts), _p(np.ascontiguousarray(valid, dtype=np.uint8)),
                           _p(np.ascontiguousarray(prior, dtype=np.float64)),
                           _p(np.ascontiguousarray(active, dtype=np.uint8)), C.byref(op), C.c_uint32(n_boot),
                           C.c_uint64(seed), _p(alphas), _p(samp))
    return alphas, samp, rc


def gibbs(eq, valid, eff_len, alphas_init, use_vbem, per_txp_prior, vb_prior, n_samples, thinning, no_gamma_draw,
          num_mapped_frags, seed):
    lib = load()
    M = eq.n_txps
    out = np.zeros((n_samples, M))
    rc = lib.orc_gibbs(C.c_uint64(eq.n_classes), C.c_uint32(M), _p(eq.off), _p(eq.tids), _p(eq.weights),
                       _p(eq.counts), _p(np.ascontiguousarray(valid, dtype=np.uint8)),
                       _p(np.ascontiguousarray(eff_len, dtype=np.float64)),
                       _p(np.ascontiguousarray(alphas_init, dtype=np.float64)), C.c_int(use_vbem),
                       C.c_int(per_txp_prior), C.c_double(vb_prior), C.c_uint32(n_samples), C.c_uint32(thinning),
                       C.c_int(no_gamma_draw), C.c_double(num_mapped_frags), C.c_uint64(seed), _p(out))
    assert rc == 0
    return out
