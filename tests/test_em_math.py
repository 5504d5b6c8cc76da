"""The fused VBEM transform exp(digamma(x) - logNorm) of the product (em_math.h, compiled for the host) against the
oracle's Boost-style digamma (itself pinned to mpmath, tests/golden/digamma_golden.json) and against mpmath directly."""
import ctypes as C
import json
import math
import os
import subprocess

import numpy as np

import oracle_lib as O

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SO = os.path.join(ROOT, "tests", "_build", "libhostemmath.so")
SRC = os.path.join(ROOT, "tests", "host_em_math.cpp")
HDR = os.path.join(ROOT, "salmon_b200", "csrc", "em_math.h")


def _lib():
    os.makedirs(os.path.dirname(SO), exist_ok=True)
    if (not os.path.exists(SO)) or os.path.getmtime(SO) < max(os.path.getmtime(SRC), os.path.getmtime(HDR)):
        subprocess.check_call(["/usr/bin/g++", "-O2", "-std=c++17", "-ffp-contract=off", "-fPIC", "-shared", "-o", SO, SRC])
    return C.CDLL(SO)


def _run(x, ln):
    lib = _lib()
    x = np.ascontiguousarray(x, dtype=np.float64); ln = np.ascontiguousarray(ln, dtype=np.float64)
    out = np.empty_like(x)
    lib.hem_exp_digamma(C.c_uint64(len(x)), x.ctypes.data_as(C.c_void_p), ln.ctypes.data_as(C.c_void_p),
                        out.ctypes.data_as(C.c_void_p))
    return out


def test_fused_transform_vs_oracle_digamma():
    rng = np.random.default_rng(5)
    x = np.concatenate([10.0 ** rng.uniform(-10, 10, 200000), rng.uniform(0.5, 12.0, 100000),
                        np.array([1e-10 * (1 + 1e-9), 0.01, 1.0, 2.0, 9.999999999, 10.0, 10.000000001, 1e8])])
    ln = rng.uniform(0.0, 20.0, len(x))
    got = _run(x, ln)
    psi = np.array([O.digamma(v) for v in x])
    want = np.exp(psi - ln)
    # the exponent carries an absolute rounding error of a few ulp(|psi - logNorm|) in BOTH evaluations
    tol = 16 * np.spacing(np.maximum(1.0, np.maximum(np.abs(psi), np.abs(ln)))) + 4e-16
    ok = want > 1e-300
    rel = np.abs(got[ok] - want[ok]) / want[ok]
    assert np.all(rel <= tol[ok]), (float(rel.max()), x[ok][np.argmax(rel / tol[ok])])
    assert np.all(got[~ok] < 1e-299)
    # where digamma is O(1..20) both evaluations are good to a few ulp of the exponent (<= 20)
    mid = ok & (x > 0.5)
    assert np.max(np.abs(got[mid] - want[mid]) / want[mid]) < 1e-14


def test_fused_transform_vs_mpmath_golden():
    """the committed mpmath digamma values (all branches of the Boost algorithm) through the fused form"""
    g = json.load(open(os.path.join(ROOT, "tests", "golden", "digamma_golden.json")))
    pts = [(float.fromhex(r["x"]), float(r["digamma"])) for r in g["rows"] if float.fromhex(r["x"]) > 1e-10]
    x = np.array([p[0] for p in pts]); psi = np.array([p[1] for p in pts])
    got = _run(x, np.zeros_like(x))
    ok = (psi > -700) & (psi < 700)
    want = np.exp(psi[ok])
    rel = np.abs(got[ok] - want) / want
    tol = 16 * np.spacing(np.maximum(1.0, np.abs(psi[ok]))) + 4e-16
    assert np.all(rel <= tol), float(rel.max())
