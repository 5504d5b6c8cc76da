"""TEST INFRASTRUCTURE: runs the product's per-read Stage A logic (map_core.h) compiled for the
host (tests/host_map_core.cpp) -- the same code the CUDA kernels execute -- so it can be checked
against the independent oracle without a GPU."""
import ctypes as C
import os
import subprocess

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SO = os.path.join(ROOT, "tests", "_build", "libhostmap.so")
SRC = os.path.join(ROOT, "tests", "host_map_core.cpp")
HDR = os.path.join(ROOT, "salmon_b200", "csrc", "map_core.h")


def build():
    os.makedirs(os.path.dirname(SO), exist_ok=True)
    if (not os.path.exists(SO)) or os.path.getmtime(SO) < max(os.path.getmtime(SRC), os.path.getmtime(HDR)):
        subprocess.check_call(["/usr/bin/g++", "-O3", "-std=c++17", "-ffp-contract=off", "-fopenmp", "-fPIC", "-shared",
                               "-I" + os.path.join(ROOT, "include"), "-o", SO, SRC])
    return C.CDLL(SO)


def fld_tables(params):
    """the four FLD tables of the prior, as the product's host code builds them"""
    import oracle_lib as O
    import math
    nf = params.max_frag_len + 1
    fld = np.zeros(4 * nf)
    pm = np.zeros(nf); cm = np.zeros(nf)
    O.load().orc_fld_tables(C.c_double(params.fld_mean), C.c_double(params.fld_sd), C.c_uint32(params.max_frag_len),
                            pm.ctypes.data_as(C.c_void_p), cm.ctypes.data_as(C.c_void_p))
    fld[:nf] = pm
    la = lambda x, y: y if math.isinf(x) else (max(x, y) + math.log(1 + math.exp(min(x, y) - max(x, y))))
    tm = math.inf
    for i in range(nf):
        tm = la(tm, pm[i])
    cum = math.inf; cq = math.inf; le = math.log(0.375e-10)
    for i in range(nf):
        pc = pm[i] - tm
        fld[nf + i] = pc
        cum = la(cum, pc); fld[2 * nf + i] = cum
        cq = la(cq, le); fld[3 * nf + i] = cq
    return fld


def map_throughput(index, params, left, right, frag_counter=0, n_threads=0):
    """CPU port (OpenMP over reads) of the per-read Stage A logic; returns (seconds, n_aln, counters)."""
    import time
    lib = build()
    ha = index.host_arrays()
    n, L = left.shape
    left = np.ascontiguousarray(left, dtype=np.uint8); right = np.ascontiguousarray(right, dtype=np.uint8)
    fld = fld_tables(params)
    n_aln = np.zeros(n, dtype=np.uint32)
    ctr = np.zeros(7, dtype=np.uint64)
    t0 = time.perf_counter()
    rc = lib.hmc_map_throughput(C.c_uint32(index.n_txps), C.c_uint32(index.k), C.c_void_p(ha["tx_off"]),
                                C.c_void_p(ha["codes"]), C.c_void_p(ha["table"]), C.c_uint64(ha["table_capacity"]),
                                C.c_void_p(ha["postings"]), C.byref(params), fld.ctypes.data_as(C.c_void_p),
                                left.ctypes.data_as(C.c_void_p), right.ctypes.data_as(C.c_void_p), C.c_uint32(n),
                                C.c_uint32(L), C.c_uint64(frag_counter), C.c_int(n_threads),
                                n_aln.ctypes.data_as(C.c_void_p), ctr.ctypes.data_as(C.c_void_p))
    dt = time.perf_counter() - t0
    assert rc == 0
    return dt, n_aln, dict(zip(("lookups", "postings", "seeds", "candidates", "kept", "label_entries", "mapped"),
                               [int(x) for x in ctr]))


def map_reads(index, params, left, right, frag_counter=0):
    """index: salmon_b200._capi.Index; params: sb_map_params."""
    import oracle_lib as O
    from salmon_b200._capi import alloc_alignment_arrays
    lib = build()
    ha = index.host_arrays()
    n, L = left.shape
    left = np.ascontiguousarray(left, dtype=np.uint8); right = np.ascontiguousarray(right, dtype=np.uint8)
    nf = params.max_frag_len + 1
    # FLD tables exactly as the product builds them on the host: same libm calls as the oracle's
    fld = np.zeros(4 * nf)
    olib = O.load()
    pm = np.zeros(nf); cm = np.zeros(nf)
    olib.orc_fld_tables(C.c_double(params.fld_mean), C.c_double(params.fld_sd), C.c_uint32(params.max_frag_len),
                        pm.ctypes.data_as(C.c_void_p), cm.ctypes.data_as(C.c_void_p))
    fld[:nf] = pm
    # cached / quirk tables: recompute like the oracle (logAdd chains)
    def logadd(x, y):
        if np.isinf(x): return y
        if np.isinf(y): return x
        if y > x: x, y = y, x
        return x + np.log1p(np.exp(y - x)) if False else x + np.log(1 + np.exp(y - x))
    import math
    tm = math.inf
    for i in range(nf):
        tm = pm[i] if math.isinf(tm) else (max(tm, pm[i]) + math.log(1 + math.exp(min(tm, pm[i]) - max(tm, pm[i]))))
    cum = math.inf; cq = math.inf; le = math.log(0.375e-10)
    for i in range(nf):
        pc = pm[i] - tm
        fld[nf + i] = pc
        cum = pc if math.isinf(cum) else (max(cum, pc) + math.log(1 + math.exp(min(cum, pc) - max(cum, pc))))
        fld[2 * nf + i] = cum
        cq = le if math.isinf(cq) else (max(cq, le) + math.log(1 + math.exp(min(cq, le) - max(cq, le))))
        fld[3 * nf + i] = cq
    a = alloc_alignment_arrays(n, params.max_read_occ)
    ctr = np.zeros(7, dtype=np.uint64)
    rc = lib.hmc_map_reads(C.c_uint32(index.n_txps), C.c_uint32(index.k), C.c_void_p(ha["tx_off"]), C.c_void_p(ha["codes"]),
                           C.c_void_p(ha["table"]), C.c_uint64(ha["table_capacity"]), C.c_void_p(ha["postings"]),
                           C.byref(params), fld.ctypes.data_as(C.c_void_p), left.ctypes.data_as(C.c_void_p),
                           right.ctypes.data_as(C.c_void_p), C.c_uint32(n), C.c_uint32(L), C.c_uint64(frag_counter),
                           *[a[k].ctypes.data_as(C.c_void_p) for k in ("n_aln", "tid", "score", "prob", "pos", "mate_pos",
                                                                      "flags", "flen", "label", "weight")],
                           ctr.ctypes.data_as(C.c_void_p))
    assert rc == 0
    a["counters"] = dict(zip(("lookups", "postings", "seeds", "candidates", "kept", "label_entries", "mapped"),
                             [int(x) for x in ctr]))
    return a
