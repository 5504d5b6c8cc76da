"""CPU-side checks of the C-ABI library: it loads, exports every symbol that
include/salmon_b200.h declares, and fails loudly (no CPU fallback) without a GPU."""
import ctypes
import os
import re

import pytest

from salmon_b200 import _capi

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_symbols():
    names = set()
    inc = os.path.join(ROOT, "include")
    for fn in os.listdir(inc):
        if fn.endswith(".h"):
            txt = open(os.path.join(inc, fn)).read()
            txt = re.sub(r"/\*.*?\*/", "", txt, flags=re.S)
            names |= set(re.findall(r"\b(sb_[a-z0-9_]+)\s*\(", txt))
    return names


def test_library_exports_every_declared_symbol():
    lib = ctypes.CDLL(_capi.LIB_PATH)
    decl = declared_symbols()
    assert len(decl) >= 15
    for name in sorted(decl):
        assert hasattr(lib, name), f"{name} declared in include/*.h but not exported"
    # and the Python binding covers them all
    assert decl <= set(_capi.SYMBOLS), sorted(decl - set(_capi.SYMBOLS))


def test_version_and_params():
    lib = _capi.load()
    assert lib.sb_version() >= 100
    p = _capi.default_params()
    assert p.use_vbem == 1 and p.per_txp_prior == 1
    assert p.vb_prior == 1e-2 and p.tol == 0.01 and p.min_iter == 100 and p.max_iter == 10000
    assert p.num_required_frags == 5e7


def test_no_cpu_fallback_without_gpu():
    lib = _capi.load()
    if lib.sb_device_count() > 0:
        pytest.skip("a GPU is present")
    with pytest.raises(_capi.SalmonB200Error, match="no CUDA device"):
        _capi.EMContext(0)


def test_no_collective_inside_rank_conditional_blocks():
    """a collective that only one rank reaches hangs the run (that happened: an all-reduce inside the rank-0 output block
    of sb_quant_files): in the C++ multi-GPU driver no sb_comm_all* call may sit inside an `if (... shard_index == 0 ...)`
    or `if (r != 0 ...)` block"""
    import re
    src = open(os.path.join(ROOT, "salmon_b200", "csrc", "pipeline.cu")).read()
    bad = []
    for m in re.finditer(r"if \([^\n]*(shard_index [!=]= 0|\br [!=]= 0|rank [!=]= 0)[^\n]*\) \{", src):
        depth, i = 1, m.end()
        while depth and i < len(src):
            depth += {"{": 1, "}": -1}.get(src[i], 0)
            i += 1
        if "sb_comm_all" in src[m.end():i]:
            bad.append(src[m.start():m.end()])
    assert not bad, bad


def test_library_type_detection_rule():
    """sb_detect_lib_type restates LibraryTypeDetector::mostLikelyType (LibraryTypeDetector.hpp:34-140) for inward pairs
    and unmated reads: sense fraction < 0.3 -> antisense type, < 0.7 -> unstranded, else sense; nothing seen -> -1"""
    import ctypes as C
    from salmon_b200 import _capi
    lib = _capi.load()
    IU, ISF, ISR, U, SF, SR = range(6)

    def det(paired, isf, isr, sf, sr):
        a = (C.c_uint64 * 4)(isf, isr, sf, sr)
        return lib.sb_detect_lib_type(1 if paired else 0, a)
    assert det(True, 0, 0, 5, 5) == -1 and det(False, 7, 7, 0, 0) == -1
    assert det(True, 29, 71, 0, 0) == ISR and det(True, 30, 70, 0, 0) == IU and det(True, 69, 31, 0, 0) == IU
    assert det(True, 70, 30, 0, 0) == ISF and det(True, 100, 0, 0, 0) == ISF and det(True, 0, 9, 0, 0) == ISR
    assert det(False, 0, 0, 29, 71) == SR and det(False, 0, 0, 50, 50) == U and det(False, 0, 0, 70, 30) == SF
    assert lib.sb_detect_lib_type(1, None) == -1
