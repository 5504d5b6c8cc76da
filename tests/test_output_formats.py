"""CPU tests of the output seam (host code of the library): TPM as writeAbundances computes it, quant.sf and
eq_classes.txt[.gz] in the reference's layouts (src/output/GZipWriter.cpp:64-168, 684-739)."""
import gzip

import numpy as np

from salmon_b200 import _capi


def test_tpm_matches_oracle(oracle):
    rng = np.random.default_rng(3)
    alpha = rng.random(500) * 100
    alpha[rng.random(500) < 0.3] = 0.0
    eff = rng.random(500) * 2000 + 1
    got = _capi.tpm(alpha, eff)
    assert abs(got.sum() - 1e6) < 1e-3
    ref = (alpha / eff) / (alpha / eff).sum() * 1e6
    np.testing.assert_allclose(got, ref, rtol=1e-12)


def test_quant_sf_layout(tmp_path):
    a = np.array([10.0, 0.0, 5.5, 100.25]); e = np.array([100.0, 50.0, 20.0, 1000.0])
    p = tmp_path / "quant.sf"
    _capi.write_quant_sf(p, ["t1", "t2", "t3", "t4"], [300, 200, 100, 1200], e, a)
    lines = p.read_text().splitlines()
    assert lines[0] == "Name\tLength\tEffectiveLength\tTPM\tNumReads"
    assert lines[1] == "t1\t300\t100.000\t210415.570752\t10.000"
    assert lines[2] == "t2\t200\t50.000\t0.000000\t0.000"
    assert len(lines) == 5


def test_eq_classes_layouts(tmp_path):
    names = ["t1", "t2", "t3", "t4"]
    off, tids, counts = [0, 2, 3, 5], [0, 1, 2, 0, 1], [5, 7, 1]
    p = tmp_path / "eq_classes.txt.gz"
    _capi.write_eq_classes(p, names, off, tids, counts, weights=[0.25, 0.75, 1.0, 1 / 3, 2 / 3])
    txt = gzip.open(p, "rt").read().splitlines()
    assert txt[:2] == ["4", "3"] and txt[2:6] == names
    assert txt[6] == "2\t0\t1\t0.25\t0.75\t5" and txt[7] == "1\t2\t1\t7" and txt[8] == "2\t0\t1\t0.333333\t0.666667\t1"
    # without weights range-factorised classes collapse by transcript set (GZipWriter.cpp:86-113)
    p2 = tmp_path / "eq_classes.txt"
    _capi.write_eq_classes(p2, names, off, tids, counts)
    txt = p2.read_text().splitlines()
    assert txt[:2] == ["4", "2"] and txt[6:] == ["2\t0\t1\t6", "1\t2\t7"]
