"""File-level pipelines on the GPU: FASTQ files -> quant.sf (quant_files) and --eqclasses -> quant.sf (quant_eqclasses).
The mapping / EM kernels are checked against the oracle elsewhere; here the file seams must not change results."""
import gzip
import os

import numpy as np
import pytest

import oracle_lib as O
from salmon_b200 import _capi
from salmon_b200._capi import Index
from salmon_b200.synth import synth_eq, synth_reads, synth_txome

pytestmark = pytest.mark.gpu
LETTERS = np.frombuffer(b"ACGTN", dtype=np.uint8)


def write_fastq(path, codes, lens=None, gz=True):
    op = gzip.open if gz else open
    with op(path, "wb") as f:
        for i, row in enumerate(codes):
            L = len(row) if lens is None else int(lens[i])
            s = LETTERS[row[:L]].tobytes()
            f.write(b"@r%d\n%s\n+\n%s\n" % (i, s, b"I" * L))


def test_quant_files_equals_quant_reads(tmp_path):
    from salmon_b200.quant import quant_files, quant_reads
    txps, _ = synth_txome(seed=41, n_genes=120)
    left, right, truth = synth_reads(txps, seed=43, n=20000)
    names = [f"ENST{i:05d}" for i in range(len(txps))]
    idx = Index(txps, names=names)
    f1, f2 = str(tmp_path / "r_1.fq.gz"), str(tmp_path / "r_2.fq.gz")
    write_fastq(f1, left); write_fastq(f2, right)
    ipath = str(tmp_path / "sb_index.bin")
    idx.save(ipath)
    a = quant_reads(idx, left, right, batch=8192)
    b = quant_files(ipath, f1, f2, out_dir=str(tmp_path / "out"), batch=8192, max_read_len=left.shape[1],
                    dump_eq_weights=True, num_bootstraps=3, seed=7)
    assert b["n_observed"] == 20000 and a["n_mapped"] == b["n_mapped"]
    assert np.array_equal(a["classes"]["counts"], b["classes"]["counts"])
    assert np.array_equal(a["classes"]["tids"], b["classes"]["tids"])
    assert np.array_equal(a["alpha"], b["alpha"])          # same batches, same order: bit-identical
    lines = (tmp_path / "out" / "quant.sf").read_text().splitlines()
    assert lines[1].split("\t")[0] == "ENST00000" and len(lines) == len(txps) + 1
    raw = gzip.open(tmp_path / "out" / "aux_info" / "bootstrap" / "bootstraps.gz", "rb").read()
    boots = np.frombuffer(raw, dtype=np.float64).reshape(3, len(txps))
    assert np.array_equal(boots, b["bootstraps"])
    assert np.all(np.abs(boots.sum(axis=1) - b["n_mapped"]) < 1e-3 * b["n_mapped"])
    # the eq file written by the run feeds the --eqclasses mode
    from salmon_b200.quant import quant_eqclasses
    c = quant_eqclasses(str(tmp_path / "out" / "aux_info" / "eq_classes.txt.gz"))
    assert abs(c["alpha"].sum() - b["n_mapped"]) < 1e-6 * b["n_mapped"]


def test_quant_files_variable_read_lengths(tmp_path):
    from salmon_b200.quant import quant_files
    txps, _ = synth_txome(seed=41, n_genes=120)
    left, right, truth = synth_reads(txps, seed=44, n=12000)
    n, L = left.shape
    rng = np.random.default_rng(3)
    ll = np.full(n, L); lr = np.full(n, L)
    trim = rng.random(n) < 0.4                      # adapter-trimmed reads: 3' ends removed
    ll[trim] = rng.integers(20, L + 1, size=int(trim.sum()))
    lr[trim] = np.where(rng.random(int(trim.sum())) < 0.5, ll[trim], rng.integers(20, L + 1, size=int(trim.sum())))
    f1, f2 = str(tmp_path / "v_1.fq"), str(tmp_path / "v_2.fq")
    write_fastq(f1, left, ll, gz=False); write_fastq(f2, right, lr, gz=False)
    idx = Index(txps)
    out = quant_files(idx, f1, f2, batch=4096, max_read_len=128)
    M = len(txps)
    mappable = truth["tid"] >= 0
    true_counts = np.bincount(truth["tid"][mappable], minlength=M).astype(float)
    assert out["n_observed"] == n
    assert out["n_mapped"] >= 0.9 * mappable.sum()      # reads trimmed below k (31) cannot map
    assert abs(out["alpha"].sum() - out["n_mapped"]) < 1e-6 * out["n_mapped"]
    assert np.corrcoef(out["alpha"], true_counts)[0, 1] > 0.95


def test_quant_eqclasses_matches_oracle(tmp_path):
    from salmon_b200.quant import quant_eqclasses
    eq, proj, eff, uniq = synth_eq(seed=5, C=20000, M=6000, total_count=400000)
    names = [f"t{i}" for i in range(eq.n_txps)]
    path = str(tmp_path / "eq_classes.txt.gz")
    _capi.write_eq_classes(path, names, eq.off, eq.tids, eq.counts, eq.weights)
    with gzip.open(path, "at") as f:
        for i in range(eq.n_txps):
            f.write(f"{names[i]}\t{eff[i]:.17g}\n")
    got = quant_eqclasses(path, out_dir=str(tmp_path / "o"))
    f = _capi.read_eq_classes(path)
    eq2 = _capi.EqClasses(f["n_txps"], f["off"], f["tids"], f["weights"], f["counts"])
    p = _capi.default_params(eq_class_mode=1, init_uniform=1)
    ref, rst = O.em_optimize(eq2, np.zeros(eq.n_txps), f["eff_len"], np.zeros(eq.n_txps, np.uint64), p)
    assert got["em_stats"].iters == rst.iters
    np.testing.assert_allclose(got["alpha"], ref, rtol=1e-9, atol=1e-9)
    assert np.array_equal(f["eff_len"], eff)
    assert len((tmp_path / "o" / "quant.sf").read_text().splitlines()) == eq.n_txps + 1


def test_native_driver_and_cli_match_python_pipeline(tmp_path):
    """sb_quant_files (C++ reader thread + GPU thread) and the sb_salmon command line give the numbers of the Python
    mirror on the same files (uniform read length: same batches in the same order -> bit-identical alphas)."""
    import subprocess
    from salmon_b200.quant import quant_files
    txps, _ = synth_txome(seed=41, n_genes=120)
    left, right, truth = synth_reads(txps, seed=45, n=20000)
    names = [f"ENST{i:05d}" for i in range(len(txps))]
    idx = Index(txps, names=names)
    f1, f2 = str(tmp_path / "r_1.fq.gz"), str(tmp_path / "r_2.fq.gz")
    write_fastq(f1, left); write_fastq(f2, right)
    py = quant_files(idx, f1, f2, batch=8192, max_read_len=128)
    alpha, sm = _capi.quant_files_native(idx, f1, f2, out_dir=str(tmp_path / "nat"), batch=8192, max_read_len=128,
                                         dump_eq_weights=1, num_bootstraps=2, seed=5)
    assert sm["n_observed"] == 20000 and sm["n_mapped"] == py["n_mapped"] and sm["n_read_lengths"] == 1
    assert np.array_equal(alpha, py["alpha"])
    assert (tmp_path / "nat" / "aux_info" / "bootstrap" / "names.tsv.gz").exists()
    raw = gzip.open(tmp_path / "nat" / "aux_info" / "bootstrap" / "bootstraps.gz", "rb").read()
    assert len(raw) == 2 * len(txps) * 8
    # the command line: index from FASTA, quant from the files
    fa = tmp_path / "t.fa"
    with open(fa, "wb") as f:
        for nm, t in zip(names, txps):
            f.write(b">" + nm.encode() + b" x\n" + LETTERS[t].tobytes() + b"\n")
    exe = os.path.join(os.path.dirname(_capi.LIB_PATH), "sb_salmon")
    r = subprocess.run([exe, "index", "-t", str(fa), "-i", str(tmp_path / "idx"), "--no-clip", "--keepDuplicates"],
                       capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    r = subprocess.run([exe, "quant", "-i", str(tmp_path / "idx"), "-l", "IU", "-1", f1, "-2", f2, "-o", str(tmp_path / "cli"),
                        "--batch", "8192", "--maxReadLen", "128", "--dumpEq"], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    a = (tmp_path / "nat" / "quant.sf").read_text().splitlines()
    b = (tmp_path / "cli" / "quant.sf").read_text().splitlines()
    assert a == b and a[1].startswith("ENST00000\t")
    assert (tmp_path / "cli" / "aux_info" / "meta_info.json").exists()
    # variable read lengths through the native driver: same totals as the Python mirror
    n = 6000
    rng = np.random.default_rng(4)
    ll = np.full(n, 100); lr = np.full(n, 100)
    trim = rng.random(n) < 0.5
    ll[trim] = rng.integers(25, 101, size=int(trim.sum())); lr[trim] = rng.integers(25, 101, size=int(trim.sum()))
    g1, g2 = str(tmp_path / "v_1.fq"), str(tmp_path / "v_2.fq")
    write_fastq(g1, left[:n], ll, gz=False); write_fastq(g2, right[:n], lr, gz=False)
    alpha2, sm2 = _capi.quant_files_native(idx, g1, g2, batch=4096, max_read_len=128)
    py2 = quant_files(idx, g1, g2, batch=4096, max_read_len=128)
    assert sm2["n_observed"] == n and sm2["n_read_lengths"] > 10 and sm2["n_too_short"] > 0
    assert abs(alpha2.sum() - sm2["n_mapped"]) < 1e-6 * sm2["n_mapped"]
    assert abs(int(sm2["n_mapped"]) - py2["n_mapped"]) <= 0.01 * py2["n_mapped"]     # different batch composition
    assert np.corrcoef(alpha2, py2["alpha"])[0, 1] > 0.999


def test_decoys_are_dropped_and_run_metadata_is_written(tmp_path):
    """with a decoy-aware index the optimiser, quant.sf and eq_classes.txt.gz see the real targets
    only (readExp.dropDecoyTranscripts(), SalmonQuantify.cpp:2479); plus the run-metadata files of the drop-in contract
    (GZipWriter.cpp:294-640, ReadExperiment.inl:219-350, MappingPipelineStages.cpp:164-173)."""
    import json
    import subprocess
    from salmon_b200.quant import quant_files
    txps, _ = synth_txome(seed=47, n_genes=100)
    n_real = len(txps) - 25                      # the last 25 references act as decoys
    left, right, truth = synth_reads(txps, seed=48, n=15000)
    names = [f"ENST{i:05d}" for i in range(len(txps))]
    idx = Index(txps, names=names, first_decoy=n_real)
    f1, f2 = str(tmp_path / "d_1.fq"), str(tmp_path / "d_2.fq")
    write_fastq(f1, left, gz=False); write_fastq(f2, right, gz=False)
    out = tmp_path / "out"
    alpha, sm = _capi.quant_files_native(idx, f1, f2, out_dir=str(out), batch=8192, max_read_len=128, dump_eq_weights=1, seed=3)
    assert np.all(alpha[n_real:] == 0.0) and abs(alpha.sum() - sm["n_mapped"]) < 1e-6 * sm["n_mapped"]
    rows = (out / "quant.sf").read_text().splitlines()
    assert len(rows) == n_real + 1 and rows[-1].split("\t")[0] == names[n_real - 1]
    eqf = _capi.read_eq_classes(str(out / "aux_info" / "eq_classes.txt.gz"))
    assert eqf["n_txps"] == n_real and int(eqf["tids"].max()) < n_real
    py = quant_files(idx, f1, f2, batch=8192, max_read_len=128)
    assert len(py["alpha"]) == n_real and np.array_equal(py["alpha"], alpha[:n_real])
    # the optimiser's uniform prior is totalWeight / M with M = real targets (CollapsedEMOptimizer.cpp:803): the oracle on
    # the same classes with n_real transcripts reproduces alpha
    eq = _capi.EqClasses(n_real, py["classes"]["off"], py["classes"]["tids"], py["classes"]["weights"], py["classes"]["counts"])
    ref, _ = O.em_optimize(eq, py["projected_counts"], py["eff_len"], py["unique_counts"], _capi.default_params())
    np.testing.assert_allclose(py["alpha"], ref, rtol=1e-9, atol=1e-9)
    # ---- run metadata
    meta = json.loads((out / "aux_info" / "meta_info.json").read_text())
    assert meta["num_valid_targets"] == n_real and meta["num_decoy_targets"] == 25
    assert meta["num_processed"] == 15000 and meta["num_mapped"] == sm["n_mapped"] and meta["opt_type"] == "vb"
    assert meta["samp_type"] == "none" and meta["mapping_type"] == "mapping" and meta["serialized_eq_classes"] is True
    assert abs(meta["percent_mapped"] - 100.0 * sm["n_mapped"] / 15000) < 1e-4
    fld = np.frombuffer(gzip.open(out / "aux_info" / "fld.gz", "rb").read(), dtype=np.int32)
    assert fld.shape[0] == 1001 and int(fld.sum()) == 10000 and abs(meta["frag_length_mean"] - 250) < 25
    assert abs(float((np.arange(1001) * fld).sum()) / 10000 - meta["frag_length_mean"]) < 5
    pmf = np.array((out / "libParams" / "flenDist.txt").read_text().split(), dtype=float)
    assert pmf.shape[0] == 1001 and abs(pmf.sum() - 1.0) < 1e-3
    lfc = json.loads((out / "lib_format_counts.json").read_text())
    assert lfc["expected_format"] == "IU" and lfc["num_assigned_fragments"] == sm["n_mapped"]
    amb = (out / "aux_info" / "ambig_info.tsv").read_text().splitlines()
    assert amb[0] == "UniqueCount\tAmbigCount" and len(amb) == n_real + 1
    assert sum(int(x.split("\t")[0]) + int(x.split("\t")[1]) for x in amb[1:]) >= sm["n_mapped"]
    # cmd_info.json comes from the command line
    ipath = tmp_path / "idx"; ipath.mkdir()
    idx.save(str(ipath / "sb_index.bin"))
    exe = os.path.join(os.path.dirname(_capi.LIB_PATH), "sb_salmon")
    r = subprocess.run([exe, "quant", "-i", str(ipath), "-l", "IU", "-1", f1, "-2", f2, "-o", str(tmp_path / "cli"),
                        "--batch", "8192", "--maxReadLen", "128"], capture_output=True, text=True)   # (--seed would also seed the mapping's FLD draws)
    assert r.returncode == 0, r.stderr
    ci = json.loads((tmp_path / "cli" / "cmd_info.json").read_text())
    assert ci["libType"] == "IU" and ci["mates1"] == f1 and ci["batch"] == "8192"
    assert (tmp_path / "cli" / "quant.sf").read_text() == (out / "quant.sf").read_text()


def test_cli_eqclasses_bootstraps(tmp_path):
    """`sb_salmon quant -e` through sb_quant_eqclasses: quant.sf + bootstraps.gz, equal to the Python mirror"""
    import subprocess
    from salmon_b200.quant import quant_eqclasses
    eq, proj, eff, uniq = synth_eq(seed=15, C=6000, M=1500, total_count=200000)
    names = [f"t{i}" for i in range(eq.n_txps)]
    path = str(tmp_path / "eq_classes.txt.gz")
    _capi.write_eq_classes(path, names, eq.off, eq.tids, eq.counts, eq.weights)
    exe = os.path.join(os.path.dirname(_capi.LIB_PATH), "sb_salmon")
    r = subprocess.run([exe, "quant", "-e", path, "-o", str(tmp_path / "o"), "--numBootstraps", "4", "--seed", "11"],
                       capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    py = quant_eqclasses(path, num_bootstraps=4, seed=11)
    raw = gzip.open(tmp_path / "o" / "aux_info" / "bootstrap" / "bootstraps.gz", "rb").read()
    boots = np.frombuffer(raw, dtype=np.float64).reshape(4, eq.n_txps)
    assert np.array_equal(boots, py["bootstraps"])
    rows = (tmp_path / "o" / "quant.sf").read_text().splitlines()
    got = np.array([float(x.split("\t")[4]) for x in rows[1:]])
    np.testing.assert_allclose(got, py["alpha"], rtol=0, atol=6e-4)     # %.3f


def test_cli_single_end_and_stranded_libraries(tmp_path):
    """`sb_salmon quant -l SF -r reads.fq` and `-l ISF -1 -2`: quant.sf, lib_format_counts.json with the observed formats"""
    import json
    import subprocess
    txps, _ = synth_txome(seed=49, n_genes=100)
    left, right, truth = synth_reads(txps, seed=50, n=12000)
    names = [f"ENST{i:05d}" for i in range(len(txps))]
    idx = Index(txps, names=names)
    ipath = tmp_path / "idx"; ipath.mkdir()
    idx.save(str(ipath / "sb_index.bin"))
    f1, f2 = str(tmp_path / "s_1.fq"), str(tmp_path / "s_2.fq")
    write_fastq(f1, left, gz=False); write_fastq(f2, right, gz=False)
    exe = os.path.join(os.path.dirname(_capi.LIB_PATH), "sb_salmon")
    runs = {}
    for tag, args in (("U", ["-l", "U", "-r", f1]), ("SF", ["-l", "SF", "-r", f1]), ("SR", ["-l", "SR", "-r", f1]),
                      ("IU", ["-l", "IU", "-1", f1, "-2", f2]), ("ISF", ["-l", "ISF", "-1", f1, "-2", f2])):
        out = tmp_path / tag
        r = subprocess.run([exe, "quant", "-i", str(ipath)] + args + ["-o", str(out), "--batch", "4096", "--maxReadLen", "128"],
                           capture_output=True, text=True)
        assert r.returncode == 0, r.stderr
        lfc = json.loads((out / "lib_format_counts.json").read_text())
        meta = json.loads((out / "aux_info" / "meta_info.json").read_text())
        assert lfc["expected_format"] == tag and meta["library_types"] == [tag]
        rows = (out / "quant.sf").read_text().splitlines()
        assert len(rows) == len(txps) + 1
        runs[tag] = (lfc, meta["num_mapped"])
    assert runs["U"][1] == runs["SF"][1] + runs["SR"][1] or runs["U"][1] <= runs["SF"][1] + runs["SR"][1]
    assert runs["SF"][0]["SR"] == 0 and runs["SR"][0]["SF"] == 0 and runs["U"][0]["ISF"] == 0
    assert runs["ISF"][0]["ISR"] == 0 and runs["IU"][0]["ISF"] + runs["IU"][0]["ISR"] > 0
    assert 0 < runs["ISF"][1] < runs["IU"][1]
    r = subprocess.run([exe, "quant", "-i", str(ipath), "-l", "SF", "-1", f1, "-2", f2, "-o", str(tmp_path / "bad")], capture_output=True, text=True)
    assert r.returncode != 0 and "does not fit the input" in r.stderr
