"""BASELINE.json configs[0]: the reference's bundled sample_data (15 transcripts, 10 000 simulated 2x50 bp pairs) as a
committed fixture (tests/golden/sample_data/, made by tests/golden/make_sample_fixture.py).  The read names carry the truth
(`@<i>:<transcript>:<position>:<fragment length>`), which is what the reference's own test data offers for this path
(SURVEY.md 8c): the oracle's mapping core is checked against it on the CPU, the CUDA path against the oracle and the
truth through the command line on the GPU."""
import gzip
import os
import subprocess

import numpy as np
import pytest

from salmon_b200 import _capi

HERE = os.path.dirname(os.path.abspath(__file__))
FIX = os.path.join(HERE, "golden", "sample_data")
CODE = {c: i for i, c in enumerate("ACGT")}


def encode(seq):
    return np.array([CODE.get(c, 4) for c in seq.upper()], dtype=np.uint8)


def load_fixture():
    names, seqs = [], []
    for ln in gzip.open(os.path.join(FIX, "transcripts.fasta.gz"), "rt"):
        ln = ln.strip()
        if ln.startswith(">"):
            names.append(ln[1:].split()[0]); seqs.append([])
        elif ln:
            seqs[-1].append(ln)
    txps = [encode("".join(s)) for s in seqs]
    reads, truth = [], []
    for fn in ("reads_1.fastq.gz", "reads_2.fastq.gz"):
        lines = gzip.open(os.path.join(FIX, fn), "rt").read().split("\n")
        reads.append(np.stack([encode(s) for s in lines[1::4] if s]))
        if not truth:
            truth = [h[1:].split(":") for h in lines[0::4] if h]
    tid = np.array([names.index(t[1]) for t in truth])
    flen = np.array([int(t[3]) for t in truth])
    return names, txps, reads[0], reads[1], tid, flen


def test_fixture_shape():
    names, txps, left, right, tid, flen = load_fixture()
    assert len(names) == 15 and left.shape == right.shape == (10000, 50)
    assert sum(len(t) for t in txps) > 20000
    assert tid.min() >= 0 and flen.min() >= 50


def test_oracle_maps_sample_reads_to_their_origin(oracle):
    """MAPSPEC (the oracle's mapping core) on the reference's own simulated reads: nearly every pair maps, the true
    transcript is among the reported alignments, and when the alignment is unique the fragment length is the true one."""
    names, txps, left, right, tid, flen = load_fixture()
    p = oracle.map_params()
    m = oracle.map_reads(oracle.MapIndex(txps), p, left, right, 0)
    mapped = m["n_aln"] > 0
    assert mapped.mean() > 0.97, mapped.mean()
    cap = p.max_read_occ
    sel = np.arange(cap)[None, :] < m["n_aln"][:, None]
    hit = ((m["tid"] == tid[:, None]) & sel).any(axis=1)
    assert hit[mapped].mean() > 0.995, hit[mapped].mean()
    uniq = m["n_aln"] == 1
    ok_len = (np.abs(m["flen"][uniq, 0] - flen[uniq]) <= 2).mean()
    assert ok_len > 0.98, ok_len


@pytest.mark.gpu
def test_sample_data_through_the_command_line(oracle, tmp_path):
    """`sb_salmon index` + `sb_salmon quant` on the fixture files (gzip FASTQ, FASTA): classes bit-exact against the
    oracle mapping the same reads in one batch, NumReads close to the simulated truth."""
    names, txps, left, right, tid, flen = load_fixture()
    exe = os.path.join(os.path.dirname(_capi.LIB_PATH), "sb_salmon")
    idx, out = str(tmp_path / "idx"), str(tmp_path / "out")
    r = subprocess.run([exe, "index", "-t", os.path.join(FIX, "transcripts.fasta.gz"), "-i", idx, "--no-clip", "--keepDuplicates"],
                       capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    r = subprocess.run([exe, "quant", "-i", idx, "-l", "IU", "-1", os.path.join(FIX, "reads_1.fastq.gz"), "-2",
                        os.path.join(FIX, "reads_2.fastq.gz"), "-o", out, "--dumpEqWeights", "--batch", "16384", "--maxReadLen", "64"],
                       capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    rows = [ln.split("\t") for ln in open(os.path.join(out, "quant.sf")).read().splitlines()[1:]]
    assert [x[0] for x in rows] == names and [int(x[1]) for x in rows] == [len(t) for t in txps]
    num_reads = np.array([float(x[4]) for x in rows])
    truth = np.bincount(tid, minlength=len(names)).astype(float)
    assert abs(num_reads.sum() - 10000) < 300                       # nearly every simulated pair is assigned
    assert np.corrcoef(num_reads, truth)[0, 1] > 0.995
    assert np.abs(num_reads - truth).sum() / truth.sum() < 0.08     # total mis-assignment across 15 transcripts
    assert abs(sum(float(x[3]) for x in rows) - 1e6) < 1.0          # TPM column
    # the classes the run dumped == the oracle's classes for the same reads (one batch: same frozen online state)
    p = oracle.map_params()
    m = oracle.map_reads(oracle.MapIndex(txps), p, left, right, 0)
    e = oracle.eq_aggregate(m, p.max_read_occ, True)
    f = _capi.read_eq_classes(os.path.join(out, "aux_info", "eq_classes.txt.gz"))
    got = sorted((tuple(f["tids"][int(f["off"][c]):int(f["off"][c + 1])].tolist()), int(f["counts"][c])) for c in range(len(f["counts"])))
    eo = e["off"].astype(np.int64)
    want = sorted((tuple(e["tids"][eo[c]:eo[c + 1]].tolist()), int(e["counts"][c])) for c in range(len(e["counts"])))
    assert sum(c for _, c in got) == sum(c for _, c in want) == int((m["n_aln"] > 0).sum())
    assert got == want


def test_product_host_logic_matches_oracle_on_sample_reads(oracle):
    """map_core.h (the per-read logic the CUDA kernels share, compiled for the host) against the oracle on the reference's
    simulated 2x50 bp reads: alignments, scores, probabilities, labels and weights bit-exact."""
    from test_map_host import run_both
    names, txps, left, right, tid, flen = load_fixture()
    run_both(oracle, txps, left[:4000], right[:4000])
    run_both(oracle, txps, left[4000:6000], right[4000:6000], frag_counter=6_000_000)


@pytest.mark.gpu
def test_gpu_alignments_against_the_simulated_truth():
    """VERDICT r1 next #3d: the truth carried by the reference's simulated reads, checked on the CUDA PATH itself (not via
    the oracle): every pair maps, the true transcript is among the alignments of every pair, and a uniquely mapped
    pair's fragment length and leftmost position are the simulated ones."""
    from salmon_b200._capi import Index, MapContext, map_default_params
    names, txps, left, right, tid, flen = load_fixture()
    p = map_default_params()
    ctx = MapContext(Index(txps), p, batch_cap=16384, max_read_len=left.shape[1])
    ctx.map_batch(left, right)
    a = ctx.last_alignments()
    ctx.close()
    na = a["n_aln"]
    assert (na > 0).all()                                            # 100 % of the simulated pairs map
    cap = p.max_read_occ
    valid = np.arange(cap)[None, :] < na[:, None]
    has_truth = ((a["tid"] == tid[:, None]) & valid).any(axis=1)
    assert has_truth.all()                                           # the true transcript is always among the alignments
    uniq = na == 1
    assert uniq.sum() > 1000
    assert np.array_equal(a["tid"][uniq, 0], tid[uniq])
    assert np.array_equal(a["flen"][uniq, 0], flen[uniq])            # exact fragment length
    # the alignment on the true transcript carries the simulated fragment length for multi-mappers too
    j = np.argmax((a["tid"] == tid[:, None]) & valid, axis=1)
    paired = ((a["flags"][np.arange(len(na)), j] >> 2) & 3) == 0       # mate status 0 = properly paired
    fl_true = a["flen"][np.arange(len(na)), j]
    assert (fl_true[paired] == flen[paired]).mean() > 0.999
