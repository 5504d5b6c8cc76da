"""Input seam (host code, no GPU): FASTQ/FASTA reader, --eqclasses reader, bootstraps.gz writer, transcript FASTA
loader.  The checker is an independent pure-Python parse of the same files; formats follow the reference
(FQFeeder records as consumed in src/quant/SalmonQuantify.cpp:1118-1141; src/util/SalmonUtils.cpp:1024-1122;
src/output/GZipWriter.cpp:64-168,765-789; src/index/BuildSalmonIndex.cpp:72-124)."""
import gzip
import os

import numpy as np
import pytest

from salmon_b200 import _capi

CODE = {c: i for i, c in enumerate("ACGT")}
CODE.update({c.lower(): i for c, i in list(CODE.items())})


def encode(seq):
    return np.array([CODE.get(c, 4) for c in seq], dtype=np.uint8)


def rand_seq(rng, n, alphabet="ACGT"):
    return "".join(rng.choice(list(alphabet), size=n))


def write_text(path, text, gz):
    if gz:
        with gzip.open(path, "wt") as f:
            f.write(text)
    else:
        with open(path, "w") as f:
            f.write(text)


def make_reads(rng, n, lo=30, hi=120):
    out = []
    for i in range(n):
        L = int(rng.integers(lo, hi + 1))
        s = rand_seq(rng, L, "ACGTNacgt" if i % 7 == 0 else "ACGT")
        out.append(s)
    return out


def fastq_text(reads, tag, crlf=False, final_newline=True):
    nl = "\r\n" if crlf else "\n"
    recs = [f"@r{i}/{tag} extra{nl}{s}{nl}+{nl}{'I' * len(s)}" for i, s in enumerate(reads)]
    t = nl.join(recs)
    return t + (nl if final_newline else "")


@pytest.mark.parametrize("gz", [False, True])
@pytest.mark.parametrize("crlf", [False, True])
def test_fastq_pairs_match_python_parse(tmp_path, gz, crlf):
    rng = np.random.default_rng(5)
    n = 5000
    r1, r2 = make_reads(rng, n), make_reads(rng, n)
    f1 = tmp_path / ("a_1.fq.gz" if gz else "a_1.fq")
    f2 = tmp_path / ("a_2.fq.gz" if gz else "a_2.fq")
    write_text(f1, fastq_text(r1, 1, crlf), gz)
    write_text(f2, fastq_text(r2, 2, crlf, final_newline=False), gz)
    got_l, got_r = [], []
    with _capi.ReadFiles(str(f1), str(f2), n_threads=3) as rf:
        while True:
            k, left, right, ll, lr = rf.next_batch(777, 128)
            if k == 0:
                break
            for i in range(k):
                got_l.append(left[i, :ll[i]].copy())
                got_r.append(right[i, :lr[i]].copy())
                assert np.all(left[i, ll[i]:] == 4) and np.all(right[i, lr[i]:] == 4)   # padding
    assert len(got_l) == n and len(got_r) == n
    for i in range(n):
        assert np.array_equal(got_l[i], encode(r1[i])), i
        assert np.array_equal(got_r[i], encode(r2[i])), i


def test_multiple_files_single_end_and_fasta(tmp_path):
    rng = np.random.default_rng(6)
    a, b = make_reads(rng, 300), make_reads(rng, 411)
    fa = tmp_path / "a.fq"
    fb = tmp_path / "b.fa.gz"
    write_text(fa, fastq_text(a, 1), False)
    write_text(fb, "".join(f">s{i} d\n{s}\n" for i, s in enumerate(b)), True)
    rf = _capi.ReadFiles([str(fa), str(fb)], None, n_threads=2)
    seqs = []
    while True:
        k, left, right, ll, lr = rf.next_batch(128, 120)
        if k == 0:
            break
        assert right is None and lr is None
        seqs += [left[i, :ll[i]].copy() for i in range(k)]
    rf.close()
    want = a + b
    assert len(seqs) == len(want)
    assert all(np.array_equal(x, encode(y)) for x, y in zip(seqs, want))


def test_large_block_boundaries(tmp_path):
    # > 8 MiB per file so that records straddle the splitter's chunk boundaries
    rng = np.random.default_rng(7)
    n = 60000
    base = rng.integers(0, 4, size=(n, 100), dtype=np.uint8)
    lut = np.frombuffer(b"ACGT", dtype=np.uint8)
    seqs = [lut[row].tobytes().decode() for row in base]
    f1, f2 = tmp_path / "l_1.fq", tmp_path / "l_2.fq"
    write_text(f1, fastq_text(seqs, 1), False)
    write_text(f2, fastq_text(seqs[::-1], 2), False)
    assert os.path.getsize(f1) > (8 << 20)
    tot = 0
    with _capi.ReadFiles(str(f1), str(f2), n_threads=4) as rf:
        while True:
            k, left, right, ll, lr = rf.next_batch(16384, 100)
            if k == 0:
                break
            assert np.all(ll == 100) and np.all(lr == 100)
            assert np.array_equal(left, base[tot:tot + k])
            assert np.array_equal(right, base[::-1][tot:tot + k])
            tot += k
    assert tot == n


def test_reader_errors(tmp_path):
    rng = np.random.default_rng(8)
    a = make_reads(rng, 50)
    f1, f2 = tmp_path / "e_1.fq", tmp_path / "e_2.fq"
    write_text(f1, fastq_text(a, 1), False)
    write_text(f2, fastq_text(a[:40], 2), False)
    rf = _capi.ReadFiles(str(f1), str(f2))
    with pytest.raises(_capi.SalmonB200Error, match="different numbers of records"):
        while rf.next_batch(64, 128)[0]:
            pass
    rf.close()
    # a read longer than the stride
    rf = _capi.ReadFiles(str(f1), None)
    with pytest.raises(_capi.SalmonB200Error, match="exceeds the buffer stride"):
        rf.next_batch(64, 20)
    rf.close()
    # quality / sequence length mismatch
    bad = tmp_path / "bad.fq"
    bad.write_text("@x\nACGT\n+\nIII\n")
    rf = _capi.ReadFiles(str(bad), None)
    with pytest.raises(_capi.SalmonB200Error, match="quality and sequence lengths differ"):
        rf.next_batch(4, 16)
    rf.close()
    # truncated record
    bad.write_text("@x\nACGT\n+\nIIII\n@y\nAC")
    rf = _capi.ReadFiles(str(bad), None)
    with pytest.raises(_capi.SalmonB200Error, match="truncated"):
        while rf.next_batch(4, 16)[0]:
            pass
    rf.close()
    # missing file
    rf = _capi.ReadFiles(str(tmp_path / "nope.fq"), None)
    with pytest.raises(_capi.SalmonB200Error, match="cannot open"):
        rf.next_batch(4, 16)
    rf.close()


def test_eq_classes_round_trip(tmp_path):
    rng = np.random.default_rng(9)
    M, Cn = 40, 200
    names = [f"tx{i}|g" for i in range(M)]
    sizes = rng.integers(1, 6, size=Cn)
    off = np.concatenate([[0], np.cumsum(sizes)]).astype(np.uint64)
    tids = np.concatenate([np.sort(rng.choice(M, size=k, replace=False)) for k in sizes]).astype(np.uint32)
    w = rng.random(int(off[-1]))
    counts = rng.integers(1, 1000, size=Cn).astype(np.uint64)
    for gz in (False, True):
        path = str(tmp_path / ("eq.txt.gz" if gz else "eq.txt"))
        _capi.write_eq_classes(path, names, off, tids, counts, w)
        got = _capi.read_eq_classes(path)
        assert got["names"] == names and got["has_weights"] and got["n_missing_eff_len"] == M
        assert np.array_equal(got["off"], off) and np.array_equal(got["tids"], tids)
        assert np.array_equal(got["counts"], counts)
        np.testing.assert_allclose(got["weights"], w, rtol=1e-5)   # the reference prints 6 significant digits
        assert np.all(got["eff_len"] == 100.0)                       # SalmonUtils.cpp:1109-1116
    # the trailer with effective lengths, as the reader expects it (readEquivCounts :1095-1107)
    path = str(tmp_path / "eq_trailer.txt")
    _capi.write_eq_classes(path, names, off, tids, counts, w)
    eff = rng.random(M) * 1000 + 1
    with open(path, "a") as f:
        for i in rng.permutation(M)[:30]:
            f.write(f"{names[i]}\t{eff[i]:.17g}\n")
            eff[i] = -eff[i]
    got = _capi.read_eq_classes(path)
    assert got["n_missing_eff_len"] == 10
    given = eff < 0
    assert np.array_equal(got["eff_len"][given], -eff[given]) and np.all(got["eff_len"][~given] == 100.0)
    # without weights (--dumpEq only): classes collapse by transcript set
    path = str(tmp_path / "eq_now.txt")
    _capi.write_eq_classes(path, names, off, tids, counts, None)
    got = _capi.read_eq_classes(path)
    assert not got["has_weights"] and got["weights"] is None
    assert int(got["counts"].sum()) == int(counts.sum())
    # malformed
    bad = tmp_path / "bad.txt"
    bad.write_text("2\n1\na\nb\n2\t0\t5\t7\n")
    with pytest.raises(_capi.SalmonB200Error):
        _capi.read_eq_classes(str(bad))


def test_bootstrap_writer_format(tmp_path):
    rng = np.random.default_rng(10)
    path = str(tmp_path / "bootstraps.gz")
    w = _capi.BootstrapWriter(path)
    samples = [rng.random(123) * 50 for _ in range(7)]
    for smp in samples:
        w.write(smp)
    assert w.close() == 7
    raw = gzip.open(path, "rb").read()
    got = np.frombuffer(raw, dtype=np.float64).reshape(7, 123)   # tximport / fishpond read it exactly like this
    assert np.array_equal(got, np.stack(samples))
    # samples of several deflate slices (the writer compresses 128 KiB slices in parallel), sparse like real counts;
    # one gzip member with a valid CRC / length (gzip.open checks both), also for an empty file
    import subprocess
    path2 = str(tmp_path / "big.gz")
    w = _capi.BootstrapWriter(path2)
    big = [np.where(rng.random(300_001) < 0.3, rng.random(300_001) * 1000, 0.0) for _ in range(3)]
    for smp in big:
        w.write(smp)
    assert w.close() == 3
    raw = gzip.open(path2, "rb").read()
    assert np.array_equal(np.frombuffer(raw, dtype=np.float64).reshape(3, 300_001), np.stack(big))
    assert subprocess.run(["gzip", "-t", path2]).returncode == 0
    assert raw[:0] == b"" and open(path2, "rb").read(3) == b"\x1f\x8b\x08"
    path3 = str(tmp_path / "none.gz")
    w = _capi.BootstrapWriter(path3)
    assert w.close() == 0 and gzip.open(path3, "rb").read() == b"" and subprocess.run(["gzip", "-t", path3]).returncode == 0


def test_txome_fasta_options(tmp_path):
    rng = np.random.default_rng(11)
    s0 = rand_seq(rng, 300)
    s1 = rand_seq(rng, 200) + "A" * 25           # poly-A tail: clipped
    s2 = s0                                        # duplicate of tx0
    s3 = rand_seq(rng, 20)                         # shorter than k
    s4 = rand_seq(rng, 150) + "NNNN" + rand_seq(rng, 50) + "A" * 8   # short A run stays
    d0 = rand_seq(rng, 500) + "A" * 30             # decoy: never clipped
    def wrap(s):
        return "\n".join(s[i:i + 60] for i in range(0, len(s), 60))
    fa = tmp_path / "t.fa.gz"
    write_text(fa, "".join(f">{n} desc\n{wrap(s)}\n" for n, s in [
        ("ENST0|ENSG0|x", s0), ("ENST1|ENSG1|x", s1), ("ENST2|ENSG0|x", s2), ("ENST3|g", s3), ("ENST4|g", s4),
        ("chr1", d0)]), True)
    dec = tmp_path / "decoys.txt"
    dec.write_text("chr1\n")
    t = _capi.read_txome_fasta(str(fa), k=31, gencode=True, decoys=str(dec))
    assert t["names"] == ["ENST0", "ENST1", "ENST3", "ENST4", "chr1"]
    assert t["first_decoy"] == 4 and t["n_duplicates_removed"] == 1 and t["n_clipped"] == 1 and t["n_short"] == 1
    assert list(t["complete_len"]) == [300, 225, 20, len(s4), 530]
    assert np.array_equal(t["seqs"][0], encode(s0))
    assert np.array_equal(t["seqs"][1], encode(s1[:200].rstrip("A")))
    assert np.array_equal(t["seqs"][3], encode(s4)) and int((t["seqs"][3] == 4).sum()) == 4
    assert np.array_equal(t["seqs"][4], encode(d0))
    # --keepDuplicates --no-clip, names up to the first white space
    t2 = _capi.read_txome_fasta(str(fa), k=31, keep_duplicates=True, no_clip=True)
    assert t2["names"][0] == "ENST0|ENSG0|x" and len(t2["names"]) == 6 and t2["first_decoy"] == 6
    assert np.array_equal(t2["seqs"][1], encode(s1))
    # a decoy in the middle is an error
    dec.write_text("ENST1|ENSG1|x\n")
    with pytest.raises(_capi.SalmonB200Error, match="decoys must come last"):
        _capi.read_txome_fasta(str(fa), k=31, decoys=str(dec))


def test_long_decoys_are_split_into_overlapping_pieces(tmp_path):
    """chromosome-sized decoys exceed the index's per-reference limit (2^21 - 1 bases): they are stored as pieces of 2 Mb
    that overlap by 1024 bases, so every window of up to 1025 bases of the chromosome lies inside one piece"""
    rng = np.random.default_rng(13)
    tx = rand_seq(rng, 400)
    chrom = np.frombuffer(b"ACGT", dtype=np.uint8)[rng.integers(0, 4, 4_500_000)].tobytes().decode()
    small = rand_seq(rng, 1000)
    fa = tmp_path / "g.fa"
    with open(fa, "w") as f:
        f.write(f">tx0\n{tx}\n>chrBig\n")
        for i in range(0, len(chrom), 80):
            f.write(chrom[i:i + 80] + "\n")
        f.write(f">chrSmall\n{small}\n")
    dec = tmp_path / "d.txt"
    dec.write_text("chrBig\nchrSmall\n")
    t = _capi.read_txome_fasta(str(fa), k=31, decoys=str(dec))
    assert t["names"] == ["tx0", "chrBig:0", "chrBig:1", "chrBig:2", "chrSmall"] and t["first_decoy"] == 1
    lens = [len(x) for x in t["seqs"]]
    assert max(lens) < (1 << 21)
    full = encode(chrom)
    starts = [0, 2_000_000 - 1024, 2 * (2_000_000 - 1024)]
    for piece, a in zip(t["seqs"][1:4], starts):
        assert np.array_equal(piece, full[a:a + len(piece)])
    assert starts[2] + lens[3] == len(chrom) and lens[1] == lens[2] == 2_000_000
    assert list(t["complete_len"][1:4]) == lens[1:4]
    # the pieces build (the limit applies per reference)
    idx = _capi.Index(t["seqs"], names=t["names"], first_decoy=t["first_decoy"])
    assert idx.n_txps == 5


def test_index_save_load_round_trip(tmp_path):
    # host-side only: build (sb_index_build never touches the device), save, load, compare every array
    import ctypes as C
    rng = np.random.default_rng(12)
    fa = tmp_path / "t.fa"
    seqs = [rand_seq(rng, int(rng.integers(40, 400))) for _ in range(30)]
    seqs[3] = seqs[3][:50] + "NN" + seqs[3][52:]
    fa.write_text("".join(f">t{i}|g{i // 3} x\n{s}\n" for i, s in enumerate(seqs)))
    ix = _capi.Index.from_fasta(str(fa), k=21, gencode=True)
    path = str(tmp_path / "sb_index.bin")
    ix.save(path)
    iy = _capi.Index.load(path)
    assert iy.n_txps == ix.n_txps == 30 and iy.k == 21
    ma, mb = ix.meta(), iy.meta()
    assert ma["names"] == mb["names"] == [f"t{i}" for i in range(30)]
    assert np.array_equal(ma["complete_len"], mb["complete_len"]) and mb["first_decoy"] == 30
    assert ix.info() == iy.info()
    ha, hb = ix.host_arrays(), iy.host_arrays()
    assert ha["table_capacity"] == hb["table_capacity"] and ha["n_postings"] == hb["n_postings"]

    def raw(ptr, nbytes):
        return bytes((C.c_char * nbytes).from_address(ptr))
    total = int(ix.tx_lengths().sum())
    assert raw(ha["tx_off"], 8 * 31) == raw(hb["tx_off"], 8 * 31)
    assert raw(ha["codes"], total) == raw(hb["codes"], total)
    assert raw(ha["table"], 16 * ha["table_capacity"]) == raw(hb["table"], 16 * hb["table_capacity"])
    assert raw(ha["postings"], 8 * ha["n_postings"]) == raw(hb["postings"], 8 * hb["n_postings"])
    # a truncated file is rejected
    data = open(path, "rb").read()
    open(path, "wb").write(data[: len(data) // 2])
    with pytest.raises(_capi.SalmonB200Error, match="truncated or corrupt"):
        _capi.Index.load(path)
    open(path, "wb").write(b"not an index at all, just text padding to be long enough for a header " * 2)
    with pytest.raises(_capi.SalmonB200Error, match="bad header"):
        _capi.Index.load(path)


def test_cli_index_and_no_gpu_failure(tmp_path):
    """sb_salmon (C++ front end over the C ABI): `index` is host-only and runs anywhere; `quant` must fail loudly
    without a CUDA device (no CPU fallback)."""
    import subprocess
    exe = os.path.join(os.path.dirname(_capi.LIB_PATH), "sb_salmon")
    assert os.path.exists(exe), "build it with `make` (or __graft_entry__.build())"
    rng = np.random.default_rng(13)
    fa = tmp_path / "t.fa"
    seqs = [rand_seq(rng, int(rng.integers(100, 600))) for _ in range(12)]
    fa.write_text("".join(f">tx{i}|gene{i // 2}|more\n{s}\n" for i, s in enumerate(seqs)) + ">chrD\n" + rand_seq(rng, 900) + "\n")
    (tmp_path / "decoys.txt").write_text("chrD\n")
    idir = tmp_path / "idx"
    r = subprocess.run([exe, "index", "-t", str(fa), "-i", str(idir), "-k", "25", "--gencode", "-d", str(tmp_path / "decoys.txt")],
                       capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    ix = _capi.Index.load(str(idir / "sb_index.bin"))
    m = ix.meta()
    assert m["k"] == 25 and m["n_txps"] == 13 and m["first_decoy"] == 12 and m["names"][0] == "tx0" and m["names"][-1] == "chrD"
    import json
    info = json.load(open(idir / "info.json"))
    assert info["k"] == 25 and info["num_references"] == 13 and info["first_decoy"] == 12
    # even k is rejected like the reference does (BuildSalmonIndex.cpp:206-211)
    r = subprocess.run([exe, "index", "-t", str(fa), "-i", str(idir), "-k", "30"], capture_output=True, text=True)
    assert r.returncode == 1 and "odd" in r.stderr
    # options outside the hot path are refused, not ignored
    r = subprocess.run([exe, "quant", "-i", str(idir), "-l", "IU", "-1", "a.fq", "-2", "b.fq", "-o", str(tmp_path / "o"), "--gcBias"],
                       capture_output=True, text=True)
    assert r.returncode == 1 and "outside the hot path" in r.stderr
    if _capi.load().sb_device_count() == 0:
        f1, f2 = tmp_path / "a_1.fq", tmp_path / "a_2.fq"
        f1.write_text("@r\n" + seqs[0][:60] + "\n+\n" + "I" * 60 + "\n")
        f2.write_text("@r\n" + seqs[0][100:160] + "\n+\n" + "I" * 60 + "\n")
        r = subprocess.run([exe, "quant", "-i", str(idir), "-l", "IU", "-1", str(f1), "-2", str(f2), "-o", str(tmp_path / "o")],
                           capture_output=True, text=True)
        assert r.returncode == 1 and "no CUDA device" in r.stderr


def test_peek_and_skip(tmp_path):
    rng = np.random.default_rng(14)
    n = 3000
    seqs = [rand_seq(rng, 76) for _ in range(n)]
    lens2 = np.full(n, 76)
    lens2[1500] = 50                                   # one shorter read in the second mate file
    f1, f2 = tmp_path / "p_1.fq", tmp_path / "p_2.fq"
    write_text(f1, fastq_text(seqs, 1), False)
    write_text(f2, fastq_text([s[:lens2[i]] for i, s in enumerate(seqs)], 2), True)
    with _capi.ReadFiles(str(f1), str(f2), n_threads=2) as rf:
        assert rf.peek(1000) == (1000, 76)             # uniform
        assert rf.peek(1000) == (1000, 76)             # peeking delivers nothing
        k, left, right, ll, lr = rf.next_batch(400, 76)
        assert k == 400 and np.array_equal(left[399], encode(seqs[399]))
        assert rf.skip(600) == 600
        assert rf.peek(1000) == (1000, 0)              # records 1000..1999 hold the 50-base read
        k, left, right, ll, lr = rf.next_batch(1000, 76)
        assert k == 1000 and lr[500] == 50 and np.array_equal(left[0], encode(seqs[1000]))
        assert np.array_equal(right[500, :50], encode(seqs[1500][:50])) and np.all(right[500, 50:] == 4)
        assert rf.peek(5000) == (1000, 76)             # only 1000 left
        assert rf.skip(5000) == 1000
        assert rf.peek(10) == (0, 0) and rf.next_batch(10, 76)[0] == 0


def _pairs_with_lengths(rng, n, uniform):
    r1, r2 = [], []
    for i in range(n):
        if uniform:
            a = b = 80
        else:
            a = int(rng.integers(10, 101)); b = a if rng.random() < 0.6 else int(rng.integers(10, 101))
        r1.append(rand_seq(rng, a, "ACGTN" if i % 11 == 0 else "ACGT")); r2.append(rand_seq(rng, b))
    return r1, r2


@pytest.mark.parametrize("uniform", [True, False])
def test_bucketed_batches(tmp_path, uniform):
    """sb_reads_bucketed (the reader + length grouping of sb_quant_files, with a callback in place of sb_map_batch):
    every pair is delivered exactly once, in a batch of its own length, at the shorter mate's length."""
    rng = np.random.default_rng(21 + uniform)
    n, batch, k = 7000, 1024, 31
    r1, r2 = _pairs_with_lengths(rng, n, uniform)
    f1, f2 = tmp_path / "b_1.fq", tmp_path / "b_2.fq.gz"
    write_text(f1, fastq_text(r1, 1), False); write_text(f2, fastq_text(r2, 2), True)
    got, sizes = [], []

    def take(left, right, L):
        assert left.shape == right.shape and left.shape[1] == L and 0 < left.shape[0] <= batch
        sizes.append((L, left.shape[0]))
        got.extend((L, left[i].tobytes(), right[i].tobytes()) for i in range(left.shape[0]))
    with _capi.ReadFiles(str(f1), str(f2), n_threads=3) as rf:
        st = rf.bucketed(take, min_len=k, batch=batch, max_read_len=100, threads=3)
    want = []
    short = trimmed = 0
    for a, b in zip(r1, r2):
        L = min(len(a), len(b))
        trimmed += len(a) != len(b)
        if L < k:
            short += 1
            continue
        want.append((L, encode(a[:L]).tobytes(), encode(b[:L]).tobytes()))
    assert st["n_observed"] == n and st["n_delivered"] == len(want) == len(got) and st["n_too_short"] == short
    assert st["n_batches"] == len(sizes) and st["n_read_lengths"] == len({L for L, _ in sizes})
    if uniform:
        assert got == want and sizes == [(80, 1024)] * 6 + [(80, n - 6 * 1024)]     # order kept, no regrouping
        assert st["n_trimmed_mates"] == 0
    else:
        assert sorted(got) == sorted(want) and st["n_trimmed_mates"] == trimmed
        # inside one length the stream order is kept
        for L in {L for L, _ in sizes}:
            assert [g for g in got if g[0] == L] == [w for w in want if w[0] == L]


def test_bucketed_shards_and_abort(tmp_path):
    rng = np.random.default_rng(23)
    n, batch = 5000, 512
    r1, r2 = _pairs_with_lengths(rng, n, True)
    f1, f2 = tmp_path / "s_1.fq", tmp_path / "s_2.fq"
    write_text(f1, fastq_text(r1, 1), False); write_text(f2, fastq_text(r2, 2), False)
    all_rows = []
    for shard in range(3):
        rows = []
        with _capi.ReadFiles(str(f1), str(f2), n_threads=2) as rf:
            st = rf.bucketed(lambda l, r, L: rows.extend(l[i].tobytes() for i in range(l.shape[0])), batch=batch,
                             max_read_len=80, shard_index=shard, shard_count=3)
        # global batches g = shard, shard + 3, ... of 512 records each
        want = [encode(r1[i]).tobytes() for i in range(n) if (i // batch) % 3 == shard]
        assert rows == want and st["n_observed"] == n and st["n_delivered"] == len(want)
        all_rows += rows
    assert sorted(all_rows) == sorted(encode(s).tobytes() for s in r1)
    # a failing callback stops the stream and is reported
    calls = []
    with _capi.ReadFiles(str(f1), str(f2), n_threads=2) as rf:
        with pytest.raises(_capi.SalmonB200Error):
            rf.bucketed(lambda l, r, L: calls.append(1) or (-7 if len(calls) == 3 else 0), batch=batch, max_read_len=80)
    assert len(calls) == 3
    # single-end: right is None
    with _capi.ReadFiles(str(f1), None, n_threads=2) as rf:
        seen = []
        st = rf.bucketed(lambda l, r, L: seen.append((r is None, l.shape[0])), batch=2048, max_read_len=80)
    assert all(x for x, _ in seen) and sum(c for _, c in seen) == n


def test_parallel_splitter_adversarial_qualities(tmp_path, monkeypatch):
    """Plain files are memory-mapped and cut at record starts found from arbitrary offsets: quality lines that begin
    with '@' or '+' (legal Phred+33 characters) must not be taken for headers; same records as the serial splitter."""
    monkeypatch.setenv("SB_READS_SCANNERS", "4")
    rng = np.random.default_rng(31)
    n = 90000
    L = rng.integers(60, 121, size=n)
    qual_alphabet = np.frombuffer(b"@+IIIIFF#;", dtype=np.uint8)
    lut = np.frombuffer(b"ACGT", dtype=np.uint8)
    recs, seqs = [], []
    for i in range(n):
        sq = lut[rng.integers(0, 4, size=L[i])].tobytes()
        q = qual_alphabet[rng.integers(0, len(qual_alphabet), size=L[i])].tobytes()
        if i % 3 == 0:
            q = b"@" + q[1:]
        elif i % 3 == 1:
            q = b"+" + q[1:]
        seqs.append(sq)
        recs.append(b"@read%d\n%s\n+\n%s\n" % (i, sq, q))
    f1, f2 = tmp_path / "q_1.fq", tmp_path / "q_2.fq"
    f1.write_bytes(b"".join(recs)); f2.write_bytes(b"".join(recs[::-1]))
    assert os.path.getsize(f1) > (16 << 20)
    got1, got2 = [], []
    with _capi.ReadFiles(str(f1), str(f2), n_threads=4) as rf:
        while True:
            k, left, right, ll, lr = rf.next_batch(20000, 120)
            if k == 0:
                break
            got1 += [left[i, :ll[i]].tobytes() for i in range(k)]
            got2 += [right[i, :lr[i]].tobytes() for i in range(k)]
    want = [encode(s.decode()).tobytes() for s in seqs]
    assert got1 == want and got2 == want[::-1]
    # single-line FASTA reads through the same path
    fa = tmp_path / "q.fa"
    fa.write_bytes(b"".join(b">r%d\n%s\n" % (i, s) for i, s in enumerate(seqs)))
    got = []
    with _capi.ReadFiles(str(fa), None, n_threads=4) as rf:
        while True:
            k, left, right, ll, lr = rf.next_batch(30000, 120)
            if k == 0:
                break
            got += [left[i, :ll[i]].tobytes() for i in range(k)]
    assert got == want


def test_quant_files_argument_checks():
    """argument checks of the C++ driver happen before any device work (a sharded run needs the communicator id;
    posterior samples / class dumps need the whole table on one GPU)"""
    import ctypes as C
    idx = _capi.Index([encode(rand_seq(np.random.default_rng(1), 200))], k=31)
    for kw, msg in ((dict(shard_count=2), "communicator id"), (dict(shard_count=2, shard_index=2), "below shard_count"),
                    (dict(max_read_len=1000), "max_read_len"), (dict(num_bootstraps=2, num_gibbs=2), "not both")):
        with pytest.raises(_capi.SalmonB200Error, match=msg):
            _capi.quant_files_native(idx, "a.fq", "b.fq", **kw)


def test_quant_files_auto_library_type_arguments():
    """SB_LIB_AUTO_PAIRED / SB_LIB_AUTO_SINGLE (-l A) are resolved to their unstranded family before the mate-file check;
    without a GPU the call then fails loudly at sb_map_create (the library has no CPU path)"""
    idx = _capi.Index([encode(rand_seq(np.random.default_rng(2), 300))], k=31)
    auto_pe, auto_se = _capi.map_default_params(lib_type=6), _capi.map_default_params(lib_type=7)
    with pytest.raises(_capi.SalmonB200Error, match="needs both mate files"):
        _capi.quant_files_native(idx, "a.fq", None, map_params=auto_pe)
    with pytest.raises(_capi.SalmonB200Error, match="unmated reads only"):
        _capi.quant_files_native(idx, "a.fq", "b.fq", map_params=auto_se)
    import torch
    if not torch.cuda.is_available():
        with pytest.raises(_capi.SalmonB200Error, match="no CUDA device"):
            _capi.quant_files_native(idx, "a.fq", "b.fq", map_params=auto_pe)
    # the context itself takes the six concrete types only
    with pytest.raises(_capi.SalmonB200Error):
        _capi.MapContext(idx, _capi.map_default_params(lib_type=6), device=0, batch_cap=1024, max_read_len=128)
