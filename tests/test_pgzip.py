"""Parallel gzip inflate of the read files (salmon_b200/csrc/pgzip.h): byte-exact against zlib on every gzip flavour a
FASTQ file comes in (gzip -1/-6/-9, concatenated members, pigz-style sync-flushed blocks, stored blocks, BGZF), at
chunk sizes from far below to above a deflate block, with 1..8 threads; corrupt and truncated files must be reported;
and the reader (sb_reads_*) must deliver the same records from .gz as from the plain file."""
import gzip
import os
import subprocess
import zlib

import numpy as np
import pytest

from salmon_b200 import _capi

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
EXE = os.path.join(ROOT, "tests", "_build", "host_pgzip")
SRC = os.path.join(ROOT, "tests", "host_pgzip.cpp")
HDR = os.path.join(ROOT, "salmon_b200", "csrc", "pgzip.h")


def _exe():
    os.makedirs(os.path.dirname(EXE), exist_ok=True)
    if (not os.path.exists(EXE)) or os.path.getmtime(EXE) < max(os.path.getmtime(SRC), os.path.getmtime(HDR)):
        subprocess.check_call(["/usr/bin/g++", "-O2", "-std=c++17", "-Wall", "-o", EXE, SRC, "-lz", "-lpthread"])
    return EXE


def _run(path, threads, chunk, mode=None):
    r = subprocess.run([_exe(), str(path), str(threads), str(chunk)] + ([mode] if mode else []), capture_output=True, text=True, timeout=300)
    return r.returncode, r.stdout.strip()


def _fastq(rng, n, L=100):
    bases = np.frombuffer(b"ACGT", dtype=np.uint8)
    out = []
    for i in range(n):
        s = bases[rng.integers(0, 4, L)].tobytes()
        q = bytes(rng.integers(33, 74, L, dtype=np.uint8))
        out.append(b"@SRR1234567.%d %d/1 length=%d\n" % (i + 1, i + 1, L) + s + b"\n+\n" + q + b"\n")
    return b"".join(out)


def _bgzf(data):
    out = []
    for a in range(0, len(data), 65280):
        blk = data[a:a + 65280]
        c = zlib.compressobj(6, zlib.DEFLATED, -15)
        cd = c.compress(blk) + c.flush()
        hdr = b"\x1f\x8b\x08\x04\x00\x00\x00\x00\x00\xff\x06\x00BC\x02\x00" + (len(cd) + 25).to_bytes(2, "little")
        out.append(hdr + cd + zlib.crc32(blk).to_bytes(4, "little") + len(blk).to_bytes(4, "little"))
    out.append(bytes.fromhex("1f8b08040000000000ff0600424302001b0003000000000000000000"))   # the BGZF end-of-file block
    return b"".join(out)


@pytest.fixture(scope="module")
def files(tmp_path_factory):
    d = tmp_path_factory.mktemp("pgz")
    rng = np.random.default_rng(3)
    data = _fastq(rng, 60000)
    paths = {}
    for lvl in (1, 6, 9):
        p = d / f"l{lvl}.fq.gz"
        with gzip.open(p, "wb", compresslevel=lvl) as f:
            f.write(data)
        paths[f"l{lvl}"] = p
    p = d / "multi.fq.gz"
    with open(p, "wb") as f:
        step = len(data) // 7
        for a in range(0, len(data), step):
            f.write(gzip.compress(data[a:a + step], 6))
    paths["multi"] = p
    p = d / "flush.fq.gz"
    co = zlib.compressobj(6, zlib.DEFLATED, 31)
    with open(p, "wb") as f:
        for a in range(0, len(data), 131072):
            f.write(co.compress(data[a:a + 131072])); f.write(co.flush(zlib.Z_SYNC_FLUSH))
        f.write(co.flush())
    paths["flush"] = p
    p = d / "stored.fq.gz"
    with gzip.open(p, "wb", compresslevel=0) as f:
        f.write(data[:3_000_000])
    paths["stored"] = p
    p = d / "bgzf.fq.gz"
    p.write_bytes(_bgzf(data))
    paths["bgzf"] = p
    # incompressible bytes (stored blocks inside a -6 stream), a long run (distance-1 matches), 2-bit text
    mix = bytes(rng.integers(0, 256, 1_500_000, dtype=np.uint8)) + b"A" * 300000 + b"ACGT" * 200000 + \
        bytes(rng.integers(65, 69, 1_000_000, dtype=np.uint8))
    p = d / "mix.gz"
    with gzip.open(p, "wb", compresslevel=6) as f:
        f.write(mix)
    paths["mix"] = p
    p = d / "fixed.gz"       # tiny members use the fixed Huffman code
    with open(p, "wb") as f:
        for i in range(300):
            f.write(gzip.compress(b"@r%d\nACGTACGT\n+\nIIIIIIII\n" % i, 9))
    paths["fixed"] = p
    p = d / "empty.gz"
    with gzip.open(p, "wb") as f:
        pass
    paths["empty"] = p
    p = d / "padded.fq.gz"   # zero padding after the last member (tape blocks): ignored like zlib does
    p.write_bytes(paths["l6"].read_bytes() + b"\0" * 1000)
    paths["padded"] = p
    return paths


@pytest.mark.parametrize("name", ["l1", "l6", "l9", "multi", "flush", "stored", "bgzf", "mix", "fixed", "empty", "padded"])
def test_inflate_matches_zlib(files, name):
    for chunk in (3000, 50000, 700000, 2 << 20):
        for threads in (1, 3, 8):
            rc, out = _run(files[name], threads, chunk)
            assert rc == 0 and out.startswith("OK"), (name, chunk, threads, out)
    if name == "bgzf":
        assert out.split()[3] == "1"        # recognised as BGZF (member-parallel path)


def test_corrupt_and_truncated_files_are_reported(files, tmp_path):
    raw = files["l6"].read_bytes()
    rng = np.random.default_rng(9)
    for trial in range(12):
        b = bytearray(raw)
        pos = int(rng.integers(200, len(b) - 200))
        b[pos] ^= 1 << int(rng.integers(0, 8))
        p = tmp_path / f"c{trial}.gz"
        p.write_bytes(bytes(b))
        rc, out = _run(p, 4, 100000, "self")
        # a flipped bit gives a decode error or a CRC mismatch, reported by the inflater itself
        assert rc != 0 and out.startswith("FAIL error"), (trial, pos, out)
    for cut in (len(raw) - 1, len(raw) - 9, len(raw) // 2, 15):
        p = tmp_path / f"t{cut}.gz"
        p.write_bytes(raw[:cut])
        rc, out = _run(p, 4, 100000, "self")
        assert rc != 0 and "FAIL" in out, (cut, out)
    bg = files["bgzf"].read_bytes()
    b = bytearray(bg); b[len(b) // 2] ^= 0x10
    p = tmp_path / "cb.gz"; p.write_bytes(bytes(b))
    rc, out = _run(p, 4, 100000, "self")
    assert rc != 0 and "FAIL error" in out, out


ENC = np.full(256, 4, np.uint8)
for _c, _v in zip(b"ACGTacgt", [0, 1, 2, 3, 0, 1, 2, 3]):
    ENC[_c] = _v


def _read_all(f1, f2, threads, stride=128):
    got = []
    with _capi.ReadFiles(str(f1), str(f2) if f2 else None, n_threads=threads) as rf:
        while True:
            k, left, right, ll, lr = rf.next_batch(50000, stride)
            if k == 0:
                break
            for i in range(k):
                got.append((left[i, :ll[i]].tobytes(), right[i, :lr[i]].tobytes() if f2 else b""))
    return got


@pytest.mark.parametrize("flavour", ["gzip", "multi", "bgzf"])
def test_reader_same_records_from_gz_and_plain(tmp_path, monkeypatch, flavour):
    monkeypatch.setenv("SB_READS_INFLATERS", "4")
    monkeypatch.setenv("SB_READS_SCANNERS", "3")
    rng = np.random.default_rng(17)
    n = 120000
    bases = np.frombuffer(b"ACGTN", dtype=np.uint8)
    texts = []
    seqs = []
    for tag in (1, 2):
        recs, ss = [], []
        for i in range(n):
            L = int(rng.integers(40, 121))
            s = bases[rng.integers(0, 5 if i % 11 == 0 else 4, L)].tobytes()
            q = bytes(rng.integers(33, 74, L, dtype=np.uint8))
            if i % 5 == 0:
                q = b"@" + q[1:]
            recs.append(b"@SRR99.%d/%d\n%s\n+\n%s\n" % (i, tag, s, q))
            ss.append(ENC[np.frombuffer(s, dtype=np.uint8)].tobytes())
        t = b"".join(recs)
        if tag == 2:
            t = t[:-1]                      # the last record without its newline
        texts.append(t); seqs.append(ss)

    def pack(t):
        if flavour == "gzip":
            return gzip.compress(t, 6)
        if flavour == "multi":
            step = len(t) // 5 + 1
            return b"".join(gzip.compress(t[a:a + step], 4) for a in range(0, len(t), step))
        return _bgzf(t)
    f1, f2 = tmp_path / "r_1.fq.gz", tmp_path / "r_2.fq.gz"
    f1.write_bytes(pack(texts[0])); f2.write_bytes(pack(texts[1]))
    got = _read_all(f1, f2, threads=8)
    assert len(got) == n
    assert [g[0] for g in got] == seqs[0] and [g[1] for g in got] == seqs[1]
    # single inflate thread: zlib's gzread path gives the same
    monkeypatch.setenv("SB_READS_INFLATERS", "1")
    assert _read_all(f1, f2, threads=8) == got


def test_reader_gz_errors(tmp_path, monkeypatch):
    monkeypatch.setenv("SB_READS_INFLATERS", "4")
    rng = np.random.default_rng(23)
    t = _fastq(rng, 30000)
    z = gzip.compress(t, 6)
    f = tmp_path / "t.fq.gz"
    f.write_bytes(z[: len(z) // 2])
    with pytest.raises(_capi.SalmonB200Error):
        _read_all(f, None, threads=8)
    b = bytearray(z); b[len(b) // 3] ^= 0x40
    f.write_bytes(bytes(b))
    with pytest.raises(_capi.SalmonB200Error):
        _read_all(f, None, threads=8)
    # a record cut off at the end of the text (valid gzip, truncated FASTQ)
    f.write_bytes(gzip.compress(t[:-40], 6))
    with pytest.raises(_capi.SalmonB200Error, match="truncated|malformed"):
        _read_all(f, None, threads=8)
    # long records (longer than a piece's head room) still travel: 200 kb "reads"
    big = b"".join(b">c%d\n%s\n" % (i, bytes(rng.integers(65, 69, 200000, dtype=np.uint8))) for i in range(30))
    f.write_bytes(gzip.compress(big, 1))
    with _capi.ReadFiles(str(f), None, n_threads=8) as rf:
        with pytest.raises(_capi.SalmonB200Error, match="exceeds"):
            rf.next_batch(100, 256)


def test_eq_file_through_parallel_inflate(tmp_path):
    """whole-file readers (--eqclasses input, transcript FASTA) inflate regular gzip files above 8 MB with pgzip.h"""
    from salmon_b200.synth import synth_eq
    eq, proj, eff, uniq = synth_eq(seed=2, C=300000, M=80000, total_count=3_000_000)
    p = tmp_path / "eq.txt.gz"
    names = [f"t{i}" for i in range(eq.n_txps)]
    _capi.write_eq_classes(str(p), names, eq.off, eq.tids, eq.counts, eq.weights)
    assert os.path.getsize(p) > (8 << 20)
    got = _capi.read_eq_classes(str(p))
    plain = tmp_path / "eq.txt"
    plain.write_bytes(gzip.open(p, "rb").read())
    want = _capi.read_eq_classes(str(plain))
    assert got.keys() == want.keys()
    for k in got:
        a, b = got[k], want[k]
        if isinstance(a, np.ndarray):
            assert np.array_equal(a, b), k
        else:
            assert a == b, k


def test_mutated_files_never_crash_hang_or_pass(files, tmp_path):
    """byte flips, truncations, a garbage stretch, a header followed by noise: every mutated file must end in a reported
    error (240 such files were also run under ASan / UBSan while this was written)"""
    rng = np.random.default_rng(77)
    srcs = [files["l6"].read_bytes()[:400000], files["multi"].read_bytes()[:300000], files["bgzf"].read_bytes()[:200000]]
    for t in range(36):
        b = bytearray(srcs[t % 3])
        kind = t % 4
        if kind == 0:
            for _ in range(int(rng.integers(1, 20))):
                b[int(rng.integers(10, len(b)))] ^= int(rng.integers(1, 256))
        elif kind == 1:
            b = b[: int(rng.integers(11, len(b) - 1))]
        elif kind == 2:
            a = int(rng.integers(100, len(b) - 5000)); b[a:a + 4000] = bytes(rng.integers(0, 256, 4000, dtype=np.uint8))
        else:
            b = bytearray(b[:10]) + bytes(rng.integers(0, 256, int(rng.integers(10, 100000)), dtype=np.uint8))
        p = tmp_path / "m.gz"
        p.write_bytes(bytes(b))
        rc, out = _run(p, 4, int(rng.choice([3000, 40000, 500000])), "self")
        assert rc == 1 and out.startswith("FAIL"), (t, kind, rc, out)
