"""Host reader benchmark (no GPU needed): FASTQ -> base-code batches through sb_reads_bucketed with a no-op consumer.
usage: bench_reader.py n_pairs read_len threads [gz]"""
import os, sys, time, gzip
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from salmon_b200._capi import ReadFiles as Reads
n = int(sys.argv[1]); L = int(sys.argv[2]); thr = int(sys.argv[3]); gz = len(sys.argv) > 4
D = "/dev/shm/rb"; os.makedirs(D, exist_ok=True)
paths = [f"{D}/r_{t}_{n}_{L}.fq" + (".gz" if gz else "") for t in (1, 2)]
if not all(os.path.exists(p) for p in paths):
    rng = np.random.default_rng(1)
    lut = np.frombuffer(b"ACGT", dtype=np.uint8)
    for p in paths:
        codes = rng.integers(0, 4, size=(n, L), dtype=np.uint8)
        rec = np.empty((n, 3 + L + 3 + L + 1), dtype=np.uint8)
        rec[:, 0:3] = np.frombuffer(b"@r\n", dtype=np.uint8); rec[:, 3:3 + L] = lut[codes]
        rec[:, 3 + L:6 + L] = np.frombuffer(b"\n+\n", dtype=np.uint8); rec[:, 6 + L:6 + 2 * L] = ord("I"); rec[:, 6 + 2 * L] = 10
        if gz:
            with gzip.open(p, "wb", compresslevel=4) as f: f.write(rec.tobytes())
        else:
            rec.tofile(p)
os.environ["SB_READS_PROFILE"] = "1"
for rep in range(2):
    t0 = time.time()
    rd = Reads(paths[0], paths[1], n_threads=thr)
    tot = [0]
    def fn(l, r, LL): tot[0] += l.shape[0]
    st = rd.bucketed(fn, batch=262144, max_read_len=L + 28, threads=thr)
    rd.close()
    dt = time.time() - t0
    print(f"{tot[0]} pairs in {dt:.2f} s = {tot[0] / dt / 1e6:.2f} M pairs/s ({2 * os.path.getsize(paths[0]) / dt / 1e9:.2f} GB/s of text)", flush=True)
