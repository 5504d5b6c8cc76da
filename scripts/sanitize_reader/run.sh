#!/bin/bash
# Host-side sanitizer runs of the reader / length-bucketing code (salmon_b200/csrc/ingest.cu, pipeline.cu), no GPU needed:
#   tsan : ThreadSanitizer, built WITHOUT OpenMP (libgomp's barriers are invisible to TSan and drown the report), so the
#          std::thread / mutex / condition-variable logic is what is checked
#   asan : AddressSanitizer + UndefinedBehaviorSanitizer + LeakSanitizer, with OpenMP, plain (memory-mapped) and gzip input
# usage: scripts/sanitize_reader/run.sh   (from the repository root)
set -e
HERE=$(cd "$(dirname "$0")" && pwd); ROOT=$(cd "$HERE/../.." && pwd); W=$(mktemp -d)
python - "$W" <<'PY'
import sys, numpy as np
rng = np.random.default_rng(3); lut = np.frombuffer(b"ACGT", dtype=np.uint8); n = 60000
def write(path, lens):
    with open(path, "wb") as f:
        for i, L in enumerate(lens):
            f.write(b"@r%d\n%s\n+\n%s\n" % (i, lut[rng.integers(0, 4, size=L)].tobytes(), b"I" * L))
l1 = np.where(rng.random(n) < 0.7, 100, rng.integers(20, 101, size=n)); l2 = np.where(rng.random(n) < 0.7, l1, rng.integers(20, 101, size=n))
write(sys.argv[1] + "/t_1.fq", l1); write(sys.argv[1] + "/t_2.fq", l2)
PY
gzip -k "$W/t_1.fq" "$W/t_2.fq"
SRC="$HERE/driver.cpp $ROOT/salmon_b200/csrc/ingest.cu $ROOT/salmon_b200/csrc/pipeline.cu $HERE/stubs.cu"
NV="nvcc -O1 -g -std=c++17 -gencode arch=compute_100a,code=sm_100a"
$NV -Xcompiler -fsanitize=thread -Xcompiler -Wno-unknown-pragmas -o "$W/tsan" $SRC -lz -ltsan
$NV -Xcompiler -fsanitize=address -Xcompiler -fsanitize=undefined -Xcompiler -fopenmp -o "$W/asan" $SRC -lgomp -lz -lasan -lubsan
echo "== tsan"; TSAN_OPTIONS=halt_on_error=1 "$W/tsan" "$W/t_1.fq" "$W/t_2.fq"
echo "== asan plain"; ASAN_OPTIONS=detect_leaks=1:protect_shadow_gap=0 "$W/asan" "$W/t_1.fq" "$W/t_2.fq"
echo "== asan gzip"; ASAN_OPTIONS=detect_leaks=1:protect_shadow_gap=0 "$W/asan" "$W/t_1.fq.gz" "$W/t_2.fq.gz"
rm -rf "$W"; echo "sanitizers clean"
