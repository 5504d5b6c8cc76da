// host_pgzip.cpp -- test driver: inflate a .gz file with sb::pgz::ParallelGz and compare with zlib's gzread.
// usage: host_pgzip file.gz threads chunk_bytes [bench]      prints "OK <bytes> <pieces> <bgzf>" or "FAIL ..."
#include <fcntl.h>
#include <stdio.h>
#include <sys/mman.h>
#include <sys/stat.h>
#include <time.h>
#include <unistd.h>

#include "../salmon_b200/csrc/pgzip.h"

static double wall() { timespec ts; clock_gettime(CLOCK_MONOTONIC, &ts); return ts.tv_sec + 1e-9 * ts.tv_nsec; }

int main(int argc, char** argv) {
  if (argc < 4) { fprintf(stderr, "usage\n"); return 2; }
  const char* path = argv[1];
  const int threads = atoi(argv[2]);
  const size_t chunk = (size_t)atol(argv[3]);
  const bool bench = argc > 4;   // "bench": time both; "self": no comparison, only what the inflater itself reports
  int fd = open(path, O_RDONLY);
  if (fd < 0) { printf("FAIL open\n"); return 1; }
  struct stat sb;
  fstat(fd, &sb);
  const size_t len = (size_t)sb.st_size;
  const uint8_t* d = (const uint8_t*)mmap(nullptr, len ? len : 1, PROT_READ, MAP_PRIVATE, fd, 0);
  std::vector<uint8_t> ref;
  double t_ref = 0;
  {
    const double t0 = wall();
    gzFile g = gzopen(path, "rb");
    gzbuffer(g, 1 << 20);
    std::vector<uint8_t> b(1 << 22);
    for (;;) {
      const int got = gzread(g, b.data(), (unsigned)b.size());
      if (got <= 0) break;
      if (!bench) ref.insert(ref.end(), b.begin(), b.begin() + got);
      else ref.resize(ref.size() + 0), t_ref += 0;
    }
    gzclose(g);
    t_ref = wall() - t0;
  }
  const double t0 = wall();
  sb::pgz::ParallelGz pg(d, len, threads, chunk);
  std::string err;
  if (!pg.start(err)) { printf("FAIL start: %s\n", err.c_str()); return 1; }
  size_t total = 0, pieces = 0;
  bool ok = true;
  sb::pgz::Piece pc;
  while (pg.next(pc, err)) {
    if (!bench) {
      if (total + pc.len > ref.size() || memcmp(ref.data() + total, pc.data, pc.len) != 0) {
        size_t k = 0;
        while (total + k < ref.size() && k < pc.len && ref[total + k] == pc.data[k]) ++k;
        printf("FAIL mismatch at byte %zu (piece %zu, len %zu)\n", total + k, pieces, pc.len);
        ok = false;
        break;
      }
      // the head room must be writable
      memset(pc.data - sb::pgz::HEAD, 0xAB, sb::pgz::HEAD);
    }
    total += pc.len;
    ++pieces;
  }
  const double t1 = wall() - t0;
  if (!err.empty()) { printf("FAIL error: %s\n", err.c_str()); return 1; }
  if (ok && !bench && total != ref.size()) { printf("FAIL length %zu vs %zu\n", total, ref.size()); return 1; }
  if (!ok) return 1;
  printf("OK %zu %zu %d", total, pieces, pg.is_bgzf() ? 1 : 0);
  if (bench) printf("  pgz %.3f s (%.2f GB/s; workers %.3f cpu-s of which marker replacement + CRC %.3f, consumer %.3f cpu-s, %.1f%% marker symbols)  zlib %.3f s (%.2f GB/s)", t1, total / t1 / 1e9,
                    pg.worker_cpu_s(), pg.resolve_cpu_s(), pg.consumer_cpu_s(), 100.0 * pg.marker_symbols() / (total ? total : 1), t_ref, total / t_ref / 1e9);
  printf("\n");
  return 0;
}
