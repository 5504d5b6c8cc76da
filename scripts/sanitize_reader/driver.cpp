#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <vector>
#include <atomic>
#include "../../include/salmon_b200.h"
static std::atomic<unsigned long long> rows{0}, sum{0};
static int cb(void* user, const uint8_t* l, const uint8_t* r, uint32_t n, uint32_t L) {
  unsigned long long s = 0;
  for (size_t i = 0; i < (size_t)n * L; i += 97) s += l[i] + (r ? r[i] : 0);
  rows += n; sum += s;
  return 0;
}
int main(int argc, char** argv) {
  const char* f1[1] = {argv[1]}; const char* f2[1] = {argv[2]};
  for (int rep = 0; rep < 3; ++rep) {
    sb_reads* rd = sb_reads_open(f1, f2, 1, 4);
    if (!rd) { fprintf(stderr, "open: %s\n", sb_last_error()); return 1; }
    sb_bucket_stats st;
    int rc = sb_reads_bucketed(rd, 31, rep == 0 ? 4096 : 1000, 128, 4, 0, 1, cb, nullptr, &st);
    if (rc) { fprintf(stderr, "bucketed: %s\n", sb_last_error()); return 1; }
    printf("rep %d: observed %llu delivered %llu batches %llu lengths %u rows %llu\n", rep, (unsigned long long)st.n_observed,
           (unsigned long long)st.n_delivered, (unsigned long long)st.n_batches, st.n_read_lengths, (unsigned long long)rows.load());
    sb_reads_close(rd);
  }
  // plain next() loop too
  sb_reads* rd = sb_reads_open(f1, f2, 1, 4);
  std::vector<uint8_t> a(8192 * 128), b(8192 * 128); std::vector<uint32_t> la(8192), lb(8192);
  long long tot = 0, k;
  while ((k = sb_reads_next(rd, 8192, 128, a.data(), b.data(), la.data(), lb.data())) > 0) tot += k;
  printf("next loop: %lld\n", tot);
  sb_reads_close(rd);
  return 0;
}
