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
  // the other file parsers of ingest.cu
  if (argc > 4) {
    sb_eq_file* f = nullptr;
    if (sb_eq_file_read(argv[3], &f) != 0) { fprintf(stderr, "eq: %s\n", sb_last_error()); return 1; }
    printf("eq file: %u transcripts, %llu classes, weights %u, missing eff %u\n", f->n_txps, (unsigned long long)f->n_classes,
           f->has_weights, f->n_missing_eff_len);
    sb_bootstrap_writer* w = sb_bootstrap_writer_open("/dev/null");
    if (w) { sb_bootstrap_writer_write(w, f->eff_len, f->n_txps); sb_bootstrap_writer_close(w); }
    sb_eq_file_free(f);
    sb_txome* t = nullptr;
    if (sb_txome_read_fasta(argv[4], 31, 1, nullptr, 0, 0, &t) != 0) { fprintf(stderr, "txome: %s\n", sb_last_error()); return 1; }
    printf("txome: %u sequences, %u duplicates removed, %u clipped, %u short\n", t->n_txps, t->n_duplicates_removed, t->n_clipped,
           t->n_short);
    sb_txome_free(t);
    // malformed inputs must fail cleanly
    sb_eq_file* g = nullptr;
    if (sb_eq_file_read(argv[4], &g) == 0) { fprintf(stderr, "a FASTA parsed as an eq file\n"); return 1; }
    sb_txome* u = nullptr;
    if (sb_txome_read_fasta(argv[3], 31, 0, nullptr, 0, 0, &u) == 0) { fprintf(stderr, "an eq file parsed as FASTA\n"); return 1; }
  }
  return 0;
}
