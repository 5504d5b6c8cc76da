// sb_salmon -- thin command-line front end over the C ABI (include/salmon_b200.h), keeping the `salmon index` /
// `salmon quant` invocations of the hot path (option names from src/core/ProgramOptionsGenerator.cpp:85-289 and
// src/index/BuildSalmonIndex.cpp:72-124).  No logic of its own: argument parsing, then sb_txome_read_fasta +
// sb_index_build + sb_index_save, or sb_index_load + sb_quant_files, or sb_eq_file_read + sb_em_optimize.
// Options outside the hot path (bias models, alignment mode, SAM output, ...) are rejected with a message.
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <sys/stat.h>

#include <string>
#include <vector>

#include "../../include/salmon_b200.h"

namespace {

int die(const char* what) {
  fprintf(stderr, "sb_salmon: %s: %s\n", what, sb_last_error());
  return 1;
}

struct Args {
  std::vector<std::string> v;
  size_t i = 0;
  bool more() const { return i < v.size(); }
  const std::string& peek() const { return v[i]; }
  std::string next() { return v[i++]; }
  bool value(std::string& out) {
    if (!more()) return false;
    out = next();
    return true;
  }
  // file lists: everything up to the next option
  void list(std::vector<std::string>& out) {
    while (more() && !(peek().size() > 1 && peek()[0] == '-')) {
      std::string a = next();
      size_t b = 0;   // also accept comma-separated lists
      while (b <= a.size()) {
        size_t e = a.find(',', b);
        if (e == std::string::npos) e = a.size();
        if (e > b) out.push_back(a.substr(b, e - b));
        b = e + 1;
      }
    }
  }
};

int usage() {
  fprintf(stderr,
          "sb_salmon (salmon-b200 %d): B200-native hot path of salmon\n"
          "  sb_salmon index -t transcripts.fa[.gz] -i index_dir [-k 31] [--gencode] [-d decoys.txt] [--keepDuplicates] [--no-clip]\n"
          "  sb_salmon quant -i index_dir -l IU -1 r1.fq[.gz] ... -2 r2.fq[.gz] ... -o out_dir [-p threads] [--dumpEq] [--dumpEqWeights]\n"
          "                  [--numBootstraps N | --numGibbsSamples N] [--thinningFactor 16] [--noGammaDraw] [--useEM] [--vbPrior 0.01]\n"
          "                  [--perNucleotidePrior] [--maxReadOcc 200] [--maxOccsPerHit 1000] [--minScoreFraction 0.65] [--consensusSlack 0.35]\n"
          "                  [--hardFilter] [--rangeFactorizationBins 4] [--fldMean 250] [--fldSD 25] [--fldMax 1000] [--scoreExp 1]\n"
          "                  [--numPreAuxModelSamples 5000] [--numAuxModelSamples 5000000] [--gpu 0] [--batch 262144] [--maxReadLen 256] [--seed 42]\n"
          "  sb_salmon quant -e eq_classes.txt[.gz] -o out_dir [--useEM] [--numBootstraps N]\n",
          sb_version());
  return 1;
}

int cmd_index(Args& a) {
  std::string fasta, dir, decoys, v;
  uint32_t k = 31;
  int gencode = 0, keep_dup = 0, no_clip = 0;
  while (a.more()) {
    const std::string o = a.next();
    if (o == "-t" || o == "--transcripts") { if (!a.value(fasta)) return usage(); }
    else if (o == "-i" || o == "--index") { if (!a.value(dir)) return usage(); }
    else if (o == "-k" || o == "--kmerLen") { if (!a.value(v)) return usage(); k = (uint32_t)atoi(v.c_str()); }
    else if (o == "-d" || o == "--decoys") { if (!a.value(decoys)) return usage(); }
    else if (o == "--gencode") gencode = 1;
    else if (o == "--keepDuplicates") keep_dup = 1;
    else if (o == "-n" || o == "--no-clip") no_clip = 1;
    else if (o == "-p" || o == "--threads" || o == "--type" || o == "-m" || o == "--minimizerLen" || o == "-f" ||
             o == "--filterSize" || o == "--tmpdir") { a.value(v); /* accepted, not needed by this index */ }
    else if (o == "--keepFixedFasta" || o == "--features") { fprintf(stderr, "sb_salmon index: %s is not supported\n", o.c_str()); return 1; }
    else { fprintf(stderr, "sb_salmon index: unknown option %s\n", o.c_str()); return usage(); }
  }
  if (fasta.empty() || dir.empty()) return usage();
  if (k == 0) k = 31;
  if ((k & 1) == 0) { fprintf(stderr, "Error: k must be an odd value, you chose %u.\n", k); return 1; }   // BuildSalmonIndex.cpp:206-211
  if (k > 31) { fprintf(stderr, "Error: k must not be larger than 31, you chose %u.\n", k); return 1; }
  sb_txome* t = nullptr;
  if (sb_txome_read_fasta(fasta.c_str(), k, gencode, decoys.empty() ? nullptr : decoys.c_str(), no_clip, keep_dup, &t) != 0)
    return die("reading the transcripts");
  fprintf(stderr, "read %u sequences (%u decoys), %llu bases; %u duplicates removed, %u poly-A tails clipped, %u shorter than k\n",
          t->n_txps, t->n_txps - t->first_decoy, (unsigned long long)t->seq_off[t->n_txps], t->n_duplicates_removed, t->n_clipped,
          t->n_short);
  sb_index* ix = sb_index_build(t->n_txps, t->seq_off, t->codes, k);
  if (!ix) return die("building the index");
  sb_index_set_meta(ix, t->names, t->complete_len, t->first_decoy);
  mkdir(dir.c_str(), 0777);
  if (sb_index_save(ix, (dir + "/sb_index.bin").c_str()) != 0) return die("writing the index");
  uint64_t info[4] = {0, 0, 0, 0};
  sb_index_info(ix, info);
  if (FILE* f = fopen((dir + "/info.json").c_str(), "w")) {
    fprintf(f, "{\n  \"index_type\": \"sb_kmer_table\",\n  \"k\": %u,\n  \"num_references\": %u,\n  \"first_decoy\": %u,\n"
               "  \"num_kmers\": %llu,\n  \"num_postings\": %llu,\n  \"table_capacity\": %llu,\n  \"bytes\": %llu\n}\n",
            k, t->n_txps, t->first_decoy, (unsigned long long)info[0], (unsigned long long)info[1], (unsigned long long)info[2],
            (unsigned long long)info[3]);
    fclose(f);
  }
  fprintf(stderr, "index: %llu distinct %u-mers, %llu postings, %.1f MB -> %s\n", (unsigned long long)info[0], k,
          (unsigned long long)info[1], (double)info[3] / 1e6, dir.c_str());
  sb_index_free(ix);
  sb_txome_free(t);
  return 0;
}

int quant_eqclasses(const std::string& eqfile, const std::string& out, sb_em_params ep, const sb_quant_opts& qo) {
  sb_eq_file* f = nullptr;
  if (sb_eq_file_read(eqfile.c_str(), &f) != 0) return die("reading the equivalence classes");
  if (!f->has_weights) { fprintf(stderr, "sb_salmon: --eqclasses needs the weights (write the file with --dumpEqWeights)\n"); return 1; }
  // processEqClasses, src/alignment/SalmonQuantifyAlignments.cpp:1406-1440: uniform initialisation, eq-class mode
  ep.init_uniform = 1;
  ep.eq_class_mode = 1;
  std::vector<double> zeros(f->n_txps, 0.0), alpha(f->n_txps, 0.0);
  std::vector<uint64_t> uniq(f->n_txps, 0);
  sb_eq_csr eq;
  eq.n_classes = f->n_classes; eq.n_txps = f->n_txps; eq.off = f->off; eq.tids = f->tids; eq.weights = f->weights; eq.counts = f->counts;
  sb_em_ctx* em = sb_em_create(qo.device);
  if (!em) return die("creating the optimiser");
  sb_em_stats st;
  const int rc = sb_em_optimize(em, &eq, &ep, zeros.data(), f->eff_len, uniq.data(), alpha.data(), &st);
  if (rc < 0) return die("optimising");
  if (rc == 1) { fprintf(stderr, "The optimization algorithm failed (total alpha weight too small)\n"); return 1; }
  double n_frags = 0;
  for (uint64_t c = 0; c < f->n_classes; ++c) n_frags += (double)f->counts[c];
  mkdir(out.c_str(), 0777);
  std::vector<uint32_t> lens(f->n_txps);
  for (uint32_t t = 0; t < f->n_txps; ++t) lens[t] = (uint32_t)(f->eff_len[t] < 1.0 ? 1.0 : f->eff_len[t]);
  if (sb_write_quant_sf((out + "/quant.sf").c_str(), f->n_txps, f->names, lens.data(), f->eff_len, alpha.data(), n_frags, 3) != 0)
    return die("writing quant.sf");
  fprintf(stderr, "%llu classes, %u transcripts: %u iterations (%s), %.1f ms on the device\n", (unsigned long long)f->n_classes,
          f->n_txps, st.iters, st.converged ? "converged" : "iteration limit", st.run_ms);
  sb_em_destroy(em);
  sb_eq_file_free(f);
  return 0;
}

int cmd_quant(Args& a) {
  std::string dir, out, lib = "A", eqfile, v;
  std::vector<std::string> m1, m2, unmated;
  sb_map_params mp;
  sb_em_params ep;
  sb_quant_opts qo;
  sb_map_default_params(&mp);
  sb_em_default_params(&ep);
  sb_quant_default_opts(&qo);
  auto num = [&](double& d) { if (!a.value(v)) return false; d = atof(v.c_str()); return true; };
  double d = 0;
  bool vb_prior_given = false;
  while (a.more()) {
    const std::string o = a.next();
    if (o == "-i" || o == "--index") { if (!a.value(dir)) return usage(); }
    else if (o == "-o" || o == "--output") { if (!a.value(out)) return usage(); }
    else if (o == "-l" || o == "--libType") { if (!a.value(lib)) return usage(); }
    else if (o == "-1" || o == "--mates1") a.list(m1);
    else if (o == "-2" || o == "--mates2") a.list(m2);
    else if (o == "-r" || o == "--unmatedReads") a.list(unmated);
    else if (o == "-e" || o == "--eqclasses") { if (!a.value(eqfile)) return usage(); }
    else if (o == "-p" || o == "--threads") { if (!num(d)) return usage(); qo.threads = (uint32_t)d; }
    else if (o == "--dumpEq") qo.dump_eq = 1;
    else if (o == "-d" || o == "--dumpEqWeights") qo.dump_eq_weights = 1;
    else if (o == "--numBootstraps") { if (!num(d)) return usage(); qo.num_bootstraps = (uint32_t)d; }
    else if (o == "--numGibbsSamples") { if (!num(d)) return usage(); qo.num_gibbs = (uint32_t)d; }
    else if (o == "--thinningFactor") { if (!num(d)) return usage(); qo.thinning = (uint32_t)d; }
    else if (o == "--noGammaDraw") qo.no_gamma_draw = 1;
    else if (o == "--useEM") ep.use_vbem = 0;
    else if (o == "--useVBOpt") ep.use_vbem = 1;
    else if (o == "--vbPrior") { if (!num(ep.vb_prior)) return usage(); vb_prior_given = true; }
    else if (o == "--perNucleotidePrior") ep.per_txp_prior = 0;
    else if (o == "--perTranscriptPrior") ep.per_txp_prior = 1;
    else if (o == "--initUniform") ep.init_uniform = 1;
    else if (o == "--noLengthCorrection") ep.no_length_correction = 1;
    else if (o == "--noRichEqClasses") ep.no_rich_eq = 1;
    else if (o == "--maxReadOcc") { if (!num(d)) return usage(); mp.max_read_occ = (uint32_t)d; }
    else if (o == "--maxOccsPerHit") { if (!num(d)) return usage(); mp.max_occs_per_hit = (uint32_t)d; }
    else if (o == "--minScoreFraction") { if (!num(mp.min_score_fraction)) return usage(); }
    else if (o == "--consensusSlack") { if (!num(d)) return usage(); mp.consensus_frac = 1.0 - d; }
    else if (o == "--hardFilter") mp.hard_filter = 1;
    else if (o == "--rangeFactorizationBins") { if (!num(d)) return usage(); mp.range_bins = (uint32_t)d; }
    else if (o == "--fldMean") { if (!num(mp.fld_mean)) return usage(); }
    else if (o == "--fldSD") { if (!num(mp.fld_sd)) return usage(); }
    else if (o == "--fldMax") { if (!num(d)) return usage(); mp.max_frag_len = (uint32_t)d; }
    else if (o == "--scoreExp") { if (!num(mp.score_exp)) return usage(); }
    else if (o == "--decoyThreshold") { if (!num(mp.decoy_threshold)) return usage(); }
    else if (o == "--ma") { if (!num(d)) return usage(); mp.ma = (int32_t)d; }
    else if (o == "--mp") { if (!num(d)) return usage(); mp.mp = (int32_t)d; }
    else if (o == "--go") { if (!num(d)) return usage(); mp.go = (int32_t)d; }
    else if (o == "--ge") { if (!num(d)) return usage(); mp.ge = (int32_t)d; }
    else if (o == "--bandwidth") { if (!num(d)) return usage(); mp.band = (uint32_t)d; }
    else if (o == "--numPreAuxModelSamples") { if (!num(d)) return usage(); mp.num_pre_burnin = (uint64_t)d; }
    else if (o == "--numAuxModelSamples") { if (!num(d)) return usage(); mp.num_burnin = (uint64_t)d; }
    else if (o == "--gpu") { if (!num(d)) return usage(); qo.device = (int32_t)d; }
    else if (o == "--batch") { if (!num(d)) return usage(); qo.batch = (uint32_t)d; }
    else if (o == "--maxReadLen") { if (!num(d)) return usage(); qo.max_read_len = (uint32_t)d; }
    else if (o == "--seed") { if (!num(d)) return usage(); qo.seed = (uint64_t)d; mp.seed = (uint64_t)d; }
    else if (o == "--validateMappings" || o == "--softclipOverhangs" || o == "-q" || o == "--quiet") { /* default behaviour / no-op */ }
    else if (o == "--seqBias" || o == "--gcBias" || o == "--posBias" || o == "--writeMappings" || o == "-z" || o == "-a" ||
             o == "--alignments" || o == "--recoverOrphans" || o == "-g" || o == "--geneMap" || o == "--sketchMode") {
      fprintf(stderr, "sb_salmon quant: %s is outside the hot path this build replaces (DESIGN.md, out of scope)\n", o.c_str());
      return 1;
    } else { fprintf(stderr, "sb_salmon quant: unknown option %s\n", o.c_str()); return usage(); }
  }
  if (out.empty()) return usage();
  // --perNucleotidePrior without an explicit --vbPrior: the reference switches the default to 1e-5
  // (src/cli/QuantOptionsUtils.cpp:569-572)
  if (ep.use_vbem && !ep.per_txp_prior && !vb_prior_given) ep.vb_prior = 1e-5;
  if (!eqfile.empty()) return quant_eqclasses(eqfile, out, ep, qo);
  if (dir.empty()) return usage();
  if (!unmated.empty() || m1.empty() || m1.size() != m2.size()) {
    fprintf(stderr, "sb_salmon quant: paired-end input (-1 / -2, the same number of files) is what this build maps; single-end "
                    "reads (-r) are not supported yet\n");
    return 1;
  }
  if (lib != "A" && lib != "IU") {
    fprintf(stderr, "sb_salmon quant: library type %s: only IU (or A, taken as IU) is supported\n", lib.c_str());
    return 1;
  }
  sb_index* ix = sb_index_load((dir + "/sb_index.bin").c_str());
  if (!ix) return die("loading the index");
  std::vector<const char*> p1, p2;
  for (auto& s : m1) p1.push_back(s.c_str());
  for (auto& s : m2) p2.push_back(s.c_str());
  sb_quant_summary sum;
  if (sb_quant_files(ix, p1.data(), p2.data(), (uint32_t)p1.size(), &mp, &ep, &qo, out.c_str(), nullptr, &sum) != 0)
    return die("quant");
  const double rate = sum.n_observed ? 100.0 * (double)sum.n_mapped / (double)sum.n_observed : 0.0;
  fprintf(stderr, "%llu fragments observed, %llu mapped (%.4f%%), %llu equivalence classes; %u read lengths, %llu batches\n",
          (unsigned long long)sum.n_observed, (unsigned long long)sum.n_mapped, rate, (unsigned long long)sum.n_classes,
          sum.n_read_lengths, (unsigned long long)sum.n_batches);
  fprintf(stderr, "mapping %.2f s (%.1f ms on the device, %.2f M fragments/s end to end), optimiser %u iterations in %.2f s, total %.2f s\n",
          sum.map_seconds, sum.map_device_ms, sum.map_seconds > 0 ? (double)sum.n_observed / sum.map_seconds / 1e6 : 0.0, sum.em_iters,
          sum.em_seconds, sum.total_seconds);
  if (FILE* f = fopen((out + "/aux_info/meta_info.json").c_str(), "w")) {   // the fields downstream tools read (tximport: num_bootstraps ...)
    fprintf(f, "{\n  \"salmon_version\": \"sb-%d\",\n  \"samp_type\": \"%s\",\n  \"num_bootstraps\": %u,\n  \"num_processed\": %llu,\n"
               "  \"num_mapped\": %llu,\n  \"percent_mapped\": %.6f,\n  \"num_eq_classes\": %llu,\n  \"mapping_type\": \"mapping\"\n}\n",
            sb_version(), qo.num_gibbs ? "gibbs" : (qo.num_bootstraps ? "bootstrap" : "none"), qo.num_gibbs ? qo.num_gibbs : qo.num_bootstraps,
            (unsigned long long)sum.n_observed, (unsigned long long)sum.n_mapped, rate, (unsigned long long)sum.n_classes);
    fclose(f);
  }
  sb_index_free(ix);
  return 0;
}

}  // namespace

int main(int argc, char** argv) {
  if (argc < 2) return usage();
  Args a;
  for (int i = 2; i < argc; ++i) a.v.push_back(argv[i]);
  const std::string cmd = argv[1];
  if (cmd == "index") return cmd_index(a);
  if (cmd == "quant") return cmd_quant(a);
  if (cmd == "--version" || cmd == "-v") { printf("sb_salmon %d\n", sb_version()); return 0; }
  return usage();
}
