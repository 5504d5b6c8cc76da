// output.cu -- the output seam of the hot path (host code): quant.sf and eq_classes.txt[.gz] exactly in the
// reference's formats (src/output/GZipWriter.cpp:684-739 writeAbundances, :64-168 writeEquivCounts;
// doc/source/file_formats.rst).  No device work here; the numbers come from sb_em_optimize / sb_map_finish.
#include <stdio.h>
#include <string.h>
#include <zlib.h>

#include <algorithm>
#include <map>
#include <string>
#include <vector>

#include "common.cuh"

// TPM as writeAbundances computes it (GZipWriter.cpp:719-736): npm = count / numMappedFrags,
// tfracDenom = sum npm / effLen, tpm = (npm / effLen) / tfracDenom * 1e6
extern "C" int sb_tpm(uint32_t n, const double* alpha, const double* eff_len, double num_mapped_frags, double* tpm_out) {
  if (!alpha || !eff_len || !tpm_out) { sb::set_error("null argument"); return SB_ERR_INVALID; }
  double tfracDenom = 0.0;
  for (uint32_t i = 0; i < n; ++i) tfracDenom += (alpha[i] / num_mapped_frags) / eff_len[i];
  for (uint32_t i = 0; i < n; ++i) {
    const double npm = alpha[i] / num_mapped_frags;
    tpm_out[i] = ((npm / eff_len[i]) / tfracDenom) * 1000000.0;
  }
  return SB_OK;
}

extern "C" int sb_write_quant_sf(const char* path, uint32_t n, const char* const* names, const uint32_t* complete_len,
                                 const double* eff_len, const double* alpha, double num_mapped_frags, int sig_digits) {
  if (!path || !names || !complete_len || !eff_len || !alpha) { sb::set_error("null argument"); return SB_ERR_INVALID; }
  std::vector<double> tpm(n);
  sb_tpm(n, alpha, eff_len, num_mapped_frags, tpm.data());
  FILE* f = fopen(path, "w");
  if (!f) { sb::set_error("cannot open %s", path); return SB_ERR_INVALID; }
  bool ok = fprintf(f, "Name\tLength\tEffectiveLength\tTPM\tNumReads\n") > 0;
  for (uint32_t i = 0; i < n && ok; ++i)
    ok = fprintf(f, "%s\t%u\t%.*f\t%f\t%.*f\n", names[i], complete_len[i], sig_digits, eff_len[i], tpm[i], sig_digits, alpha[i]) > 0;
  ok = (fclose(f) == 0) && ok;     // a full disk shows up here at the latest
  if (!ok) { sb::set_error("write error on %s", path); return SB_ERR_INVALID; }
  return SB_OK;
}

namespace {
struct Sink {   // plain or gzip text sink
  FILE* f = nullptr; gzFile g = nullptr;
  bool open(const char* path) {
    const size_t L = strlen(path);
    if (L > 3 && !strcmp(path + L - 3, ".gz")) g = gzopen(path, "wb"); else f = fopen(path, "w");
    return f || g;
  }
  bool ok = true;
  void put(const std::string& s) {
    if (s.empty()) return;
    if (g) ok = ok && gzwrite(g, s.data(), (unsigned)s.size()) == (int)s.size();
    else ok = ok && fwrite(s.data(), 1, s.size(), f) == s.size();
  }
  bool close() {
    if (g) ok = (gzclose(g) == Z_OK) && ok;
    if (f) ok = (fclose(f) == 0) && ok;
    g = nullptr; f = nullptr;
    return ok;
  }
};
}  // namespace

// eq_classes.txt[.gz]: N, M, N names, then per class `k t_1..t_k [w_1..w_k] count` (tab separated).  weights == NULL
// is the reference's default (--dumpEq without --dumpEqWeights): range-factorised classes are collapsed by
// transcript set (GZipWriter.cpp:86-113); here they are written in label order (the reference's order is that of a
// hash map).  Weights are printed like an ostream prints a double (6 significant digits).
extern "C" int sb_write_eq_classes(const char* path, uint32_t n_txps, const char* const* names, uint64_t n_classes,
                                   const uint64_t* off, const uint32_t* tids, const double* weights,
                                   const uint64_t* counts) {
  if (!path || !names || !off || !tids || !counts) { sb::set_error("null argument"); return SB_ERR_INVALID; }
  Sink out;
  if (!out.open(path)) { sb::set_error("cannot open %s", path); return SB_ERR_INVALID; }
  char buf[64];
  std::string s;
  if (weights) {
    s = std::to_string(n_txps) + "\n" + std::to_string(n_classes) + "\n";
    for (uint32_t i = 0; i < n_txps; ++i) { s += names[i]; s += '\n'; }
    out.put(s);
    for (uint64_t c = 0; c < n_classes; ++c) {
      s.clear();
      const uint64_t b = off[c], e = off[c + 1];
      s += std::to_string(e - b); s += '\t';
      for (uint64_t j = b; j < e; ++j) { s += std::to_string(tids[j]); s += '\t'; }
      for (uint64_t j = b; j < e; ++j) { snprintf(buf, sizeof buf, "%g", weights[j]); s += buf; s += '\t'; }
      s += std::to_string(counts[c]); s += '\n';
      out.put(s);
    }
  } else {
    std::map<std::vector<uint32_t>, uint64_t> collapsed;
    for (uint64_t c = 0; c < n_classes; ++c)
      collapsed[std::vector<uint32_t>(tids + off[c], tids + off[c + 1])] += counts[c];
    s = std::to_string(n_txps) + "\n" + std::to_string(collapsed.size()) + "\n";
    for (uint32_t i = 0; i < n_txps; ++i) { s += names[i]; s += '\n'; }
    out.put(s);
    for (auto& kv : collapsed) {
      s.clear();
      s += std::to_string(kv.first.size()); s += '\t';
      for (uint32_t t : kv.first) { s += std::to_string(t); s += '\t'; }
      s += std::to_string(kv.second); s += '\n';
      out.put(s);
    }
  }
  if (!out.close()) { sb::set_error("write error on %s", path); return SB_ERR_INVALID; }
  return SB_OK;
}
