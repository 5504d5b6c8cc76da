// pipeline.cu -- the host driver of the hot path (C++, no kernels of its own): what processReadLibrary / quantifyLibrary
// / stageFinalizeMappingOutputs do around the seams in the reference (src/quant/SalmonQuantify.cpp:2339-2730 spawn the
// parser + worker threads; src/quant/pipeline/MappingPipelineStages.cpp:37-206 optimise and write), here as
//
//     reader thread:  sb_reads_next -> rows grouped by read length into pinned [n, L] buffers   (host cores)
//     caller thread:  sb_map_batch per full buffer                                              (GPU)
//     then:           sb_map_finish -> sb_em_optimize -> (sb_bootstrap | sb_gibbs) -> quant.sf, eq_classes, bootstraps.gz
//
// so parsing / packing of the next batch overlaps the kernels of the current one.  One GPU per call; multi-GPU runs
// shard the read stream over processes (shard_index / shard_count) and do the once-per-run reduction in the host layer
// (salmon_b200/dist.py).  sb_map_batch takes one read length per call: a pair whose mates differ in length is mapped
// at the shorter length (documented deviation until the kernels take per-mate lengths).
#include <errno.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <sys/stat.h>
#include <time.h>
#include <zlib.h>

#include <condition_variable>
#include <deque>
#include <map>
#include <memory>
#include <mutex>
#include <string>
#include <thread>
#include <vector>

#include "common.cuh"

namespace {

double now_s() {
  timespec ts;
  clock_gettime(CLOCK_MONOTONIC, &ts);
  return (double)ts.tv_sec + 1e-9 * (double)ts.tv_nsec;
}

struct HostBuf {   // pinned [cap, L] byte matrix pair
  uint8_t* left = nullptr;
  uint8_t* right = nullptr;
  uint32_t cap = 0, L = 0, n = 0;
  bool busy = false;   // handed to the consumer side
  bool pinned = false;
  int alloc(uint32_t cap_, uint32_t L_) {
    cap = cap_; L = L_; n = 0;
    const size_t bytes = (size_t)cap * L;
    // page-locked when a CUDA device is there (the consumer copies these to the GPU), plain memory otherwise
    if (cudaMallocHost(&left, bytes) == cudaSuccess && cudaMallocHost(&right, bytes) == cudaSuccess) { pinned = true; return SB_OK; }
    cudaGetLastError();
    if (left) { cudaFreeHost(left); left = nullptr; }
    left = (uint8_t*)malloc(bytes ? bytes : 1);
    right = (uint8_t*)malloc(bytes ? bytes : 1);
    if (!left || !right) { sb::set_error("cannot allocate %zu bytes of host memory", 2 * bytes); return SB_ERR_NOMEM; }
    return SB_OK;
  }
  void release() {
    if (pinned) { if (left) cudaFreeHost(left); if (right) cudaFreeHost(right); }
    else { free(left); free(right); }
    left = right = nullptr;
  }
};

struct Bucket {   // reads of one length: two buffers, the reader fills one while the GPU consumes the other
  HostBuf buf[2];
  int fill = 0;
  uint32_t cnt = 0;        // rows in buf[fill] (reader-private; HostBuf::n is set at submit)
  bool checked = true;     // buf[fill] is known to be released by the GPU side
};

struct Job {
  HostBuf* b;
};

struct Pipe {
  std::mutex mu;
  std::condition_variable cv_job, cv_free;
  std::deque<Job> jobs;
  bool reader_done = false;
  bool abort = false;
  std::string err;
  uint64_t n_observed = 0, n_too_short = 0, n_trimmed_mates = 0;
};

struct BootUser {
  sb_bootstrap_writer* w;
  int rc;
};
int boot_cb(const double* alpha, uint32_t n, void* user) {
  BootUser* u = (BootUser*)user;
  u->rc = sb_bootstrap_writer_write(u->w, alpha, n);
  return u->rc;
}

bool make_dirs(const std::string& path) {
  std::string cur;
  for (size_t i = 0; i <= path.size(); ++i) {
    if (i == path.size() || path[i] == '/') {
      if (!cur.empty() && cur != "/") {
        if (mkdir(cur.c_str(), 0777) != 0 && errno != EEXIST) return false;
      }
    }
    if (i < path.size()) cur += path[i];
  }
  return true;
}

}  // namespace

extern "C" int sb_reads_bucketed(sb_reads* rd, uint32_t min_len, uint32_t batch, uint32_t max_read_len, uint32_t threads,
                                 uint32_t shard_index, uint32_t shard_count, sb_batch_cb cb, void* user,
                                 sb_bucket_stats* stats) {
  if (!rd || !cb) { sb::set_error("sb_reads_bucketed: null argument"); return SB_ERR_INVALID; }
  if (batch < 1) batch = 1;
  if (max_read_len < 1) { sb::set_error("sb_reads_bucketed: max_read_len must be positive"); return SB_ERR_INVALID; }
  if (shard_count == 0 || shard_index >= shard_count) { sb::set_error("bad shard index / count"); return SB_ERR_INVALID; }
  if (threads == 0) threads = 1;
  const bool paired = sb_reads_paired(rd) != 0;
  Pipe P;
  std::map<uint32_t, std::unique_ptr<Bucket>> buckets;
  const uint32_t stride = max_read_len;
  // ---- reader side ---------------------------------------------------------------------------------------------
  auto submit = [&](HostBuf* b) {   // hand a filled buffer to the consumer side
    std::unique_lock<std::mutex> lk(P.mu);
    b->busy = true;
    P.jobs.push_back(Job{b});
    P.cv_job.notify_one();
  };
  auto reader = [&]() {
    std::vector<uint8_t> sl, sr;          // staging for batches of mixed lengths (allocated on first use)
    std::vector<uint32_t> ll(batch), lr(batch);
    std::string err;
    // the bucket of length L with room for at least one row; nullptr on error / abort
    auto bucket_for = [&](uint32_t L) -> Bucket* {
      std::unique_ptr<Bucket>& bp = buckets[L];
      if (!bp) {
        bp.reset(new Bucket());
        // the first length seen gets full-size buffers; rarer lengths smaller ones
        const uint32_t cap = buckets.size() == 1 ? batch : std::max<uint32_t>(batch / 8, std::min<uint32_t>(batch, 4096));
        if (bp->buf[0].alloc(cap, L) != SB_OK || bp->buf[1].alloc(cap, L) != SB_OK) { err = sb_last_error(); return nullptr; }
      }
      HostBuf* b = &bp->buf[bp->fill];
      if (!bp->checked) {   // first row after a flip: the consumer must have released this buffer
        std::unique_lock<std::mutex> lk(P.mu);
        P.cv_free.wait(lk, [&] { return !b->busy || P.abort; });
        if (P.abort) { err = "aborted"; return nullptr; }
        bp->checked = true;
      }
      return bp.get();
    };
    auto filled = [&](Bucket* bp, uint32_t rows) {
      HostBuf* b = &bp->buf[bp->fill];
      bp->cnt += rows;
      if (bp->cnt == b->cap) { b->n = bp->cnt; submit(b); bp->fill ^= 1; bp->cnt = 0; bp->checked = false; }
    };
    // rows [i0, i1) of the staging buffers -> the bucket of length L (copied by an OpenMP team when it is a run)
    auto put_rows = [&](uint32_t L, int64_t i0, int64_t i1) {
      while (i0 < i1 && err.empty()) {
        Bucket* bp = bucket_for(L);
        if (!bp) return;
        HostBuf* b = &bp->buf[bp->fill];
        const int64_t take = std::min<int64_t>(i1 - i0, (int64_t)(b->cap - bp->cnt));
        const int nt = (int)std::max<int64_t>(1, std::min<int64_t>(threads, take / 4096));
#pragma omp parallel for schedule(static) num_threads(nt)
        for (int64_t q = 0; q < take; ++q) {
          memcpy(b->left + (size_t)(bp->cnt + q) * L, sl.data() + (size_t)(i0 + q) * stride, L);
          if (paired) memcpy(b->right + (size_t)(bp->cnt + q) * L, sr.data() + (size_t)(i0 + q) * stride, L);
        }
        i0 += take;
        filled(bp, (uint32_t)take);
      }
    };
    // the read stream is cut into global batches of `batch` records; batch g belongs to shard g % shard_count
    for (uint64_t g = 0; err.empty(); ++g) {
      { std::lock_guard<std::mutex> lk(P.mu); if (P.abort) break; }
      uint32_t L0 = 0;
      int64_t n = sb_reads_peek(rd, batch, &L0);
      if (n < 0) { err = sb_last_error(); break; }
      if (n == 0) break;
      P.n_observed += (uint64_t)n;
      if (g % shard_count != shard_index) {
        if (sb_reads_skip(rd, (uint32_t)n) != n) { err = sb_last_error(); break; }
        continue;
      }
      if (L0 >= min_len && L0 > 0 && L0 <= stride) {
        // one length (the usual case): translated straight into the bucket buffer, no staging, no copy
        int64_t left_n = n;
        while (left_n > 0 && err.empty()) {
          Bucket* bp = bucket_for(L0);
          if (!bp) break;
          HostBuf* b = &bp->buf[bp->fill];
          const uint32_t take = (uint32_t)std::min<int64_t>(left_n, (int64_t)(b->cap - bp->cnt));
          const int64_t got = sb_reads_next(rd, take, L0, b->left + (size_t)bp->cnt * L0, b->right + (size_t)bp->cnt * L0,
                                            ll.data(), lr.data());
          if (got != (int64_t)take) { err = got < 0 ? sb_last_error() : "short read from the parser"; break; }
          left_n -= take;
          filled(bp, take);
        }
        continue;
      }
      if (sl.empty()) { sl.resize((size_t)batch * stride); sr.resize((size_t)batch * stride); }
      const int64_t got = sb_reads_next(rd, (uint32_t)n, stride, sl.data(), sr.data(), ll.data(), lr.data());
      if (got != n) { err = got < 0 ? sb_last_error() : "short read from the parser"; break; }
      // runs of equal length go in one piece
      int64_t run0 = 0;
      uint32_t runL = 0;
      for (int64_t i = 0; i <= n && err.empty(); ++i) {
        uint32_t L = 0;
        if (i < n) {
          L = paired ? std::min(ll[i], lr[i]) : ll[i];
          if (paired && ll[i] != lr[i]) ++P.n_trimmed_mates;
          if (L < min_len || L == 0) { ++P.n_too_short; L = 0; }   // cannot hold a k-mer: observed, never delivered
        }
        if (i == n || L != runL) {
          if (runL != 0 && i > run0) put_rows(runL, run0, i);
          run0 = i; runL = L;
        }
      }
    }
    if (err == "aborted") err.clear();
    if (err.empty())
      for (auto& kv : buckets) {   // ascending read length
        Bucket& bk = *kv.second;
        if (bk.cnt > 0) { HostBuf* b = &bk.buf[bk.fill]; b->n = bk.cnt; bk.cnt = 0; submit(b); }
      }
    std::lock_guard<std::mutex> lk(P.mu);
    if (!err.empty()) P.err = err;
    P.reader_done = true;
    P.cv_job.notify_all();
  };
  std::thread rt(reader);
  // ---- consumer side (the calling thread) ------------------------------------------------------------------------
  int rc = SB_OK;
  std::string cb_err;
  uint64_t n_batches = 0, n_delivered = 0;
  for (;;) {
    Job j{nullptr};
    {
      std::unique_lock<std::mutex> lk(P.mu);
      P.cv_job.wait(lk, [&] { return !P.jobs.empty() || P.reader_done; });
      if (P.jobs.empty()) break;
      j = P.jobs.front();
      P.jobs.pop_front();
    }
    if (rc == SB_OK) {
      rc = cb(user, j.b->left, paired ? j.b->right : nullptr, j.b->n, j.b->L);
      if (rc != SB_OK) {
        cb_err = sb_last_error();
        std::lock_guard<std::mutex> lk(P.mu);
        P.abort = true;
      } else {
        ++n_batches;
        n_delivered += j.b->n;
      }
    }
    {
      std::lock_guard<std::mutex> lk(P.mu);
      j.b->n = 0;
      j.b->busy = false;
    }
    P.cv_free.notify_all();
  }
  rt.join();
  const uint32_t n_lengths = (uint32_t)buckets.size();
  for (auto& kv : buckets) { kv.second->buf[0].release(); kv.second->buf[1].release(); }
  if (stats) {
    memset(stats, 0, sizeof(*stats));
    stats->n_observed = P.n_observed; stats->n_delivered = n_delivered; stats->n_too_short = P.n_too_short;
    stats->n_trimmed_mates = P.n_trimmed_mates; stats->n_batches = n_batches; stats->n_read_lengths = n_lengths;
  }
  if (rc != SB_OK) { sb::set_error("%s", cb_err.c_str()); return rc > 0 ? SB_ERR_STATE : rc; }
  if (!P.err.empty()) { sb::set_error("%s", P.err.c_str()); return SB_ERR_INVALID; }
  return SB_OK;
}

extern "C" void sb_quant_default_opts(sb_quant_opts* o) {
  if (!o) return;
  memset(o, 0, sizeof(*o));
  o->device = 0;
  o->batch = 262144;
  o->max_read_len = 256;
  o->threads = 8;
  o->shard_index = 0;
  o->shard_count = 1;
  o->thinning = 16;
  o->seed = 42;
}

extern "C" int sb_quant_files(sb_index* ix, const char* const* mates1, const char* const* mates2, uint32_t n_files,
                              const sb_map_params* mp_in, const sb_em_params* ep_in, const sb_quant_opts* o_in,
                              const char* out_dir, double* alpha_out, sb_quant_summary* sum) {
  if (!ix || !mates1 || !mates2 || !n_files) { sb::set_error("sb_quant_files: null argument (paired-end input only)"); return SB_ERR_INVALID; }
  sb_quant_opts o;
  if (o_in) o = *o_in; else sb_quant_default_opts(&o);
  if (o.batch < 1024) o.batch = 1024;
  if (o.max_read_len < 32 || o.max_read_len > 256) { sb::set_error("max_read_len must be in 32..256"); return SB_ERR_INVALID; }
  if (o.shard_count != 1 || o.shard_index != 0) {
    // a shard's classes alone are not a quantification: multi-GPU runs reduce the end-of-mapping statistics over the
    // ranks before the EM (salmon_b200.quant.quant_files: sb_reads_bucketed + sb_map_partial_get / sb_map_project_global)
    sb::set_error("sb_quant_files quantifies one whole library on one GPU; shard the reads with sb_reads_bucketed and the multi-GPU host layer");
    return SB_ERR_INVALID;
  }
  if (o.num_bootstraps && o.num_gibbs) { sb::set_error("choose bootstraps or Gibbs samples, not both"); return SB_ERR_INVALID; }
  sb_map_params mp;
  if (mp_in) mp = *mp_in; else sb_map_default_params(&mp);
  sb_em_params ep;
  if (ep_in) ep = *ep_in; else sb_em_default_params(&ep);
  uint32_t M = 0, k = 0, first_decoy = 0;
  const char* const* names = nullptr;
  const uint32_t* complete_len = nullptr;
  sb_index_get_meta(ix, &M, &k, &first_decoy, &names, &complete_len);
  if (first_decoy < M) mp.first_decoy = (int32_t)first_decoy;
  const double t0 = now_s();

  sb_map_ctx* ctx = sb_map_create(ix, &mp, o.device, o.batch, o.max_read_len);
  if (!ctx) return SB_ERR_CUDA;
  sb_reads* rd = sb_reads_open(mates1, mates2, n_files, o.threads);
  if (!rd) { sb_map_destroy(ctx); return SB_ERR_INVALID; }

  struct MapUser { sb_map_ctx* ctx; float device_ms; } mu{ctx, 0.0f};
  sb_batch_cb map_cb = [](void* user, const uint8_t* l, const uint8_t* r, uint32_t n, uint32_t L) -> int {
    MapUser* u = (MapUser*)user;
    sb_map_batch_stats st;
    const int rc = sb_map_batch(u->ctx, l, r, n, L, &st);
    if (rc == SB_OK) u->device_ms += st.device_ms;
    return rc;
  };
  sb_bucket_stats bs;
  int rc = sb_reads_bucketed(rd, mp.k, o.batch, o.max_read_len, o.threads, o.shard_index, o.shard_count, map_cb, &mu, &bs);
  sb_reads_close(rd);
  if (rc != SB_OK) { sb_map_destroy(ctx); return rc; }
  const float device_ms = mu.device_ms;
  const double t_map = now_s();

  // ---- classes -> EM -> outputs --------------------------------------------------------------------------------
  sb_map_result res;
  rc = sb_map_finish(ctx, &res);
  if (rc != SB_OK) { sb_map_destroy(ctx); return rc; }
  sb_eq_csr eq;
  eq.n_classes = res.n_classes; eq.n_txps = M; eq.off = res.off; eq.tids = res.tids; eq.weights = res.weights; eq.counts = res.counts;
  std::vector<double> alpha(M, 0.0);
  sb_em_stats est;
  memset(&est, 0, sizeof est);
  sb_em_ctx* em = sb_em_create(o.device);
  if (!em) { sb_map_destroy(ctx); return SB_ERR_CUDA; }
  rc = sb_em_optimize(em, &eq, &ep, res.projected_counts, res.eff_len, res.unique_counts, alpha.data(), &est);
  const double n_mapped = (double)res.n_mapped;
  std::string outs = out_dir ? out_dir : "";
  auto fail = [&](int code) { sb_em_destroy(em); sb_map_destroy(ctx); return code; };
  if (rc < 0) return fail(rc);
  if (rc == 1) { sb::set_error("The optimization algorithm failed (total alpha weight too small)"); return fail(SB_ERR_STATE); }
  const double t_em = now_s();
  if (!outs.empty()) {
    if (!make_dirs(outs + "/aux_info")) { sb::set_error("cannot create %s/aux_info", outs.c_str()); return fail(SB_ERR_INVALID); }
    std::vector<std::string> gen;
    std::vector<const char*> np;
    if (!names) {
      for (uint32_t t = 0; t < M; ++t) gen.push_back("t" + std::to_string(t));
      for (auto& s : gen) np.push_back(s.c_str());
      names = np.data();
    }
    std::vector<uint32_t> lens;
    if (!complete_len) {
      const uint64_t* off = nullptr;
      sb_index_host_arrays(ix, &off, nullptr, nullptr, nullptr, nullptr, nullptr);
      for (uint32_t t = 0; t < M; ++t) lens.push_back((uint32_t)(off[t + 1] - off[t]));
      complete_len = lens.data();
    }
    rc = sb_write_quant_sf((outs + "/quant.sf").c_str(), M, names, complete_len, res.eff_len, alpha.data(), n_mapped, 3);
    if (rc == SB_OK && (o.dump_eq || o.dump_eq_weights))
      rc = sb_write_eq_classes((outs + "/aux_info/eq_classes.txt.gz").c_str(), M, names, res.n_classes, res.off, res.tids,
                               o.dump_eq_weights ? res.weights : nullptr, res.counts);
    if (rc != SB_OK) return fail(rc);
    if (o.num_bootstraps || o.num_gibbs) {
      if (!make_dirs(outs + "/aux_info/bootstrap")) { sb::set_error("cannot create the bootstrap directory"); return fail(SB_ERR_INVALID); }
      BootUser bu{sb_bootstrap_writer_open((outs + "/aux_info/bootstrap/bootstraps.gz").c_str()), SB_OK};
      if (!bu.w) return fail(SB_ERR_INVALID);
      if (o.num_bootstraps) {
        sb_em_params bp = ep;
        bp.min_iter = 50;   // CollapsedEMOptimizer.cpp:411
        rc = sb_bootstrap(em, &bp, n_mapped, o.num_bootstraps, o.seed, boot_cb, &bu);
      } else {
        rc = sb_gibbs(em, alpha.data(), ep.use_vbem, ep.per_txp_prior, ep.vb_prior, o.num_gibbs, o.thinning ? o.thinning : 16,
                      o.no_gamma_draw, n_mapped, o.seed, boot_cb, &bu);
      }
      sb_bootstrap_writer_close(bu.w);
      if (rc < 0 || bu.rc != SB_OK) return fail(rc < 0 ? rc : bu.rc);
      // names of the columns of bootstraps.gz (GZipWriter writes names.tsv.gz next to it)
      std::string nm;
      for (uint32_t t = 0; t < M; ++t) { nm += names[t]; nm += (t + 1 < M) ? '\t' : '\n'; }
      gzFile g = gzopen((outs + "/aux_info/bootstrap/names.tsv.gz").c_str(), "wb");
      const bool wrote = g && gzwrite(g, nm.data(), (unsigned)nm.size()) == (int)nm.size();
      if (!g || (gzclose(g) != Z_OK) || !wrote) { sb::set_error("write error on %s/aux_info/bootstrap/names.tsv.gz", outs.c_str()); return fail(SB_ERR_INVALID); }
    }
  }
  if (alpha_out) memcpy(alpha_out, alpha.data(), (size_t)M * 8);
  if (sum) {
    memset(sum, 0, sizeof(*sum));
    sum->n_observed = bs.n_observed; sum->n_mapped = res.n_mapped; sum->n_too_short = bs.n_too_short;
    sum->n_trimmed_mates = bs.n_trimmed_mates;
    sum->n_classes = res.n_classes; sum->n_batches = bs.n_batches; sum->n_read_lengths = bs.n_read_lengths;
    sum->em_iters = est.iters; sum->em_converged = est.converged;
    sum->map_seconds = t_map - t0; sum->em_seconds = t_em - t_map; sum->total_seconds = now_s() - t0;
    sum->map_device_ms = device_ms;
  }
  sb_em_destroy(em);
  sb_map_destroy(ctx);
  return SB_OK;
}
