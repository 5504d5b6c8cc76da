// em_internal.h -- host-side state of the Stage-B optimiser context.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "../../include/salmon_b200.h"

namespace sb {

// A segmented matrix in HBM: rows with (index, weight) entries, plus the tile
// descriptors the streaming kernels walk.  Used twice: class-major
// (row = class, index = transcript id) and transcript-major (row = active
// transcript, index = compact class id).
struct SegMat {
  uint32_t n_rows = 0;
  uint32_t nnz = 0;
  uint32_t n_tiles = 0;
  uint32_t n_long = 0;
  uint32_t* off = nullptr;    // [n_rows+1]
  uint32_t* idx = nullptr;    // [nnz+16]
  double* w = nullptr;        // [nnz+16]
  uint4* tiles = nullptr;     // [n_tiles] {row0,row1,ent0,ent1}
  uint32_t* longs = nullptr;  // rows longer than LMAX
};

struct NcclApi;

}  // namespace sb

struct sb_em_ctx {
  int device = 0;
  int n_sm = 0;
  size_t l2_bytes = 0;
  cudaStream_t stream = nullptr;
  cudaEvent_t ev[4] = {nullptr, nullptr, nullptr, nullptr};

  // options
  int variant = 1;        // 1 = persistent cooperative kernel, 0 = one launch per phase
  int blocks_per_sm = 0;  // 0 = as many as fit
  int check_every = 1;

  // problem
  uint64_t C = 0, nnz = 0;
  uint32_t M = 0;
  double total_weight = 0.0;
  size_t h2d_bytes = 0;
  bool uploaded = false, prepared = false;
  sb_em_params params{};

  // uploaded inputs
  uint64_t* d_off = nullptr;
  uint32_t* d_tids = nullptr;
  double* d_aux = nullptr;
  uint64_t* d_counts = nullptr;
  double* d_projected = nullptr;
  double* d_eff_in = nullptr;
  uint64_t* d_unique = nullptr;

  // per transcript
  double* d_efflens = nullptr;
  double* d_prior = nullptr;
  double* d_alpha0 = nullptr;
  double* d_alpha = nullptr;
  double* d_theta = nullptr;
  double* d_base = nullptr;       // folded singleton classes
  uint32_t* d_tcnt = nullptr;
  uint32_t* d_tid_row = nullptr;
  uint32_t* d_row_tid = nullptr;

  // per class / entry
  double* d_cw = nullptr;         // combinedWeights in input order
  uint64_t* d_packed = nullptr;
  uint64_t* d_packed_scan = nullptr;
  uint8_t* d_valid = nullptr;
  double* d_cnt = nullptr;        // counts of compact classes (as f64)
  double* d_scale = nullptr;      // count/denom per compact class
  uint32_t* d_ent_cls = nullptr;
  uint32_t *d_sort_keys = nullptr, *d_sort_vals = nullptr, *d_sort_keys2 = nullptr,
           *d_sort_vals2 = nullptr;
  void* d_tmp = nullptr;
  size_t tmp_bytes = 0;

  sb::SegMat cm, tm;

  double* d_scalars = nullptr;    // 64 doubles of misc device scalars
  double* d_sum_partial = nullptr;
  uint32_t grid = 0;
  double sum0 = 0.0, inactive_sum = 0.0;

  // multi-GPU
  int rank = 0, nranks = 1;
  void* nccl_comm = nullptr;
  double* d_part = nullptr;       // per-transcript partial alpha' (send)
  double* d_part_red = nullptr;   // all-reduced (recv)

  // L2 flush
  void* d_flush = nullptr;
  size_t flush_bytes = 0;
  uint32_t flush_ctr = 0;

  // results
  uint32_t iters = 0, converged = 0, launches = 0;
  double max_rel_diff = 0.0;
  uint64_t n_degenerate = 0;
  float prepare_ms = 0, run_ms = 0;
};
