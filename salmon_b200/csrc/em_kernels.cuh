// em_kernels.cuh -- the EM / VBEM iteration kernels (sm_100a).
//
// One iteration = two segmented reductions over the class<->transcript map held
// twice in HBM (class-major and transcript-major), no atomics on the data path,
// fixed summation order:
//   P1 (class-major):  denom_c = sum_i theta[t_i] * w_ci ;  scale_c = count_c / denom_c
//   P2 (txp-major):    alpha'_t = base_t + theta_t * sum_c w_ct * scale_c, convergence
//                      test, theta'_t (VBEM: exp(digamma(alpha'+prior) - logNorm))
// Reference arithmetic: src/inference/CollapsedEMOptimizer.cpp:178-234 (EMUpdate_),
// :241-328 (VBEMUpdate_), :945-957 (convergence + swap).
//
// Layout: sliced ELL, slice height 32 (SELL-32).  Rows are ordered for gather locality
// (classes by first transcript id, transcripts by id) and then bucketed by length inside
// groups of SELL_GROUP rows, so the 32 rows of a slice have (nearly) equal length.  A
// warp owns a slice: lane = row, entry j of the 32 rows is one coalesced 128-byte (index)
// + 256-byte (weight) load, each lane accumulates its row sequentially in label order.
// Slices are dealt to warps in contiguous, work-balanced ranges computed at prepare time.
// Rows longer than LMAX are reduced by a whole block from their CSR copy.
#pragma once
#include <cooperative_groups.h>
#include <float.h>

#include "common.cuh"

namespace sb {
namespace cg = cooperative_groups;

constexpr int LMAX_DEFAULT = 96;             // rows longer than this leave the lane-per-row SELL path
constexpr int LWARP = 2048;                  // ... and are reduced by one warp (<= LWARP) or one block
constexpr int SELL_GROUP = 1024;             // rows per length-bucketing group
constexpr int EM_THREADS = 256;
constexpr double DIGAMMA_MIN = 1e-10;        // CollapsedEMOptimizer.cpp:43
constexpr double MIN_EQ_W = DBL_MIN;         // :40
constexpr double ALPHA_CHECK_CUTOFF = 1e-2;  // :884
constexpr uint16_t LEN_LONG = 0xFFFFu;       // row handled by the block path

// One segmented matrix in SELL-32 form (+ CSR copy of the long rows only).
struct Sell {
  const uint32_t* slice_ptr;   // [n_slices+1] first column of each slice
  const uint16_t* len;         // [n_rows] entries per row (LEN_LONG: block path)
  const uint32_t* idx;         // [n_cols*32] gather index, column-interleaved
  const double* w;             // [n_cols*32]
  const uint32_t* warp_begin;  // [n_warps+1] slice range per warp (work balanced)
  // long rows: (row, first entry, end entry) triples into csr_idx / csr_w
  const uint32_t* long_rows;   // [3*n_long]
  const uint32_t* csr_idx;
  const double* csr_w;
  uint32_t n_rows, n_slices, n_long;
  uint32_t n_block;            // the first n_block long rows (longest first) take the block path
  // dynamic distribution (k_em_persistent_dyn): the slices of a block cut into units of <= CH columns (whole slices;
  // a slice wider than CH is a unit of its own), handed out to the block's warps through a shared-memory counter
  const uint4* units;          // {first slice, end slice, first column, end column}, per block, costliest first
  const uint32_t* blk_unit_ptr;// [grid+1]
  uint32_t keep_pct;           // % of stream chunks loaded with L2 evict_last (rest evict_first)
};

struct EmArgs {
  Sell cm;                      // rows = kept multi-transcript classes; idx = state index
  Sell tm;                      // rows = active transcripts;           idx = class id
  const double* c_cnt;          // [classes] count as f64
  double* scale;                // [classes] count / denom
  double* raw1; double* raw2;   // batched streaming (run_phase_b): per-row sums of the SELL path, [rows of cm] / [rows of tm]
  // iteration state.  Single GPU: indexed by ROW of tm (cm.idx holds rows).  Multi GPU:
  // indexed by transcript id (cm.idx holds ids) and row_tid maps tm rows to ids.
  double* alpha; double* theta; const double* prior; const double* base;
  const uint32_t* row_tid;
  double* part_out;             // multi-GPU: this rank's alpha' share per transcript id
  // reductions
  double* sum_partial;          // [2][grid]
  unsigned long long* maxrel;   // [2] bit pattern of a non-negative double
  double inactive_sum; double sum0;
  double tol;
  double min_eq_w;              // denominator guard: DBL_MIN (optimize) / denorm_min (serial EM)
  double first_bias;            // 1.0 for optimize's first plain-EM iteration (:812,:821), else 0
  uint32_t min_iter, max_iter;
  int vbem;
  uint32_t* out;                // [0]=iters [1]=converged [2]=maxrel slot
  unsigned long long* dbg;      // optional [n_warps*8] phase timestamps (ns) of iteration dbg_it
  uint32_t dbg_it;
  // multi-GPU, fused all-reduce over peer memory (k_em_persistent_mgpu): every rank owns one exchange block
  //   [ flags: 64 u64 | part: M doubles | red: M doubles ]   mapped into every peer (CUDA IPC, NVLink P2P)
  double* const* peers;         // [nranks] base pointers of the exchange blocks (peers[rank] = own)
  uint32_t rank, nranks, M;
  unsigned long long epoch0;    // barrier epochs consumed by earlier launches
  uint32_t* xfail;              // set when a peer did not show up in time
};

__device__ __forceinline__ unsigned long long gtime_ns() {
  unsigned long long t;
  asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
  return t;
}

struct P2Acc {
  double sum;     // sum of (alpha' + prior) over my rows
  double maxrel;  // max rel diff over my rows
};

// Per-warp TMA ring: each warp streams ITS contiguous column range of the SELL arrays
// through a private double-buffered shared-memory ring with 1-D bulk copies (lane 0 is
// the producer, the warp is the consumer), so the index/weight stream is never a
// dependent load and no block-level barrier exists on the data path.
constexpr int RING = 2;   // chunks in flight per warp
template <int CH>         // CH = columns (x32 entries) per chunk
struct __align__(128) WarpRing {
  double w[RING][CH * 32];
  uint32_t idx[RING][CH * 32];
};
constexpr int EM_WARPS = EM_THREADS / 32;
template <int CH>
__host__ __device__ constexpr size_t em_smem() { return sizeof(WarpRing<CH>) * EM_WARPS + EM_WARPS * RING * 8 + 40 * 8; }

template <int CH>
struct WarpCtx {
  WarpRing<CH>* ring;
  uint64_t* bars;        // [RING]
  uint32_t phase_bits;   // mbarrier parity per stage
  double* scratch;       // block scratch (32 doubles)
  unsigned long long* dbg;  // optional: timestamp after the SELL part of a phase
};

template <int CH>
__device__ __forceinline__ void warp_setup(WarpCtx<CH>& W, unsigned char* smem) {
  const uint32_t wid = threadIdx.x >> 5;
  W.ring = reinterpret_cast<WarpRing<CH>*>(smem) + wid;
  W.bars = reinterpret_cast<uint64_t*>(smem + sizeof(WarpRing<CH>) * EM_WARPS) + wid * RING;
  W.scratch = reinterpret_cast<double*>(smem + sizeof(WarpRing<CH>) * EM_WARPS + EM_WARPS * RING * 8);
  W.phase_bits = 0;
  W.dbg = nullptr;
  if ((threadIdx.x & 31u) == 0) {
#pragma unroll
    for (int s = 0; s < RING; ++s) mbar_init(&W.bars[s], 1);
    mbar_fence_init();
  }
  __syncthreads();
}

// operands of a row's epilogue, fetched while the row's columns stream
struct RowOps {
  double x0, x1, x2, x3;
  uint32_t len;
};
template <int PHASE>
__device__ __forceinline__ RowOps load_ops(const EmArgs& A, const Sell& S, uint32_t row) {
  RowOps o;
  o.x0 = o.x1 = o.x2 = o.x3 = 0.0;
  o.len = LEN_LONG;
  if (row < S.n_rows) {
    o.len = __ldg(&S.len[row]);
    if (PHASE == 1) {
      o.x0 = __ldg(&A.c_cnt[row]);
    } else if (PHASE == 2) {
      o.x0 = A.theta[row];
      o.x1 = __ldg(&A.prior[row]);
      o.x2 = __ldg(&A.base[row]);
      o.x3 = A.alpha[row];
    }
  }
  return o;
}

template <int PHASE>
__device__ __forceinline__ void row_finish(const EmArgs& A, uint32_t row, const RowOps& o, double acc,
                                           double logNorm, double bias, P2Acc& pa) {
  if (o.len == LEN_LONG) return;  // beyond the last row, or a long row (block path)
  if (PHASE == 1) {
    A.scale[row] = (acc <= A.min_eq_w) ? 0.0 : o.x0 / acc;
  } else if (PHASE == 2) {
    const double th = o.x0, pr = o.x1;
    double na = o.x2 + bias;
    if (th > 0.0) na += th * acc;
    if (na > ALPHA_CHECK_CUTOFF) pa.maxrel = fmax(pa.maxrel, fabs(o.x3 - na) / na);
    A.alpha[row] = na;
    const double ap = na + pr;
    pa.sum += ap;
    A.theta[row] = A.vbem ? ((ap > DIGAMMA_MIN) ? exp(digamma_pos(ap) - logNorm) : 0.0) : na;
  } else {
    const uint32_t t = __ldg(&A.row_tid[row]);
    const double th = A.theta[t];
    double na = __ldg(&A.base[t]);
    if (th > 0.0) na += th * acc;
    A.part_out[t] = na;
  }
}

// static stream range of one warp in one matrix (the matrices never change)
struct WarpRange {
  uint32_t s0, s1, cbeg, cend;
};
__device__ __forceinline__ WarpRange load_range(const Sell& S, uint32_t gwarp) {
  WarpRange r;
  r.s0 = __ldg(&S.warp_begin[gwarp]);
  r.s1 = __ldg(&S.warp_begin[gwarp + 1]);
  r.cbeg = __ldg(&S.slice_ptr[r.s0]);
  r.cend = __ldg(&S.slice_ptr[r.s1]);
  return r;
}
template <int CH>
__device__ __forceinline__ void ring_issue(const Sell& S, WarpCtx<CH>& W, const WarpRange& R, uint32_t k) {
  if ((threadIdx.x & 31u) == 0) {
    const uint32_t c = R.cbeg + k * CH;
    const uint32_t cols = min((uint32_t)CH, R.cend - c);
    const int st = k % RING;
    // The two layouts together exceed what the L2 keeps under a cyclic sweep; pin a fixed
    // pseudo-random subset of chunks (evict_last) and let the rest stream (evict_first).
    const bool keep = (((c / CH) * 2654435761u) >> 16) % 100u < S.keep_pct;
    const uint64_t pol = keep ? l2_policy_evict_last() : l2_policy_evict_first();
    mbar_arrive_expect_tx(&W.bars[st], cols * 384u);
    bulk_g2s_hint(W.ring->w[st], S.w + (size_t)c * 32u, cols * 256u, &W.bars[st], pol);
    bulk_g2s_hint(W.ring->idx[st], S.idx + (size_t)c * 32u, cols * 128u, &W.bars[st], pol);
  }
}
// fill the ring with the first chunks of a phase.  The matrices are read-only, so this
// may run BEFORE the grid barrier that precedes the phase: the stream then lands while
// the grid synchronises and is never on the critical path.
template <int CH>
__device__ __forceinline__ void ring_prefetch(const Sell& S, WarpCtx<CH>& W, const WarpRange& R) {
  const uint32_t nchunks = (R.cend - R.cbeg + CH - 1) / CH;
#pragma unroll
  for (int k = 0; k < RING; ++k)
    if ((uint32_t)k < nchunks) ring_issue(S, W, R, k);
}

template <int PHASE, int CH>
__device__ __forceinline__ void run_phase(const EmArgs& A, WarpCtx<CH>& W, const WarpRange& R,
                                          uint32_t bid, uint32_t nblk, double logNorm, double bias,
                                          P2Acc& pa) {
  const Sell& S = (PHASE == 1) ? A.cm : A.tm;
  // theta / scale are rewritten by other blocks inside the persistent kernel: plain
  // coherent loads only, never ld.global.nc.
  const double* gsrc = (PHASE == 1) ? A.theta : A.scale;
  const bool em_nan_guard = (PHASE == 1) && !A.vbem;
  const uint32_t lane = threadIdx.x & 31u;
  const uint32_t s0 = R.s0, s1 = R.s1;
  if (s1 > s0) {
    const uint32_t cbeg = R.cbeg, cend = R.cend;
    const uint32_t nchunks = (cend - cbeg + CH - 1) / CH;
    auto issue = [&](uint32_t k) { ring_issue(S, W, R, k); };
    // slice boundaries, 32 at a time: lane l holds the end column of slice sbase+l
    uint32_t sbase = s0;
    uint32_t sp = (s0 + lane < s1) ? __ldg(&S.slice_ptr[s0 + lane + 1]) : cend;
    uint32_t s = s0;
    uint32_t slice_end = __shfl_sync(0xffffffffu, sp, 0);
    RowOps ops = load_ops<PHASE>(A, S, s * 32u + lane);
    double acc = 0.0;
    auto next_slice = [&]() {
      row_finish<PHASE>(A, s * 32u + lane, ops, acc, logNorm, bias, pa);
      ++s;
      acc = 0.0;
      if (s < s1) {
        if (s - sbase == 32u) {
          sbase = s;
          sp = (s + lane < s1) ? __ldg(&S.slice_ptr[s + lane + 1]) : cend;
        }
        slice_end = __shfl_sync(0xffffffffu, sp, (int)(s - sbase));
        ops = load_ops<PHASE>(A, S, s * 32u + lane);
      }
    };
    for (uint32_t k = 0; k < nchunks; ++k) {
      const int st = k % RING;
      mbar_wait(&W.bars[st], (W.phase_bits >> st) & 1u);
      W.phase_bits ^= (1u << st);
      const uint32_t* sidx = W.ring->idx[st] + lane;
      const double* sw = W.ring->w[st] + lane;
      const uint32_t c0 = cbeg + k * CH;
      const uint32_t cstop = min(cend, c0 + CH);
      uint32_t col = c0;
      while (col < cstop) {
        while (col == slice_end && s < s1) next_slice();  // (possibly zero-width slices)
        const uint32_t n = min(slice_end, cstop) - col;
        const uint32_t l0 = col - c0;
        uint32_t j = 0;
        for (; j + 4 <= n; j += 4) {
          const uint32_t o = (l0 + j) * 32u;
          const double g0 = gsrc[sidx[o]], g1 = gsrc[sidx[o + 32]];
          const double g2 = gsrc[sidx[o + 64]], g3 = gsrc[sidx[o + 96]];
          double v0 = g0 * sw[o], v1 = g1 * sw[o + 32], v2 = g2 * sw[o + 64], v3 = g3 * sw[o + 96];
          if (em_nan_guard) {
            if (isnan(v0)) v0 = 0.0;
            if (isnan(v1)) v1 = 0.0;
            if (isnan(v2)) v2 = 0.0;
            if (isnan(v3)) v3 = 0.0;
          }
          acc += v0; acc += v1; acc += v2; acc += v3;
        }
        for (; j < n; ++j) {
          const uint32_t o = (l0 + j) * 32u;
          double v = gsrc[sidx[o]] * sw[o];
          if (em_nan_guard && isnan(v)) v = 0.0;
          acc += v;
        }
        col += n;
      }
      __syncwarp();
      if (k + RING < nchunks) issue(k + RING);
    }
    while (s < s1) next_slice();
  }
  if (W.dbg && lane == 0) *W.dbg = gtime_ns();
  // very long rows: whole block per row, fixed-order tree reduction
  for (uint32_t li = bid; li < S.n_block; li += nblk) {
    const uint32_t r = __ldg(&S.long_rows[3 * li]);
    const uint32_t b = __ldg(&S.long_rows[3 * li + 1]);
    const uint32_t e = __ldg(&S.long_rows[3 * li + 2]);
    double acc = 0.0;
    for (uint32_t k = b + threadIdx.x; k < e; k += EM_THREADS) {
      double v = gsrc[__ldg(&S.csr_idx[k])] * __ldg(&S.csr_w[k]);
      if (em_nan_guard && isnan(v)) v = 0.0;
      acc += v;
    }
    acc = block_reduce<false>(acc, W.scratch);
    if (threadIdx.x == 0) {
      RowOps o = load_ops<PHASE>(A, S, r);
      o.len = 0;  // force the epilogue for this long row
      row_finish<PHASE>(A, r, o, acc, logNorm, bias, pa);
    }
    __syncthreads();
  }
  // long rows (LMAX < len <= LWARP): one warp per row, lanes stride the CSR copy, fixed
  // shuffle tree.  Sorted longest-first and dealt round-robin over all warps of the grid.
  {
    const uint32_t gw = bid * EM_WARPS + (threadIdx.x >> 5);
    const uint32_t nw = nblk * EM_WARPS;
    // lane k keeps the sum of the k-th row this warp reduced; the epilogues (digamma, exp)
    // then run lane-parallel, 32 rows at a time.
    uint32_t cnt = 0, myrow = 0xffffffffu;
    double myacc = 0.0;
    auto flush = [&]() {
      if (myrow != 0xffffffffu) {
        RowOps o = load_ops<PHASE>(A, S, myrow);
        o.len = 0;  // force the epilogue for a long row
        row_finish<PHASE>(A, myrow, o, myacc, logNorm, bias, pa);
      }
      myrow = 0xffffffffu;
      cnt = 0;
    };
    for (uint32_t li = S.n_block + gw; li < S.n_long; li += nw) {
      const uint32_t r = __ldg(&S.long_rows[3 * li]);
      const uint32_t b = __ldg(&S.long_rows[3 * li + 1]);
      const uint32_t e = __ldg(&S.long_rows[3 * li + 2]);
      double a0 = 0.0, a1 = 0.0;
      uint32_t k = b + lane;
      for (; k + 32 < e; k += 64) {
        const uint32_t i0 = __ldg(&S.csr_idx[k]), i1 = __ldg(&S.csr_idx[k + 32]);
        double v0 = gsrc[i0] * __ldg(&S.csr_w[k]);
        double v1 = gsrc[i1] * __ldg(&S.csr_w[k + 32]);
        if (em_nan_guard) {
          if (isnan(v0)) v0 = 0.0;
          if (isnan(v1)) v1 = 0.0;
        }
        a0 += v0;
        a1 += v1;
      }
      if (k < e) {
        double v = gsrc[__ldg(&S.csr_idx[k])] * __ldg(&S.csr_w[k]);
        if (em_nan_guard && isnan(v)) v = 0.0;
        a0 += v;
      }
      const double acc = warp_sum(a0 + a1);
      if (lane == cnt) { myacc = acc; myrow = r; }
      if (++cnt == 32) flush();
    }
    flush();
  }
}


// The SELL stream of one warp's slice range: raw[row] = sum_j gsrc[idx_j] * w_j in column order (see run_phase_b).
template <int CH, int NB, bool GUARD>
__device__ __forceinline__ void stream_sell(const Sell& S, WarpCtx<CH>& W, const WarpRange& R,
                                            const double* gsrc, double* raw) {
  const uint32_t lane = threadIdx.x & 31u;
  const uint32_t s0 = R.s0, s1 = R.s1;
  const uint32_t cbeg = R.cbeg, cend = R.cend;
  const uint32_t nchunks = (cend - cbeg + CH - 1) / CH;
  const uint32_t n_rows = S.n_rows;
  uint32_t sbase = s0;
  uint32_t sp = (s0 + lane < s1) ? __ldg(&S.slice_ptr[s0 + lane + 1]) : cend;
  uint32_t s = s0;
  uint32_t slice_end = __shfl_sync(0xffffffffu, sp, 0);
  uint32_t col = cbeg;
  double acc = 0.0;
  auto boundary = [&]() {   // slice s is complete
    const uint32_t row = s * 32u + lane;
    if (row < n_rows) raw[row] = acc;
    acc = 0.0;
    ++s;
    if (s < s1) {
      if (s - sbase == 32u) {
        sbase = s;
        sp = (s + lane < s1) ? __ldg(&S.slice_ptr[s + lane + 1]) : cend;
      }
      slice_end = __shfl_sync(0xffffffffu, sp, (int)(s - sbase));
    }
  };
  for (uint32_t k = 0; k < nchunks; ++k) {
    const int st = k % RING;
    mbar_wait(&W.bars[st], (W.phase_bits >> st) & 1u);
    W.phase_bits ^= (1u << st);
    const uint32_t* sidx = W.ring->idx[st] + lane;
    const double* sw = W.ring->w[st] + lane;
    const uint32_t cstop = min(cend, cbeg + (k + 1) * CH);
    auto step = [&](double gj, double wj) {
      while (col == slice_end && s < s1) boundary();   // warp-uniform (possibly zero-width slices)
      double v = gj * wj;
      if (GUARD && isnan(v)) v = 0.0;
      acc += v;
      ++col;
    };
#pragma unroll 1
    for (uint32_t l0 = 0; col < cstop; l0 += NB * 32u) {
      const uint32_t nb = cstop - col;   // columns left in this chunk (>= 1)
      double g[NB];
      if (nb >= (uint32_t)NB) {
#pragma unroll
        for (int j = 0; j < NB; ++j) g[j] = gsrc[sidx[l0 + j * 32]];
#pragma unroll
        for (int j = 0; j < NB; ++j) step(g[j], sw[l0 + j * 32]);
      } else {
#pragma unroll
        for (int j = 0; j < NB; ++j) g[j] = ((uint32_t)j < nb) ? gsrc[sidx[l0 + j * 32]] : 0.0;
#pragma unroll
        for (int j = 0; j < NB; ++j)
          if ((uint32_t)j < nb) step(g[j], sw[l0 + j * 32]);
      }
    }
    __syncwarp();
    if (k + RING < nchunks) ring_issue(S, W, R, k + RING);
  }
  while (s < s1) boundary();
}

// ---- batched streaming (MODE 1) ----------------------------------------------------------------------------------
// The lane-per-row loop above is bound by the latency of the gathers: a warp waits one L2 round trip per group of
// columns it has in flight, and with short slices (a class has ~6 members) that is 2-3 round trips per slice
// (ncu r1: 42 % of the warp-time at the grid barriers, long-scoreboard on the first use of every gather group, and
// per-warp phase times that match  #groups x L2 latency).  Here the gathers of NB consecutive COLUMNS of the chunk
// are issued together, whatever slices they belong to (the indices are already in shared memory); the slice
// boundaries are handled while the values are consumed.  The per-row epilogue (count/denominator; digamma/exp) is
// taken out of the stream: a boundary only stores the row's sum, and after the stream the warp finishes its rows
// lane-parallel with all operand loads independent.  Summation order per row is unchanged -> same bits as MODE 0.
template <int PHASE, int CH, int NB>
__device__ __forceinline__ void run_phase_b(const EmArgs& A, WarpCtx<CH>& W, const WarpRange& R,
                                            uint32_t bid, uint32_t nblk, double logNorm, double bias,
                                            P2Acc& pa) {
  static_assert(CH % NB == 0, "batch must divide the chunk");
  const Sell& S = (PHASE == 1) ? A.cm : A.tm;
  const double* gsrc = (PHASE == 1) ? A.theta : A.scale;   // rewritten by other blocks: coherent loads only
  double* raw = (PHASE == 1) ? A.raw1 : A.raw2;
  const bool em_nan_guard = (PHASE == 1) && !A.vbem;
  const uint32_t lane = threadIdx.x & 31u;
  const uint32_t s0 = R.s0, s1 = R.s1;
  // long rows first: they are the longest single items of a phase (critical path), the stream's first chunks are
  // already in flight (ring_prefetch) and land meanwhile.
  // very long rows: whole block per row, fixed-order tree reduction
  for (uint32_t li = bid; li < S.n_block; li += nblk) {
    const uint32_t r = __ldg(&S.long_rows[3 * li]);
    const uint32_t b = __ldg(&S.long_rows[3 * li + 1]);
    const uint32_t e = __ldg(&S.long_rows[3 * li + 2]);
    double a0 = 0.0, a1 = 0.0;
    uint32_t k = b + threadIdx.x;
    for (; k + EM_THREADS < e; k += 2 * EM_THREADS) {
      const uint32_t i0 = __ldg(&S.csr_idx[k]), i1 = __ldg(&S.csr_idx[k + EM_THREADS]);
      double v0 = gsrc[i0] * __ldg(&S.csr_w[k]);
      double v1 = gsrc[i1] * __ldg(&S.csr_w[k + EM_THREADS]);
      if (em_nan_guard) {
        if (isnan(v0)) v0 = 0.0;
        if (isnan(v1)) v1 = 0.0;
      }
      a0 += v0;
      a1 += v1;
    }
    if (k < e) {
      double v = gsrc[__ldg(&S.csr_idx[k])] * __ldg(&S.csr_w[k]);
      if (em_nan_guard && isnan(v)) v = 0.0;
      a0 += v;
    }
    const double acc = block_reduce<false>(a0 + a1, W.scratch);
    if (threadIdx.x == 0) {
      RowOps o = load_ops<PHASE>(A, S, r);
      o.len = 0;  // force the epilogue for this long row
      row_finish<PHASE>(A, r, o, acc, logNorm, bias, pa);
    }
    __syncthreads();
  }
  // long rows (LMAX < len <= lwarp): one warp per row, 4 independent gathers per lane in flight
  {
    const uint32_t gw = bid * EM_WARPS + (threadIdx.x >> 5);
    const uint32_t nw = nblk * EM_WARPS;
    uint32_t cnt = 0, myrow = 0xffffffffu;
    double myacc = 0.0;
    auto flush = [&]() {
      if (myrow != 0xffffffffu) {
        RowOps o = load_ops<PHASE>(A, S, myrow);
        o.len = 0;
        row_finish<PHASE>(A, myrow, o, myacc, logNorm, bias, pa);
      }
      myrow = 0xffffffffu;
      cnt = 0;
    };
    for (uint32_t li = S.n_block + gw; li < S.n_long; li += nw) {
      const uint32_t r = __ldg(&S.long_rows[3 * li]);
      const uint32_t b = __ldg(&S.long_rows[3 * li + 1]);
      const uint32_t e = __ldg(&S.long_rows[3 * li + 2]);
      double a0 = 0.0, a1 = 0.0, a2 = 0.0, a3 = 0.0;
      uint32_t k = b + lane;
      for (; k + 96 < e; k += 128) {
        const uint32_t i0 = __ldg(&S.csr_idx[k]), i1 = __ldg(&S.csr_idx[k + 32]);
        const uint32_t i2 = __ldg(&S.csr_idx[k + 64]), i3 = __ldg(&S.csr_idx[k + 96]);
        const double g0 = gsrc[i0], g1 = gsrc[i1], g2 = gsrc[i2], g3 = gsrc[i3];
        double v0 = g0 * __ldg(&S.csr_w[k]), v1 = g1 * __ldg(&S.csr_w[k + 32]);
        double v2 = g2 * __ldg(&S.csr_w[k + 64]), v3 = g3 * __ldg(&S.csr_w[k + 96]);
        if (em_nan_guard) {
          if (isnan(v0)) v0 = 0.0;
          if (isnan(v1)) v1 = 0.0;
          if (isnan(v2)) v2 = 0.0;
          if (isnan(v3)) v3 = 0.0;
        }
        a0 += v0; a1 += v1; a2 += v2; a3 += v3;
      }
      for (; k < e; k += 32) {
        double v = gsrc[__ldg(&S.csr_idx[k])] * __ldg(&S.csr_w[k]);
        if (em_nan_guard && isnan(v)) v = 0.0;
        a0 += v;
      }
      const double acc = warp_sum((a0 + a1) + (a2 + a3));
      if (lane == cnt) { myacc = acc; myrow = r; }
      if (++cnt == 32) flush();
    }
    flush();
  }
  if (s1 > s0) {
    if (em_nan_guard) stream_sell<CH, NB, true>(S, W, R, gsrc, raw);
    else stream_sell<CH, NB, false>(S, W, R, gsrc, raw);
    if (W.dbg && lane == 0) *W.dbg = gtime_ns();
    // epilogues of my rows, two slices at a time (operand loads of both in flight together)
    for (uint32_t q = s0; q < s1; q += 2) {
      const uint32_t r0 = q * 32u + lane, r1 = r0 + 32u;
      const bool has1 = (q + 1 < s1);
      RowOps o0 = load_ops<PHASE>(A, S, r0);
      RowOps o1;
      o1.x0 = o1.x1 = o1.x2 = o1.x3 = 0.0;
      o1.len = LEN_LONG;
      if (has1) o1 = load_ops<PHASE>(A, S, r1);
      const double v0 = (r0 < S.n_rows) ? raw[r0] : 0.0;
      const double v1 = (has1 && r1 < S.n_rows) ? raw[r1] : 0.0;
      row_finish<PHASE>(A, r0, o0, v0, logNorm, bias, pa);
      row_finish<PHASE>(A, r1, o1, v1, logNorm, bias, pa);
    }
  } else if (W.dbg && lane == 0) {
    *W.dbg = gtime_ns();
  }
}

template <int PHASE, int CH, int MODE>
__device__ __forceinline__ void run_phase_m(const EmArgs& A, WarpCtx<CH>& W, const WarpRange& R,
                                            uint32_t bid, uint32_t nblk, double logNorm, double bias,
                                            P2Acc& pa) {
  if constexpr (MODE == 0) run_phase<PHASE, CH>(A, W, R, bid, nblk, logNorm, bias, pa);
  else run_phase_b<PHASE, CH, (MODE == 2 && CH % 16 == 0) ? 16 : 8>(A, W, R, bid, nblk, logNorm, bias, pa);
}

// alphaSum of the iteration input, from the per-block partials of the previous P2
__device__ __forceinline__ double sum_partials(const double* part, uint32_t n, double extra,
                                               double* scratch) {
  double acc = 0.0;
  for (uint32_t i = threadIdx.x; i < n; i += blockDim.x) acc += __ldcg(&part[i]);
  acc = block_reduce<false>(acc, scratch);
  return acc + extra;
}

// lagged logNorm: alphaSum of THIS iteration's input = the per-block partials written by
// the previous P2 (complete since the last grid barrier / kernel boundary).  Any common
// factor in theta cancels in P1/P2.  Warp 0 only; result in scratch[33].
__device__ __forceinline__ void lag_lognorm_warp0(const EmArgs& A, uint32_t par, uint32_t nblk,
                                                  double* scratch) {
  if (threadIdx.x < 32) {
    const double* part = A.sum_partial + (size_t)(par ^ 1u) * nblk;
    double acc = 0.0;
    for (uint32_t i = threadIdx.x; i < nblk; i += 32) acc += __ldcg(&part[i]);
    acc = warp_sum(acc);
    if (threadIdx.x == 0) scratch[33] = digamma_pos(acc + A.inactive_sum);
  }
}

__device__ __forceinline__ void p2_finish(const EmArgs& A, double* scratch, P2Acc& pa,
                                          uint32_t par) {
  double bs = block_reduce<false>(pa.sum, scratch);
  double bm = block_reduce<true>(pa.maxrel, scratch);
  if (threadIdx.x == 0) {
    A.sum_partial[(size_t)par * gridDim.x + blockIdx.x] = bs;
    if (bm > 0.0) atomicMax(&A.maxrel[par], (unsigned long long)__double_as_longlong(bm));
  }
}

#define SB_DBG(slot)                                                        \
  if (A.dbg && it == A.dbg_it && (threadIdx.x & 31u) == 0) A.dbg[(size_t)gwarp * 8 + (slot)] = gtime_ns();

// ---- persistent cooperative kernel: the whole iteration loop, two grid barriers/iter
template <int CH, int MINB, int MODE>
__global__ void __launch_bounds__(EM_THREADS, MINB) k_em_persistent(const __grid_constant__ EmArgs A) {
  extern __shared__ __align__(128) unsigned char smem[];
  WarpCtx<CH> W;
  warp_setup(W, smem);
  double* scratch = W.scratch;
  cg::grid_group grid = cg::this_grid();
  const uint32_t bid = blockIdx.x, nblk = gridDim.x;
  const uint32_t gwarp = bid * (EM_THREADS / 32) + (threadIdx.x >> 5);
  const WarpRange R1 = load_range(A.cm, gwarp);
  const WarpRange R2 = load_range(A.tm, gwarp);
  uint32_t it = 0;
  bool converged = false;
  double logNorm = A.vbem ? digamma_pos(A.sum0) : 0.0;
  ring_prefetch(A.cm, W, R1);
  while (it < A.min_iter || (it < A.max_iter && !converged)) {
    const uint32_t par = it & 1u;
    if (bid == 0 && threadIdx.x == 0) A.maxrel[par] = 0ull;
    if (A.vbem && it > 0) lag_lognorm_warp0(A, par, nblk, scratch);  // consumed after the next barrier
    P2Acc pa{0.0, 0.0};
    SB_DBG(0)
    run_phase_m<1, CH, MODE>(A, W, R1, bid, nblk, 0.0, 0.0, pa);
    SB_DBG(1)
    ring_prefetch(A.tm, W, R2);   // P2's stream lands during the grid barrier
    grid.sync();
    SB_DBG(2)
    if (A.vbem && it > 0) logNorm = scratch[33];   // written before the grid barrier above
    const double bias = (it == 0) ? A.first_bias : 0.0;  // alphasPrime starts at 1.0 (:812,:821)
    SB_DBG(3)
    W.dbg = (A.dbg && it == A.dbg_it) ? &A.dbg[(size_t)gwarp * 8 + 7] : nullptr;
    run_phase_m<2, CH, MODE>(A, W, R2, bid, nblk, logNorm, bias, pa);
    W.dbg = nullptr;
    SB_DBG(4)
    ring_prefetch(A.cm, W, R1);   // next iteration's P1 stream (harmless if the loop ends)
    p2_finish(A, scratch, pa, par);
    SB_DBG(5)
    grid.sync();
    SB_DBG(6)
    const double mr = __longlong_as_double((long long)__ldcg(&A.maxrel[par]));
    converged = !(mr > A.tol);
    ++it;
  }
  if (bid == 0 && threadIdx.x == 0) {
    A.out[0] = it;
    A.out[1] = converged ? 1u : 0u;
    A.out[2] = (it - 1) & 1u;
  }
  // drain the speculative prefetch before the block (and its shared memory) retires
  {
    const uint32_t nchunks = (R1.cend - R1.cbeg + CH - 1) / CH;
#pragma unroll
    for (int k = 0; k < RING; ++k)
      if ((uint32_t)k < nchunks) mbar_wait(&W.bars[k], (W.phase_bits >> k) & 1u);
  }
}

// ---- dynamic distribution inside a block (MODE 3, k_em_persistent_dyn) ---------------------------------------------
// Round-1 timeline: the median warp finishes P1 in 6.4 us and P2 in 11 us, the slowest in 10.7 / 17.6 us, and 42 % of
// all warp-time is spent at the two grid barriers.  The per-warp time is the latency of the warp's own instruction
// stream, so its spread is latency noise on a static share.  Here a block's slices are cut (host side, at prepare)
// into units of <= CH columns and the block's 8 warps take units from a shared-memory counter until none is left: the
// static share becomes the block's, the variance averages over 8 warps x ~10 units.  A warp keeps its private ring;
// the chunk stream it feeds is "the chunks of the units I grabbed", issued RING chunks ahead of consumption (the unit
// after the current one is grabbed, and its bulk copies issued, while the current one is consumed; the first units of
// the next phase before the grid barrier).  Per-row arithmetic and summation order are those of MODE 0.
constexpr int UQ = 4;               // units in flight per warp (consumed + issued ahead), >= RING + 1
struct __align__(16) DynWarp {
  uint4 uq[UQ];                     // unit descriptors, ring indexed by units grabbed
  uint32_t sp[UQ][32];              // end column of each slice of the unit
  uint32_t cq[RING][2];             // column range of the chunk in each ring slot
};
template <int CH>
__host__ __device__ constexpr size_t em_smem_dyn() { return em_smem<CH>() + sizeof(DynWarp) * EM_WARPS + 16; }

struct DynState {                   // warp-uniform
  uint32_t ki = 0, kc = 0;          // chunks issued / consumed
  uint32_t i_col = 0, i_end = 0;    // issue side: next column, end column of the unit being issued
  uint32_t un_in = 0, un_out = 0;   // units grabbed / taken over by the consumer
  bool i_done = true;               // no more units in this phase
};

// issue bulk copies until RING chunks are in flight or the block has no more units
template <int CH>
__device__ __forceinline__ void dyn_issue(const Sell& S, WarpCtx<CH>& W, DynWarp* D, DynState& st, uint32_t* ctr,
                                          uint32_t ub0, uint32_t ub1) {
  const uint32_t lane = threadIdx.x & 31u;
  while (st.ki - st.kc < (uint32_t)RING) {
    if (st.i_col == st.i_end) {
      if (st.i_done || st.un_in - st.un_out >= (uint32_t)UQ) break;
      uint32_t u = 0;
      if (lane == 0) u = ub0 + atomicAdd(ctr, 1u);
      u = __shfl_sync(0xffffffffu, u, 0);
      if (u >= ub1) { st.i_done = true; break; }
      const uint4 d = __ldg(&S.units[u]);
      const uint32_t q = st.un_in % UQ;
      if (lane == 0) D->uq[q] = d;
      D->sp[q][lane] = (d.x + lane < d.y) ? __ldg(&S.slice_ptr[d.x + lane + 1]) : d.w;
      ++st.un_in;
      st.i_col = d.z; st.i_end = d.w;
    }
    const uint32_t cols = min((uint32_t)CH, st.i_end - st.i_col);
    const int slot = st.ki % RING;
    if (lane == 0) {
      const uint32_t c = st.i_col;
      const bool keep = (((c / CH) * 2654435761u) >> 16) % 100u < S.keep_pct;
      const uint64_t pol = keep ? l2_policy_evict_last() : l2_policy_evict_first();
      D->cq[slot][0] = c; D->cq[slot][1] = c + cols;
      mbar_arrive_expect_tx(&W.bars[slot], cols * 384u);
      bulk_g2s_hint(W.ring->w[slot], S.w + (size_t)c * 32u, cols * 256u, &W.bars[slot], pol);
      bulk_g2s_hint(W.ring->idx[slot], S.idx + (size_t)c * 32u, cols * 128u, &W.bars[slot], pol);
    }
    st.i_col += cols;
    ++st.ki;
  }
  __syncwarp();
}

// start issuing for a phase (its counter, its unit range); called before the grid barrier that precedes the phase
template <int CH>
__device__ __forceinline__ void dyn_begin(const Sell& S, WarpCtx<CH>& W, DynWarp* D, DynState& st, uint32_t* ctr,
                                          uint32_t ub0, uint32_t ub1) {
  st.i_col = st.i_end = 0;
  st.i_done = (ub0 >= ub1);
  dyn_issue(S, W, D, st, ctr, ub0, ub1);
}

template <int PHASE, int CH>
__device__ __forceinline__ void run_phase_dyn(const EmArgs& A, WarpCtx<CH>& W, DynWarp* D, DynState& st, uint32_t* ctr,
                                              uint32_t ub0, uint32_t ub1, uint32_t bid, uint32_t nblk, double logNorm,
                                              double bias, P2Acc& pa) {
  const Sell& S = (PHASE == 1) ? A.cm : A.tm;
  const double* gsrc = (PHASE == 1) ? A.theta : A.scale;   // rewritten by other blocks: coherent loads only
  const bool em_nan_guard = (PHASE == 1) && !A.vbem;
  const uint32_t lane = threadIdx.x & 31u;
  // the unit being consumed
  uint32_t s = 0, u_s0 = 0, u_s1 = 0, u_rem = 0, slice_end = 0, sp = 0;
  RowOps ops;
  ops.x0 = ops.x1 = ops.x2 = ops.x3 = 0.0;
  ops.len = LEN_LONG;
  double acc = 0.0;
  auto next_slice = [&]() {
    row_finish<PHASE>(A, s * 32u + lane, ops, acc, logNorm, bias, pa);
    ++s;
    acc = 0.0;
    if (s < u_s1) {
      slice_end = __shfl_sync(0xffffffffu, sp, (int)(s - u_s0));
      ops = load_ops<PHASE>(A, S, s * 32u + lane);
    }
  };
  for (;;) {
    if (st.kc == st.ki) {
      dyn_issue(S, W, D, st, ctr, ub0, ub1);
      if (st.kc == st.ki) break;           // nothing in flight, nothing left to grab
    }
    const int slot = st.kc % RING;
    mbar_wait(&W.bars[slot], (W.phase_bits >> slot) & 1u);
    W.phase_bits ^= (1u << slot);
    const uint32_t c0 = D->cq[slot][0], cstop = D->cq[slot][1];
    if (u_rem == 0) {                      // this chunk opens the next unit
      while (s < u_s1) next_slice();       // (trailing zero-width slices of the previous unit)
      const uint32_t q = st.un_out % UQ;
      const uint4 d = D->uq[q];
      sp = D->sp[q][lane];
      ++st.un_out;
      u_s0 = d.x; u_s1 = d.y; u_rem = d.w - d.z;
      s = u_s0;
      slice_end = __shfl_sync(0xffffffffu, sp, 0);
      ops = load_ops<PHASE>(A, S, s * 32u + lane);
      acc = 0.0;
    }
    const uint32_t* sidx = W.ring->idx[slot] + lane;
    const double* sw = W.ring->w[slot] + lane;
    uint32_t col = c0;
    while (col < cstop) {
      while (col == slice_end && s + 1 < u_s1) next_slice();   // (possibly zero-width slices)
      const uint32_t n = min(slice_end, cstop) - col;
      const uint32_t l0 = col - c0;
      uint32_t j = 0;
      for (; j + 4 <= n; j += 4) {
        const uint32_t o = (l0 + j) * 32u;
        const double g0 = gsrc[sidx[o]], g1 = gsrc[sidx[o + 32]];
        const double g2 = gsrc[sidx[o + 64]], g3 = gsrc[sidx[o + 96]];
        double v0 = g0 * sw[o], v1 = g1 * sw[o + 32], v2 = g2 * sw[o + 64], v3 = g3 * sw[o + 96];
        if (em_nan_guard) {
          if (isnan(v0)) v0 = 0.0;
          if (isnan(v1)) v1 = 0.0;
          if (isnan(v2)) v2 = 0.0;
          if (isnan(v3)) v3 = 0.0;
        }
        acc += v0; acc += v1; acc += v2; acc += v3;
      }
      for (; j < n; ++j) {
        const uint32_t o = (l0 + j) * 32u;
        double v = gsrc[sidx[o]] * sw[o];
        if (em_nan_guard && isnan(v)) v = 0.0;
        acc += v;
      }
      col += n;
    }
    u_rem -= cstop - c0;
    __syncwarp();
    ++st.kc;
    dyn_issue(S, W, D, st, ctr, ub0, ub1);
  }
  while (s < u_s1) next_slice();
  if (W.dbg && lane == 0) *W.dbg = gtime_ns();
  // long rows exactly as in MODE 0 (static distribution)
  for (uint32_t li = bid; li < S.n_block; li += nblk) {
    const uint32_t r = __ldg(&S.long_rows[3 * li]);
    const uint32_t b = __ldg(&S.long_rows[3 * li + 1]);
    const uint32_t e = __ldg(&S.long_rows[3 * li + 2]);
    double a = 0.0;
    for (uint32_t k = b + threadIdx.x; k < e; k += EM_THREADS) {
      double v = gsrc[__ldg(&S.csr_idx[k])] * __ldg(&S.csr_w[k]);
      if (em_nan_guard && isnan(v)) v = 0.0;
      a += v;
    }
    a = block_reduce<false>(a, W.scratch);
    if (threadIdx.x == 0) {
      RowOps o = load_ops<PHASE>(A, S, r);
      o.len = 0;
      row_finish<PHASE>(A, r, o, a, logNorm, bias, pa);
    }
    __syncthreads();
  }
  {
    const uint32_t gw = bid * EM_WARPS + (threadIdx.x >> 5);
    const uint32_t nw = nblk * EM_WARPS;
    uint32_t cnt = 0, myrow = 0xffffffffu;
    double myacc = 0.0;
    auto flush = [&]() {
      if (myrow != 0xffffffffu) {
        RowOps o = load_ops<PHASE>(A, S, myrow);
        o.len = 0;
        row_finish<PHASE>(A, myrow, o, myacc, logNorm, bias, pa);
      }
      myrow = 0xffffffffu;
      cnt = 0;
    };
    for (uint32_t li = S.n_block + gw; li < S.n_long; li += nw) {
      const uint32_t r = __ldg(&S.long_rows[3 * li]);
      const uint32_t b = __ldg(&S.long_rows[3 * li + 1]);
      const uint32_t e = __ldg(&S.long_rows[3 * li + 2]);
      double a0 = 0.0, a1 = 0.0;
      uint32_t k = b + lane;
      for (; k + 32 < e; k += 64) {
        const uint32_t i0 = __ldg(&S.csr_idx[k]), i1 = __ldg(&S.csr_idx[k + 32]);
        double v0 = gsrc[i0] * __ldg(&S.csr_w[k]);
        double v1 = gsrc[i1] * __ldg(&S.csr_w[k + 32]);
        if (em_nan_guard) {
          if (isnan(v0)) v0 = 0.0;
          if (isnan(v1)) v1 = 0.0;
        }
        a0 += v0;
        a1 += v1;
      }
      if (k < e) {
        double v = gsrc[__ldg(&S.csr_idx[k])] * __ldg(&S.csr_w[k]);
        if (em_nan_guard && isnan(v)) v = 0.0;
        a0 += v;
      }
      const double a = warp_sum(a0 + a1);
      if (lane == cnt) { myacc = a; myrow = r; }
      if (++cnt == 32) flush();
    }
    flush();
  }
}

template <int CH, int MINB>
__global__ void __launch_bounds__(EM_THREADS, MINB) k_em_persistent_dyn(const __grid_constant__ EmArgs A) {
  extern __shared__ __align__(128) unsigned char smem[];
  WarpCtx<CH> W;
  warp_setup(W, smem);
  double* scratch = W.scratch;
  DynWarp* D = reinterpret_cast<DynWarp*>(smem + em_smem<CH>()) + (threadIdx.x >> 5);
  uint32_t* s_ctr = reinterpret_cast<uint32_t*>(smem + em_smem<CH>() + sizeof(DynWarp) * EM_WARPS);   // [0] P1, [1] P2
  if (threadIdx.x == 0) { s_ctr[0] = 0u; s_ctr[1] = 0u; }
  __syncthreads();
  cg::grid_group grid = cg::this_grid();
  const uint32_t bid = blockIdx.x, nblk = gridDim.x;
  const uint32_t gwarp = bid * (EM_THREADS / 32) + (threadIdx.x >> 5);
  const uint32_t u1a = __ldg(&A.cm.blk_unit_ptr[bid]), u1b = __ldg(&A.cm.blk_unit_ptr[bid + 1]);
  const uint32_t u2a = __ldg(&A.tm.blk_unit_ptr[bid]), u2b = __ldg(&A.tm.blk_unit_ptr[bid + 1]);
  DynState st;
  uint32_t it = 0;
  bool converged = false;
  double logNorm = A.vbem ? digamma_pos(A.sum0) : 0.0;
  dyn_begin(A.cm, W, D, st, &s_ctr[0], u1a, u1b);
  while (it < A.min_iter || (it < A.max_iter && !converged)) {
    const uint32_t par = it & 1u;
    if (bid == 0 && threadIdx.x == 0) A.maxrel[par] = 0ull;
    if (A.vbem && it > 0) lag_lognorm_warp0(A, par, nblk, scratch);  // consumed after the next barrier
    P2Acc pa{0.0, 0.0};
    SB_DBG(0)
    run_phase_dyn<1, CH>(A, W, D, st, &s_ctr[0], u1a, u1b, bid, nblk, 0.0, 0.0, pa);
    SB_DBG(1)
    dyn_begin(A.tm, W, D, st, &s_ctr[1], u2a, u2b);   // first units of P2: their stream lands during the barrier
    __syncthreads();                                   // every warp of the block is done taking P1 units ...
    if (threadIdx.x == 0) s_ctr[0] = 0u;               // ... so the P1 counter can be re-armed for the next iteration
    grid.sync();
    SB_DBG(2)
    if (A.vbem && it > 0) logNorm = scratch[33];
    const double bias = (it == 0) ? A.first_bias : 0.0;
    SB_DBG(3)
    W.dbg = (A.dbg && it == A.dbg_it) ? &A.dbg[(size_t)gwarp * 8 + 7] : nullptr;
    run_phase_dyn<2, CH>(A, W, D, st, &s_ctr[1], u2a, u2b, bid, nblk, logNorm, bias, pa);
    W.dbg = nullptr;
    SB_DBG(4)
    dyn_begin(A.cm, W, D, st, &s_ctr[0], u1a, u1b);   // next iteration's P1 (harmless if the loop ends)
    p2_finish(A, scratch, pa, par);
    SB_DBG(5)
    __syncthreads();
    if (threadIdx.x == 0) s_ctr[1] = 0u;
    grid.sync();
    SB_DBG(6)
    const double mr = __longlong_as_double((long long)__ldcg(&A.maxrel[par]));
    converged = !(mr > A.tol);
    ++it;
  }
  if (bid == 0 && threadIdx.x == 0) {
    A.out[0] = it;
    A.out[1] = converged ? 1u : 0u;
    A.out[2] = (it - 1) & 1u;
  }
  // drain the speculative prefetch before the block (and its shared memory) retires
  while (st.kc < st.ki) {
    const int slot = st.kc % RING;
    mbar_wait(&W.bars[slot], (W.phase_bits >> slot) & 1u);
    W.phase_bits ^= (1u << slot);
    ++st.kc;
  }
}

// ---- multi-GPU persistent kernel: classes stay sharded per rank; alpha' is all-reduced INSIDE the kernel over
// peer memory (NVLink P2P), once per iteration, as a reduce-scatter + all-gather:
//   P1, P2-partial (this rank's share of alpha' per transcript id, into its exchange block)
//   barrier over all GPUs -> rank r sums slice r of every rank's partial in fixed rank order (remote loads) and
//   stores the sums into every rank's `red` (remote stores) -> barrier over all GPUs -> every rank runs the same
//   update on identical data (identical alpha, theta, convergence decision).
// GPU-to-GPU barrier: system-scope fence, grid barrier, block 0 pushes the epoch into each peer's flag slot
// (st.release.sys) and polls its own slots (ld.acquire.sys), grid barrier.
__device__ __forceinline__ void st_release_sys(unsigned long long* p, unsigned long long v) {
  asm volatile("st.release.sys.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}
__device__ __forceinline__ unsigned long long ld_acquire_sys(const unsigned long long* p) {
  unsigned long long v;
  asm volatile("ld.acquire.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ double ld_relaxed_sys_f64(const double* p) {
  double v;
  asm volatile("ld.relaxed.sys.global.f64 %0, [%1];" : "=d"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ void st_relaxed_sys_f64(double* p, double v) {
  asm volatile("st.relaxed.sys.global.f64 [%0], %1;" ::"l"(p), "d"(v) : "memory");
}
__device__ __forceinline__ void xgpu_barrier(cg::grid_group& grid, const EmArgs& A, unsigned long long epoch) {
  __threadfence_system();
  grid.sync();
  if (blockIdx.x == 0 && threadIdx.x < A.nranks) {
    const uint32_t q = threadIdx.x;
    st_release_sys(reinterpret_cast<unsigned long long*>(A.peers[q]) + A.rank, epoch);
    const unsigned long long* mine = reinterpret_cast<const unsigned long long*>(A.peers[A.rank]) + q;
    const unsigned long long t0 = gtime_ns();
    while (ld_acquire_sys(mine) < epoch) {
      if (gtime_ns() - t0 > 20000000000ull) { *A.xfail = 1u; break; }   // a peer is missing: give up, report
    }
  }
  grid.sync();
}

template <int CH, int MINB, int MODE>
__global__ void __launch_bounds__(EM_THREADS, MINB) k_em_persistent_mgpu(const __grid_constant__ EmArgs A) {
  extern __shared__ __align__(128) unsigned char smem[];
  WarpCtx<CH> W;
  warp_setup(W, smem);
  double* scratch = W.scratch;
  cg::grid_group grid = cg::this_grid();
  const uint32_t bid = blockIdx.x, nblk = gridDim.x;
  const uint32_t gwarp = bid * (EM_THREADS / 32) + (threadIdx.x >> 5);
  const uint32_t gtid = bid * EM_THREADS + threadIdx.x, gthreads = nblk * EM_THREADS;
  const WarpRange R1 = load_range(A.cm, gwarp);
  const WarpRange R2 = load_range(A.tm, gwarp);
  const uint32_t M = A.M, G = A.nranks;
  const uint32_t S = (M + G - 1) / G, lo = A.rank * S, hi = (lo + S < M) ? lo + S : M;
  double* my_red = A.peers[A.rank] + 64 + M;
  unsigned long long epoch = A.epoch0;
  uint32_t it = 0;
  bool converged = false;
  ring_prefetch(A.cm, W, R1);
  while (it < A.min_iter || (it < A.max_iter && !converged)) {
    const uint32_t par = it & 1u;
    if (bid == 0 && threadIdx.x == 0) A.maxrel[par] = 0ull;
    P2Acc pa{0.0, 0.0};
    run_phase_m<1, CH, MODE>(A, W, R1, bid, nblk, 0.0, 0.0, pa);
    ring_prefetch(A.tm, W, R2);
    grid.sync();
    run_phase_m<3, CH, MODE>(A, W, R2, bid, nblk, 0.0, 0.0, pa);      // A.part_out = own exchange block
    ring_prefetch(A.cm, W, R1);
    xgpu_barrier(grid, A, ++epoch);                            // every rank's partial is complete and visible
    for (uint32_t t = lo + gtid; t < hi; t += gthreads) {
      double v = 0.0;
      for (uint32_t q = 0; q < G; ++q) v += ld_relaxed_sys_f64(A.peers[q] + 64 + t);
      for (uint32_t q = 0; q < G; ++q) st_relaxed_sys_f64(A.peers[q] + 64 + M + t, v);
    }
    xgpu_barrier(grid, A, ++epoch);                            // every slice has been delivered everywhere
    // ---- update (k_em_update): identical on every rank
    double logNorm = 0.0;
    if (A.vbem) {
      if (it == 0) logNorm = digamma_pos(A.sum0);
      else logNorm = digamma_pos(sum_partials(A.sum_partial + (size_t)(par ^ 1u) * nblk, nblk, 0.0, scratch));
    }
    const double bias = (it == 0) ? A.first_bias : 0.0;
    double sum = 0.0, mx = 0.0;
    for (uint32_t t = gtid; t < M; t += gthreads) {
      const double na = __ldcg(&my_red[t]) + bias;
      const double old = A.alpha[t];
      if (na > ALPHA_CHECK_CUTOFF) mx = fmax(mx, fabs(old - na) / na);
      A.alpha[t] = na;
      const double ap = na + A.prior[t];
      sum += ap;
      A.theta[t] = A.vbem ? ((ap > DIGAMMA_MIN) ? exp(digamma_pos(ap) - logNorm) : 0.0) : na;
    }
    pa.sum = sum; pa.maxrel = mx;
    p2_finish(A, scratch, pa, par);
    grid.sync();
    const double mr = __longlong_as_double((long long)__ldcg(&A.maxrel[par]));
    converged = !(mr > A.tol);
    ++it;
  }
  if (bid == 0 && threadIdx.x == 0) {
    A.out[0] = it;
    A.out[1] = converged ? 1u : 0u;
    A.out[2] = (it - 1) & 1u;
  }
  {
    const uint32_t nchunks = (R1.cend - R1.cbeg + CH - 1) / CH;
#pragma unroll
    for (int k = 0; k < RING; ++k)
      if ((uint32_t)k < nchunks) mbar_wait(&W.bars[k], (W.phase_bits >> k) & 1u);
  }
}

// ---- one launch per phase (baseline variant; also the multi-GPU building blocks)
template <int CH, int MINB, int MODE>
__global__ void __launch_bounds__(EM_THREADS, MINB) k_em_p1(const __grid_constant__ EmArgs A) {
  extern __shared__ __align__(128) unsigned char smem[];
  WarpCtx<CH> W;
  warp_setup(W, smem);
  P2Acc pa{0.0, 0.0};
  const WarpRange R = load_range(A.cm, blockIdx.x * (EM_THREADS / 32) + (threadIdx.x >> 5));
  ring_prefetch(A.cm, W, R);
  run_phase_m<1, CH, MODE>(A, W, R, blockIdx.x, gridDim.x, 0.0, 0.0, pa);
}
template <int CH, int MINB, int MODE>
__global__ void __launch_bounds__(EM_THREADS, MINB) k_em_p2(const __grid_constant__ EmArgs A, uint32_t it) {
  extern __shared__ __align__(128) unsigned char smem[];
  WarpCtx<CH> W;
  warp_setup(W, smem);
  double* scratch = W.scratch;
  const uint32_t par = it & 1u;
  double logNorm = 0.0;
  if (A.vbem) {
    if (it == 0) {
      logNorm = digamma_pos(A.sum0);
    } else {
      lag_lognorm_warp0(A, par, gridDim.x, scratch);
      __syncthreads();
      logNorm = scratch[33];
    }
  }
  const double bias = (it == 0) ? A.first_bias : 0.0;
  P2Acc pa{0.0, 0.0};
  const WarpRange R = load_range(A.tm, blockIdx.x * (EM_THREADS / 32) + (threadIdx.x >> 5));
  ring_prefetch(A.tm, W, R);
  run_phase_m<2, CH, MODE>(A, W, R, blockIdx.x, gridDim.x, logNorm, bias, pa);
  p2_finish(A, scratch, pa, par);
}
template <int CH, int MINB, int MODE>
__global__ void __launch_bounds__(EM_THREADS, MINB) k_em_p2_partial(const __grid_constant__ EmArgs A) {
  extern __shared__ __align__(128) unsigned char smem[];
  WarpCtx<CH> W;
  warp_setup(W, smem);
  P2Acc pa{0.0, 0.0};
  const WarpRange R = load_range(A.tm, blockIdx.x * (EM_THREADS / 32) + (threadIdx.x >> 5));
  ring_prefetch(A.tm, W, R);
  run_phase_m<3, CH, MODE>(A, W, R, blockIdx.x, gridDim.x, 0.0, 0.0, pa);
}

}  // namespace sb
