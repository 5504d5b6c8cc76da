// em.cu -- Stage B: EM / VBEM over equivalence classes on B200 (sm_100a).
//
// Replaces CollapsedEMOptimizer::optimize (src/inference/CollapsedEMOptimizer.cpp:732-1035)
// and its update kernels EMUpdate_ (:178-234) / VBEMUpdate_ (:241-328).
//
// Design (DESIGN.md "Stage B"): the class<->transcript map is held twice in HBM,
// class-major and transcript-major, so that one iteration is two segmented
// reductions with NO atomics and a fixed summation order:
//   P1 (class-major):  denom_c = sum_i theta[t_i] * w_ci ;  scale_c = count_c / denom_c
//   P2 (txp-major):    alpha'_t = base_t + theta_t * sum_c w_ct * scale_c
//                      + convergence test + theta'_t (VBEM: exp(digamma(alpha'+prior) - logNorm))
// Both passes stream fixed-size windows ("tiles") of the entry arrays into shared
// memory with 1-D bulk (TMA) copies on an mbarrier ring and gather theta / scale
// from L2.  A persistent cooperative kernel runs the whole iteration loop with two
// grid barriers per iteration; the multi-kernel variant launches P1 / P2 separately
// (used for the multi-GPU path where an all-reduce sits between P2 and the update).
#include <cooperative_groups.h>
#include <cub/cub.cuh>
#include <float.h>
#include <math.h>
#include <stdarg.h>
#include <string.h>

#include <algorithm>
#include <vector>

#include "common.cuh"
#include "em_internal.h"

namespace cg = cooperative_groups;

namespace sb {

// ---------------------------------------------------------------------------
// constants
// ---------------------------------------------------------------------------
constexpr int TILE = 2048;            // entries per tile window
constexpr int LMAX = 256;             // rows longer than this take the block path
constexpr int CAP = TILE + LMAX + 8;  // shared-memory entries per stage
constexpr int STAGES = 2;
constexpr int THREADS = 256;
constexpr double DIGAMMA_MIN = 1e-10;     // CollapsedEMOptimizer.cpp:43
constexpr double MIN_EQ_W = DBL_MIN;      // :40
constexpr double ALPHA_CHECK_CUTOFF = 1e-2;  // :884

struct SmemLayout {
  // per stage: idx[CAP] u32 then w[CAP] f64 ; then barriers + scratch
  static constexpr size_t idx_bytes = size_t(CAP) * 4;
  static constexpr size_t w_bytes = size_t(CAP) * 8;
  static constexpr size_t stage_bytes = idx_bytes + w_bytes;
  static constexpr size_t bars_off = stage_bytes * STAGES;
  static constexpr size_t scratch_off = bars_off + 64;
  static constexpr size_t total = scratch_off + 64 * 8;
};

// ---------------------------------------------------------------------------
// prepare kernels
// ---------------------------------------------------------------------------

// CollapsedEMOptimizer.cpp:778-823 : per-transcript initialisation.
__global__ void k_txp_init(uint32_t M, const double* __restrict__ projected,
                           const double* __restrict__ eff_in,
                           const uint64_t* __restrict__ unique, sb_em_params p,
                           double totalWeight, double* __restrict__ effLens,
                           double* __restrict__ prior, double* __restrict__ alpha0) {
  uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= M) return;
  double el = p.no_length_correction ? 100.0 : eff_in[i];
  effLens[i] = el;
  prior[i] = p.per_txp_prior ? p.vb_prior : p.vb_prior * el;
  double uniqueCount = (double)unique[i] + 0.5;
  double wi = p.init_uniform ? 100.0 : (uniqueCount * 1e-3 * el);
  double a;
  if (p.init_uniform) {
    a = wi;
  } else {
    double uniformPrior = totalWeight / (double)M;
    double fracObserved = fmin(0.999, totalWeight / p.num_required_frags);
    double uniAbund = p.alt_init ? wi : uniformPrior;
    a = __dadd_rn(__dmul_rn(projected[i], fracObserved), __dmul_rn(uniAbund, (1.0 - fracObserved)));
  }
  alpha0[i] = a;
}

// :830-873 combined weights, :330-394 degenerate marking, singleton folding.
// One thread per class (one-time work).
__global__ void k_class_combine(uint64_t C, const uint64_t* __restrict__ off,
                                const uint32_t* __restrict__ tids,
                                const double* __restrict__ aux,
                                const uint64_t* __restrict__ counts,
                                const double* __restrict__ effLens,
                                const double* __restrict__ alpha0, sb_em_params p,
                                double* __restrict__ cw, uint64_t* __restrict__ packed,
                                double* __restrict__ single, uint8_t* __restrict__ valid,
                                unsigned long long* __restrict__ n_degenerate) {
  uint64_t c = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= C) return;
  const uint64_t b = off[c], e = off[c + 1];
  const double count = (double)counts[c];
  double wsum = 0.0;
  for (uint64_t j = b; j < e; ++j) {
    double el = effLens[tids[j]];
    if (el <= 1.0) el = 1.0;
    double w = p.no_rich_eq ? 1.0 : aux[j];
    double probStartPos = 1.0 / el;
    double wt = p.eq_class_mode ? w : __dmul_rn(__dmul_rn(count, w), probStartPos);
    cw[j] = wt;
    wsum = __dadd_rn(wsum, wt);
  }
  const double wnorm = 1.0 / wsum;
  double denom = 0.0;
  for (uint64_t j = b; j < e; ++j) {
    double v = __dmul_rn(cw[j], wnorm);
    cw[j] = v;
    double d = __dmul_rn(alpha0[tids[j]], v);
    if (!isnan(d)) denom = __dadd_rn(denom, d);
  }
  const bool ok = !(denom <= MIN_EQ_W);
  valid[c] = ok ? 1 : 0;
  uint64_t len = e - b;
  uint64_t pk = 0;
  if (!ok) {
    atomicAdd(n_degenerate, 1ull);
  } else if (len == 1) {
    atomicAdd(&single[tids[b]], count);  // integer-valued: order independent
  } else if (len > 1) {
    pk = (1ull << 32) | len;             // (class count, entry count)
  }
  packed[c] = pk;
}

// Compact valid multi-transcript classes; histogram of transcript occurrences.
__global__ void k_compact(uint64_t C, const uint64_t* __restrict__ off,
                          const uint32_t* __restrict__ tids, const double* __restrict__ cw,
                          const uint64_t* __restrict__ counts,
                          const uint64_t* __restrict__ packed,
                          const uint64_t* __restrict__ packed_scan,
                          uint32_t* __restrict__ m_off, uint32_t* __restrict__ m_idx,
                          double* __restrict__ m_w, double* __restrict__ m_cnt,
                          uint32_t* __restrict__ ent_cls, uint32_t* __restrict__ tcnt) {
  uint64_t c = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= C) return;
  if (packed[c] == 0) return;
  const uint64_t s = packed_scan[c];
  const uint32_t cid = (uint32_t)(s >> 32);
  uint32_t o = (uint32_t)(s & 0xffffffffu);
  m_off[cid] = o;
  m_cnt[cid] = (double)counts[c];
  for (uint64_t j = off[c]; j < off[c + 1]; ++j, ++o) {
    uint32_t t = tids[j];
    m_idx[o] = t;
    m_w[o] = cw[j];
    ent_cls[o] = cid;
    atomicAdd(&tcnt[t], 1u);
  }
}

__global__ void k_iota(uint32_t n, uint32_t* __restrict__ v) {
  uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) v[i] = i;
}

// per transcript: packed (active flag, occurrence count) for the row scan
__global__ void k_row_pack(uint32_t M, const uint32_t* __restrict__ tcnt,
                           uint64_t* __restrict__ packed) {
  uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < M) packed[i] = tcnt[i] ? ((1ull << 32) | tcnt[i]) : 0ull;
}
__global__ void k_row_fill(uint32_t M, const uint32_t* __restrict__ tcnt,
                           const uint64_t* __restrict__ scan, uint32_t* __restrict__ t_off,
                           uint32_t* __restrict__ row_tid, uint32_t* __restrict__ tid_row) {
  uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= M) return;
  if (tcnt[i]) {
    uint32_t row = (uint32_t)(scan[i] >> 32);
    t_off[row] = (uint32_t)(scan[i] & 0xffffffffu);
    row_tid[row] = i;
    tid_row[i] = row;
  } else {
    tid_row[i] = 0xffffffffu;
  }
}
// transcript-major entries from the stable sort permutation
__global__ void k_gather_csc(uint32_t nnz, const uint32_t* __restrict__ perm,
                             const uint32_t* __restrict__ ent_cls,
                             const double* __restrict__ m_w, uint32_t* __restrict__ t_idx,
                             double* __restrict__ t_w) {
  uint32_t k = blockIdx.x * blockDim.x + threadIdx.x;
  if (k >= nnz) return;
  uint32_t j = perm[k];
  t_idx[k] = ent_cls[j];
  t_w[k] = m_w[j];
}

// tile k owns the rows whose first entry lies in [k*TILE, (k+1)*TILE).
// desc = {row0, row1, ent0, ent1}: rows [row0,row1) with entries [ent0,ent1);
// rows longer than LMAX are skipped by the tile pass (block path), and if the
// last row of the tile is long the span stops at its first entry.
__global__ void k_tiles(uint32_t n_rows, const uint32_t* __restrict__ off, uint32_t nnz,
                        uint32_t n_tiles, uint4* __restrict__ desc) {
  uint32_t k = blockIdx.x * blockDim.x + threadIdx.x;
  if (k >= n_tiles) return;
  auto lower = [&](uint32_t target) {
    uint32_t lo = 0, hi = n_rows;  // first row with off[row] >= target
    while (lo < hi) {
      uint32_t mid = (lo + hi) >> 1;
      if (off[mid] >= target) hi = mid; else lo = mid + 1;
    }
    return lo;
  };
  uint32_t r0 = lower(k * (uint32_t)TILE);
  uint32_t r1 = (k + 1 == n_tiles) ? n_rows : lower((k + 1) * (uint32_t)TILE);
  uint32_t e0 = (r0 < n_rows) ? off[r0] : nnz;
  uint32_t e1 = off[r1];
  if (r1 > r0) {
    uint32_t lastlen = off[r1] - off[r1 - 1];
    if (lastlen > (uint32_t)LMAX) e1 = off[r1 - 1];
  } else {
    e1 = e0;
  }
  desc[k] = make_uint4(r0, r1, e0, e1);
}
__global__ void k_long_rows(uint32_t n_rows, const uint32_t* __restrict__ off,
                            uint32_t* __restrict__ list, uint32_t* __restrict__ n_long,
                            uint32_t cap) {
  uint32_t r = blockIdx.x * blockDim.x + threadIdx.x;
  if (r >= n_rows) return;
  if (off[r + 1] - off[r] > (uint32_t)LMAX) {
    uint32_t pos = atomicAdd(n_long, 1u);
    if (pos < cap) list[pos] = r;
  }
}

// deterministic single-block reduction: out[0] = sum_i f(i)
// mode 0: alpha[i]+prior[i] over all i ; mode 1: base[i]+prior[i] over inactive i
__global__ void k_sum1(uint32_t M, const double* __restrict__ a, const double* __restrict__ b,
                       const uint32_t* __restrict__ tid_row, int mode, double* __restrict__ out) {
  __shared__ double scratch[32];
  double acc = 0.0;
  for (uint32_t i = threadIdx.x; i < M; i += blockDim.x) {
    if (mode == 1 && tid_row[i] != 0xffffffffu) continue;
    acc += a[i] + b[i];
  }
  acc = block_reduce<false>(acc, scratch);
  if (threadIdx.x == 0) out[0] = acc;
}

// theta for iteration 0 (exact logNorm) + bookkeeping for inactive transcripts.
__global__ void k_theta0(uint32_t M, int vbem, const double* __restrict__ alpha0,
                         const double* __restrict__ prior, const double* __restrict__ sum0,
                         double* __restrict__ alpha, double* __restrict__ theta) {
  uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= M) return;
  double a = alpha0[i];
  alpha[i] = a;
  if (vbem) {
    double logNorm = digamma_pos(sum0[0]);
    double ap = a + prior[i];
    theta[i] = (ap > DIGAMMA_MIN) ? exp(digamma_pos(ap) - logNorm) : 0.0;
  } else {
    theta[i] = a;
  }
}

// inactive transcripts (in no valid multi-transcript class): alpha after >=1
// iteration is base (+1.0 after exactly one EM iteration, see em_internal.h).
__global__ void k_finalize_inactive(uint32_t M, const uint32_t* __restrict__ tid_row,
                                    const double* __restrict__ base, double bias,
                                    double* __restrict__ alpha) {
  uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= M) return;
  if (tid_row[i] == 0xffffffffu) alpha[i] = base[i] + bias;
}

// ---------------------------------------------------------------------------
// iteration kernels
// ---------------------------------------------------------------------------
struct Block {
  uint32_t* s_idx[STAGES];
  double* s_w[STAGES];
  uint64_t* bars;
  double* scratch;
  uint32_t uses[STAGES];
};

__device__ __forceinline__ void block_setup(Block& B, unsigned char* smem) {
#pragma unroll
  for (int s = 0; s < STAGES; ++s) {
    B.s_w[s] = reinterpret_cast<double*>(smem + s * SmemLayout::stage_bytes);
    B.s_idx[s] = reinterpret_cast<uint32_t*>(smem + s * SmemLayout::stage_bytes + SmemLayout::w_bytes);
    B.uses[s] = 0;
  }
  B.bars = reinterpret_cast<uint64_t*>(smem + SmemLayout::bars_off);
  B.scratch = reinterpret_cast<double*>(smem + SmemLayout::scratch_off);
  if (threadIdx.x == 0) {
    for (int s = 0; s < STAGES; ++s) mbar_init(&B.bars[s], 1);
    mbar_fence_init();
  }
  __syncthreads();
}

// issue the bulk loads of one tile into stage s (one elected thread)
__device__ __forceinline__ void tile_issue(Block& B, int s, const uint4 d,
                                           const uint32_t* __restrict__ idx,
                                           const double* __restrict__ w) {
  const uint32_t e0a = d.z & ~3u;
  const uint32_t e1a = (d.w + 3u) & ~3u;
  const uint32_t n = e1a - e0a;
  if (n == 0) return;
  fence_proxy_async();  // generic-proxy writes to this stage happen-before the async writes
  mbar_arrive_expect_tx(&B.bars[s], n * 12u);
  bulk_g2s(B.s_w[s], w + e0a, n * 8u, &B.bars[s]);
  bulk_g2s(B.s_idx[s], idx + e0a, n * 4u, &B.bars[s]);
}

// PHASE 1: rows = classes, gather theta, out scale[row] = cnt[row]/denom
// PHASE 2: rows = active transcripts, gather scale, epilogue per transcript
struct P2Acc {
  double sum;     // sum of (alpha' + prior) over my rows
  double maxrel;  // max rel diff over my rows
};

struct EmArgs {
  // class-major
  const uint32_t* c_off; const uint32_t* c_idx; const double* c_w; const uint4* c_tiles;
  const uint32_t* c_long; const double* c_cnt; double* scale;
  uint32_t c_ntiles, c_nlong, c_nrows;
  // transcript-major
  const uint32_t* t_off; const uint32_t* t_idx; const double* t_w; const uint4* t_tiles;
  const uint32_t* t_long; const uint32_t* row_tid;
  uint32_t t_ntiles, t_nlong, t_nrows;
  // per transcript state
  double* alpha; double* theta; const double* prior; const double* base;
  double* part_out;   // multi-GPU: partial sums per transcript row (no epilogue)
  // reductions
  double* sum_partial;            // [2][grid]
  unsigned long long* maxrel;     // [2] (bit pattern of a non-negative double)
  double inactive_sum; double sum0;
  double tol;
  uint32_t min_iter, max_iter;
  int vbem;
  uint32_t* out;                  // [0]=iters [1]=converged ; maxrel of last iter in maxrel
};

template <int PHASE>
__device__ __forceinline__ void row_epilogue(const EmArgs& A, uint32_t row, double acc,
                                             double logNorm, double bias, P2Acc& pa) {
  if (PHASE == 1) {
    A.scale[row] = (acc <= MIN_EQ_W) ? 0.0 : A.c_cnt[row] / acc;
  } else if (PHASE == 3) {
    // multi-GPU: this rank's share of alpha'_t; the update runs after the all-reduce
    const uint32_t t = A.row_tid[row];
    const double th = A.theta[t];
    double na = A.base[t];
    if (th > 0.0) na += th * acc;
    A.part_out[t] = na;
  } else {
    const uint32_t t = A.row_tid[row];
    const double th = A.theta[t];
    const double pr = A.prior[t];
    double na = A.base[t] + bias;
    if (th > 0.0) na += th * acc;
    const double old = A.alpha[t];
    if (na > ALPHA_CHECK_CUTOFF) {
      double rel = fabs(old - na) / na;
      pa.maxrel = fmax(pa.maxrel, rel);
    }
    A.alpha[t] = na;
    const double ap = na + pr;
    pa.sum += ap;
    if (A.vbem) {
      A.theta[t] = (ap > DIGAMMA_MIN) ? exp(digamma_pos(ap) - logNorm) : 0.0;
    } else {
      A.theta[t] = na;
    }
  }
}

template <int PHASE>
__device__ __forceinline__ void run_phase(const EmArgs& A, Block& B, uint32_t bid, uint32_t nblk,
                                          double logNorm, double bias, P2Acc& pa) {
  const uint32_t* __restrict__ off = (PHASE == 1) ? A.c_off : A.t_off;
  const uint32_t* __restrict__ idx = (PHASE == 1) ? A.c_idx : A.t_idx;
  const double* __restrict__ w = (PHASE == 1) ? A.c_w : A.t_w;
  const uint4* __restrict__ tiles = (PHASE == 1) ? A.c_tiles : A.t_tiles;
  // theta / scale are rewritten by other blocks inside the persistent kernel: plain
  // (coherent) loads only -- never ld.global.nc / __restrict__ for these two.
  const double* gsrc = (PHASE == 1) ? A.theta : A.scale;
  const uint32_t ntiles = (PHASE == 1) ? A.c_ntiles : A.t_ntiles;
  const uint32_t nlong = (PHASE == 1) ? A.c_nlong : A.t_nlong;
  const uint32_t* __restrict__ longs = (PHASE == 1) ? A.c_long : A.t_long;
  const bool em_nan_guard = (PHASE == 1) && !A.vbem;
  (void)A.c_nrows;

  const uint32_t n_my = (bid < ntiles) ? (ntiles - bid + nblk - 1) / nblk : 0;
  // prologue: prefetch STAGES-1 tiles
  if (threadIdx.x == 0) {
    for (uint32_t i = 0; i < (uint32_t)(STAGES - 1) && i < n_my; ++i)
      tile_issue(B, i % STAGES, tiles[bid + i * nblk], idx, w);
  }
  for (uint32_t i = 0; i < n_my; ++i) {
    const int s = i % STAGES;
    const uint4 d = tiles[bid + i * nblk];
    if (threadIdx.x == 0 && i + STAGES - 1 < n_my)
      tile_issue(B, (i + STAGES - 1) % STAGES, tiles[bid + (i + STAGES - 1) * nblk], idx, w);
    const uint32_t e0a = d.z & ~3u;
    const uint32_t n = ((d.w + 3u) & ~3u) - e0a;
    if (n) {
      mbar_wait(&B.bars[s], B.uses[s] & 1u);
      ++B.uses[s];
      uint32_t* sidx = B.s_idx[s];
      double* sw = B.s_w[s];
      // (1) gather * weight, in place
      for (uint32_t k = threadIdx.x; k < n; k += THREADS) {
        const double g = gsrc[sidx[k]];
        double v = g * sw[k];
        if (em_nan_guard && isnan(v)) v = 0.0;
        sw[k] = v;
      }
      __syncthreads();
      // (2) one thread per row: sequential sum in label order
      for (uint32_t r = d.x + threadIdx.x; r < d.y; r += THREADS) {
        const uint32_t b = off[r], e = off[r + 1];
        if (e - b > (uint32_t)LMAX) continue;  // block path
        double acc = 0.0;
        for (uint32_t k = b - e0a; k < e - e0a; ++k) acc += sw[k];
        row_epilogue<PHASE>(A, r, acc, logNorm, bias, pa);
      }
      __syncthreads();  // stage free for the next bulk load
    } else {
      // tile without short rows: only (possibly) empty span
    }
  }
  // long rows: whole block per row, fixed-order tree reduction
  for (uint32_t li = bid; li < nlong; li += nblk) {
    const uint32_t r = longs[li];
    const uint32_t b = off[r], e = off[r + 1];
    double acc = 0.0;
    for (uint32_t k = b + threadIdx.x; k < e; k += THREADS) {
      double v = gsrc[idx[k]] * w[k];
      if (em_nan_guard && isnan(v)) v = 0.0;
      acc += v;
    }
    acc = block_reduce<false>(acc, B.scratch);
    if (threadIdx.x == 0) row_epilogue<PHASE>(A, r, acc, logNorm, bias, pa);
    __syncthreads();
  }
}

// alphaSum of the iteration input, from the per-block partials of the previous P2
__device__ __forceinline__ double sum_partials(const double* part, uint32_t n, double extra,
                                               double* scratch) {
  double acc = 0.0;
  for (uint32_t i = threadIdx.x; i < n; i += THREADS) acc += part[i];
  acc = block_reduce<false>(acc, scratch);
  return acc + extra;
}

__global__ void __launch_bounds__(THREADS)
k_em_persistent(const __grid_constant__ EmArgs A) {
  extern __shared__ __align__(128) unsigned char smem[];
  cg::grid_group grid = cg::this_grid();
  Block B;
  block_setup(B, smem);
  const uint32_t bid = blockIdx.x, nblk = gridDim.x;

  uint32_t it = 0;
  bool converged = false;
  double logNorm = A.vbem ? digamma_pos(A.sum0) : 0.0;
  while (it < A.min_iter || (it < A.max_iter && !converged)) {
    const uint32_t par = it & 1u;
    if (bid == 0 && threadIdx.x == 0) A.maxrel[par] = 0ull;
    P2Acc pa{0.0, 0.0};
    run_phase<1>(A, B, bid, nblk, 0.0, 0.0, pa);
    grid.sync();
    if (A.vbem && it > 0) {
      // lagged logNorm: alphaSum of THIS iteration's input = partials written by the
      // previous P2.  Any common factor in theta cancels in P1/P2.
      double s = sum_partials(A.sum_partial + (size_t)(par ^ 1u) * nblk, nblk, A.inactive_sum,
                              B.scratch);
      logNorm = digamma_pos(s);
    }
    const double bias = (!A.vbem && it == 0) ? 1.0 : 0.0;  // alphasPrime starts at 1.0 (:812,:821)
    run_phase<2>(A, B, bid, nblk, logNorm, bias, pa);
    double bs = block_reduce<false>(pa.sum, B.scratch);
    double bm = block_reduce<true>(pa.maxrel, B.scratch);
    if (threadIdx.x == 0) {
      A.sum_partial[(size_t)par * nblk + bid] = bs;
      if (bm > 0.0) atomicMax(&A.maxrel[par], (unsigned long long)__double_as_longlong(bm));
    }
    grid.sync();
    const double mr = __longlong_as_double((long long)A.maxrel[par]);
    converged = !(mr > A.tol);
    ++it;
  }
  if (bid == 0 && threadIdx.x == 0) {
    A.out[0] = it;
    A.out[1] = converged ? 1u : 0u;
    A.out[2] = (it - 1) & 1u;  // parity slot holding the last maxrel
  }
}

// multi-kernel variant ------------------------------------------------------
__global__ void __launch_bounds__(THREADS) k_em_p1(const __grid_constant__ EmArgs A) {
  extern __shared__ __align__(128) unsigned char smem[];
  Block B;
  block_setup(B, smem);
  P2Acc pa{0.0, 0.0};
  run_phase<1>(A, B, blockIdx.x, gridDim.x, 0.0, 0.0, pa);
}
// it_par: parity of this iteration; first: iteration 0 (exact sum0)
__global__ void __launch_bounds__(THREADS)
k_em_p2(const __grid_constant__ EmArgs A, uint32_t it, const uint32_t* __restrict__ done_flag) {
  extern __shared__ __align__(128) unsigned char smem[];
  if (done_flag && *done_flag) return;
  Block B;
  block_setup(B, smem);
  const uint32_t par = it & 1u;
  double logNorm = 0.0;
  if (A.vbem) {
    if (it == 0) logNorm = digamma_pos(A.sum0);
    else
      logNorm = digamma_pos(sum_partials(A.sum_partial + (size_t)(par ^ 1u) * gridDim.x, gridDim.x,
                                         A.inactive_sum, B.scratch));
  }
  const double bias = (!A.vbem && it == 0) ? 1.0 : 0.0;
  P2Acc pa{0.0, 0.0};
  run_phase<2>(A, B, blockIdx.x, gridDim.x, logNorm, bias, pa);
  double bs = block_reduce<false>(pa.sum, B.scratch);
  double bm = block_reduce<true>(pa.maxrel, B.scratch);
  if (threadIdx.x == 0) {
    A.sum_partial[(size_t)par * gridDim.x + blockIdx.x] = bs;
    if (bm > 0.0) atomicMax(&A.maxrel[par], (unsigned long long)__double_as_longlong(bm));
  }
}
__global__ void k_reset_maxrel(unsigned long long* maxrel, uint32_t par) { maxrel[par] = 0ull; }

// multi-GPU: P2 without the per-transcript update (partial alpha' for the all-reduce)
__global__ void __launch_bounds__(THREADS) k_em_p2_partial(const __grid_constant__ EmArgs A) {
  extern __shared__ __align__(128) unsigned char smem[];
  Block B;
  block_setup(B, smem);
  P2Acc pa{0.0, 0.0};
  run_phase<3>(A, B, blockIdx.x, gridDim.x, 0.0, 0.0, pa);
}
// multi-GPU: per-transcript update over ALL transcripts from the all-reduced alpha'.
// Every rank computes the same values, so every rank takes the same decisions.
__global__ void __launch_bounds__(THREADS)
k_em_update(const __grid_constant__ EmArgs A, const double* __restrict__ red, uint32_t M,
            uint32_t it) {
  __shared__ double scratch[32];
  const uint32_t par = it & 1u;
  double logNorm = 0.0;
  if (A.vbem) {
    if (it == 0) logNorm = digamma_pos(A.sum0);
    else logNorm = digamma_pos(sum_partials(A.sum_partial + (size_t)(par ^ 1u) * gridDim.x,
                                            gridDim.x, 0.0, scratch));
  }
  const double bias = (!A.vbem && it == 0) ? 1.0 : 0.0;
  double sum = 0.0, mx = 0.0;
  for (uint32_t t = blockIdx.x * THREADS + threadIdx.x; t < M; t += gridDim.x * THREADS) {
    const double na = red[t] + bias;
    const double old = A.alpha[t];
    if (na > ALPHA_CHECK_CUTOFF) mx = fmax(mx, fabs(old - na) / na);
    A.alpha[t] = na;
    const double ap = na + A.prior[t];
    sum += ap;
    A.theta[t] = A.vbem ? ((ap > DIGAMMA_MIN) ? exp(digamma_pos(ap) - logNorm) : 0.0) : na;
  }
  double bs = block_reduce<false>(sum, scratch);
  double bm = block_reduce<true>(mx, scratch);
  if (threadIdx.x == 0) {
    A.sum_partial[(size_t)par * gridDim.x + blockIdx.x] = bs;
    if (bm > 0.0) atomicMax(&A.maxrel[par], (unsigned long long)__double_as_longlong(bm));
  }
}

}  // namespace sb

// ===========================================================================
// host side
// ===========================================================================
using namespace sb;

namespace sb {
static thread_local char g_err[512] = "";
void set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}
}  // namespace sb

extern "C" const char* sb_last_error(void) { return sb::g_err; }
extern "C" int sb_version(void) { return 100; }
extern "C" int sb_device_count(void) {
  int n = 0;
  if (cudaGetDeviceCount(&n) != cudaSuccess) {
    cudaGetLastError();
    return 0;
  }
  return n;
}

extern "C" void sb_em_default_params(sb_em_params* p) {
  memset(p, 0, sizeof(*p));
  p->use_vbem = 1;        // SalmonDefaults.hpp useVBOpt
  p->per_txp_prior = 1;   // perTranscriptPrior
  p->vb_prior = 1e-2;
  p->tol = 0.01;
  p->num_required_frags = 5e7;
  p->min_iter = 100;
  p->max_iter = 10000;
}

template <typename T>
static int dev_alloc(T** p, size_t n) {
  if (*p) { cudaFree(*p); *p = nullptr; }
  if (n == 0) n = 1;
  cudaError_t e = cudaMalloc((void**)p, n * sizeof(T));
  if (e != cudaSuccess) {
    set_error("cudaMalloc(%zu bytes) failed: %s", n * sizeof(T), cudaGetErrorString(e));
    return SB_ERR_NOMEM;
  }
  return SB_OK;
}
#define SB_TRY(x) do { int _r = (x); if (_r != SB_OK) return _r; } while (0)

extern "C" sb_em_ctx* sb_em_create(int device) {
  int n = sb_device_count();
  if (n <= 0) {
    set_error("no CUDA device available (libsalmon_b200 has no CPU fallback)");
    return nullptr;
  }
  if (device < 0 || device >= n) {
    set_error("device %d out of range (0..%d)", device, n - 1);
    return nullptr;
  }
  sb_em_ctx* c = new sb_em_ctx();
  c->device = device;
  if (cudaSetDevice(device) != cudaSuccess ||
      cudaStreamCreateWithFlags(&c->stream, cudaStreamNonBlocking) != cudaSuccess) {
    set_error("cannot initialise device %d: %s", device, cudaGetErrorString(cudaGetLastError()));
    delete c;
    return nullptr;
  }
  cudaDeviceProp prop;
  cudaGetDeviceProperties(&prop, device);
  c->n_sm = prop.multiProcessorCount;
  c->l2_bytes = (size_t)prop.l2CacheSize;
  for (int i = 0; i < 4; ++i) cudaEventCreate(&c->ev[i]);
  return c;
}

static void free_all(sb_em_ctx* c) {
  void** ptrs[] = {(void**)&c->d_off, (void**)&c->d_tids, (void**)&c->d_aux, (void**)&c->d_counts,
                   (void**)&c->d_projected, (void**)&c->d_eff_in, (void**)&c->d_unique,
                   (void**)&c->d_efflens, (void**)&c->d_prior, (void**)&c->d_alpha0,
                   (void**)&c->d_alpha, (void**)&c->d_theta, (void**)&c->d_base, (void**)&c->d_cw,
                   (void**)&c->d_packed, (void**)&c->d_packed_scan, (void**)&c->d_valid,
                   (void**)&c->d_scalars, (void**)&c->d_tcnt, (void**)&c->d_tid_row,
                   (void**)&c->cm.off, (void**)&c->cm.idx, (void**)&c->cm.w, (void**)&c->cm.tiles,
                   (void**)&c->cm.longs, (void**)&c->tm.off, (void**)&c->tm.idx, (void**)&c->tm.w,
                   (void**)&c->tm.tiles, (void**)&c->tm.longs, (void**)&c->d_cnt, (void**)&c->d_scale,
                   (void**)&c->d_ent_cls, (void**)&c->d_row_tid, (void**)&c->d_sort_keys,
                   (void**)&c->d_sort_vals, (void**)&c->d_sort_keys2, (void**)&c->d_sort_vals2,
                   (void**)&c->d_tmp, (void**)&c->d_sum_partial, (void**)&c->d_flush,
                   (void**)&c->d_part, (void**)&c->d_part_red};
  for (void** p : ptrs) {
    if (*p) cudaFree(*p);
    *p = nullptr;
  }
}

extern "C" void sb_em_destroy(sb_em_ctx* c) {
  if (!c) return;
  cudaSetDevice(c->device);
  sb_em_comm_destroy(c);
  free_all(c);
  for (int i = 0; i < 4; ++i) cudaEventDestroy(c->ev[i]);
  cudaStreamDestroy(c->stream);
  delete c;
}

extern "C" int sb_em_set_option(sb_em_ctx* c, const char* key, int64_t value) {
  if (!c || !key) { set_error("null argument"); return SB_ERR_INVALID; }
  if (!strcmp(key, "variant")) c->variant = (int)value;
  else if (!strcmp(key, "blocks_per_sm")) c->blocks_per_sm = (int)value;
  else if (!strcmp(key, "check_every")) c->check_every = (int)std::max<int64_t>(1, value);
  else { set_error("unknown option '%s'", key); return SB_ERR_INVALID; }
  return SB_OK;
}

extern "C" int sb_em_upload(sb_em_ctx* c, const sb_eq_csr* eq, const double* projected,
                            const double* eff_len, const uint64_t* unique) {
  if (!c || !eq || !projected || !eff_len || !unique) { set_error("null argument"); return SB_ERR_INVALID; }
  if (eq->n_classes && (!eq->off || !eq->counts)) { set_error("null CSR arrays"); return SB_ERR_INVALID; }
  SB_CUDA(cudaSetDevice(c->device));
  const uint64_t C = eq->n_classes;
  const uint32_t M = eq->n_txps;
  const uint64_t nnz = C ? eq->off[C] : 0;
  if (M == 0) { set_error("no transcripts"); return SB_ERR_INVALID; }
  if (nnz >= 0xfffffff0ull || C >= 0xfffffff0ull) {
    set_error("eq-class table too large for 32-bit device offsets (nnz=%llu)", (unsigned long long)nnz);
    return SB_ERR_INVALID;
  }
  if (nnz && (!eq->tids || !eq->weights)) { set_error("null CSR arrays"); return SB_ERR_INVALID; }
  c->C = C; c->M = M; c->nnz = nnz;
  c->prepared = false;
  SB_TRY(dev_alloc(&c->d_off, C + 1));
  SB_TRY(dev_alloc(&c->d_tids, nnz));
  SB_TRY(dev_alloc(&c->d_aux, nnz));
  SB_TRY(dev_alloc(&c->d_counts, C));
  SB_TRY(dev_alloc(&c->d_projected, M));
  SB_TRY(dev_alloc(&c->d_eff_in, M));
  SB_TRY(dev_alloc(&c->d_unique, M));
  cudaStream_t st = c->stream;
  if (C) {
    SB_CUDA(cudaMemcpyAsync(c->d_off, eq->off, (C + 1) * 8, cudaMemcpyHostToDevice, st));
    SB_CUDA(cudaMemcpyAsync(c->d_counts, eq->counts, C * 8, cudaMemcpyHostToDevice, st));
  } else {
    uint64_t z = 0;
    SB_CUDA(cudaMemcpyAsync(c->d_off, &z, 8, cudaMemcpyHostToDevice, st));
  }
  if (nnz) {
    SB_CUDA(cudaMemcpyAsync(c->d_tids, eq->tids, nnz * 4, cudaMemcpyHostToDevice, st));
    SB_CUDA(cudaMemcpyAsync(c->d_aux, eq->weights, nnz * 8, cudaMemcpyHostToDevice, st));
  }
  SB_CUDA(cudaMemcpyAsync(c->d_projected, projected, (size_t)M * 8, cudaMemcpyHostToDevice, st));
  SB_CUDA(cudaMemcpyAsync(c->d_eff_in, eff_len, (size_t)M * 8, cudaMemcpyHostToDevice, st));
  SB_CUDA(cudaMemcpyAsync(c->d_unique, unique, (size_t)M * 8, cudaMemcpyHostToDevice, st));
  // serial host sum, same order as the reference (:778-781)
  double tw = 0.0;
  for (uint32_t i = 0; i < M; ++i) tw += projected[i];
  c->total_weight = tw;
  c->h2d_bytes = (C + 1) * 8 + C * 8 + nnz * 12 + (size_t)M * 24;
  SB_CUDA(cudaStreamSynchronize(st));
  c->uploaded = true;
  return SB_OK;
}

static inline unsigned nblk(uint64_t n, unsigned t) { return (unsigned)((n + t - 1) / t); }

static int build_tiles(sb_em_ctx* c, SegMat& m, uint32_t* d_nlong_slot) {
  cudaStream_t st = c->stream;
  m.n_tiles = (m.nnz + TILE - 1) / TILE;
  SB_TRY(dev_alloc(&m.tiles, (size_t)m.n_tiles));
  if (m.n_tiles)
    k_tiles<<<nblk(m.n_tiles, 128), 128, 0, st>>>(m.n_rows, m.off, m.nnz, m.n_tiles, m.tiles);
  // long rows
  uint32_t cap = m.nnz / (LMAX + 1) + 1;
  SB_TRY(dev_alloc(&m.longs, (size_t)cap));
  SB_CUDA(cudaMemsetAsync(d_nlong_slot, 0, 4, st));
  if (m.n_rows)
    k_long_rows<<<nblk(m.n_rows, 256), 256, 0, st>>>(m.n_rows, m.off, m.longs, d_nlong_slot, cap);
  c->launches += 2;
  SB_CUDA(cudaMemcpyAsync(&m.n_long, d_nlong_slot, 4, cudaMemcpyDeviceToHost, st));
  SB_CUDA(cudaStreamSynchronize(st));
  if (m.n_long > 1) {
    // deterministic order of the list itself is irrelevant to results (each long row is
    // reduced independently with a fixed tree) but keep it sorted for reproducible timing.
    std::vector<uint32_t> h(m.n_long);
    SB_CUDA(cudaMemcpy(h.data(), m.longs, (size_t)m.n_long * 4, cudaMemcpyDeviceToHost));
    std::sort(h.begin(), h.end());
    SB_CUDA(cudaMemcpy(m.longs, h.data(), (size_t)m.n_long * 4, cudaMemcpyHostToDevice));
  }
  return SB_OK;
}

extern "C" int sb_em_prepare(sb_em_ctx* c, const sb_em_params* p, sb_em_stats* stats) {
  if (!c || !p) { set_error("null argument"); return SB_ERR_INVALID; }
  if (!c->uploaded) { set_error("sb_em_prepare before sb_em_upload"); return SB_ERR_STATE; }
  SB_CUDA(cudaSetDevice(c->device));
  cudaStream_t st = c->stream;
  c->params = *p;
  c->launches = 0;
  const uint64_t C = c->C;
  const uint32_t M = c->M;
  const uint64_t nnz = c->nnz;
  SB_CUDA(cudaEventRecord(c->ev[0], st));

  SB_TRY(dev_alloc(&c->d_efflens, M));
  SB_TRY(dev_alloc(&c->d_prior, M));
  SB_TRY(dev_alloc(&c->d_alpha0, M));
  SB_TRY(dev_alloc(&c->d_alpha, M));
  SB_TRY(dev_alloc(&c->d_theta, M));
  SB_TRY(dev_alloc(&c->d_base, M));
  SB_TRY(dev_alloc(&c->d_cw, nnz));
  SB_TRY(dev_alloc(&c->d_packed, std::max<uint64_t>(C, M) + 1));
  SB_TRY(dev_alloc(&c->d_packed_scan, std::max<uint64_t>(C, M) + 1));
  SB_TRY(dev_alloc(&c->d_valid, C));
  SB_TRY(dev_alloc(&c->d_scalars, 64));
  SB_TRY(dev_alloc(&c->d_tcnt, M));
  SB_TRY(dev_alloc(&c->d_tid_row, M));
  SB_CUDA(cudaMemsetAsync(c->d_base, 0, (size_t)M * 8, st));
  SB_CUDA(cudaMemsetAsync(c->d_scalars, 0, 64 * 8, st));
  SB_CUDA(cudaMemsetAsync(c->d_tcnt, 0, (size_t)M * 4, st));
  SB_CUDA(cudaMemsetAsync(c->d_packed + std::max<uint64_t>(C, M), 0, 8, st));

  k_txp_init<<<nblk(M, 256), 256, 0, st>>>(M, c->d_projected, c->d_eff_in, c->d_unique, *p,
                                            c->total_weight, c->d_efflens, c->d_prior, c->d_alpha0);
  c->launches++;
  unsigned long long* d_ndeg = (unsigned long long*)(c->d_scalars + 0);
  if (C) {
    k_class_combine<<<nblk(C, 128), 128, 0, st>>>(C, c->d_off, c->d_tids, c->d_aux, c->d_counts,
                                                   c->d_efflens, c->d_alpha0, *p, c->d_cw,
                                                   c->d_packed, c->d_base, c->d_valid, d_ndeg);
    c->launches++;
  }
  // scan (class count, entry count) packed into one u64
  size_t tmp_bytes = 0;
  cub::DeviceScan::ExclusiveSum(nullptr, tmp_bytes, c->d_packed, c->d_packed_scan, (int)(C + 1), st);
  size_t sort_tmp = 0;
  {
    uint32_t* k = nullptr; uint32_t* v = nullptr;
    cub::DeviceRadixSort::SortPairs(nullptr, sort_tmp, k, k, v, v, (int)std::max<uint64_t>(nnz, 1), 0, 32, st);
  }
  size_t scan_m = 0;
  cub::DeviceScan::ExclusiveSum(nullptr, scan_m, c->d_packed, c->d_packed_scan, (int)(M + 1), st);
  size_t need = std::max(std::max(tmp_bytes, sort_tmp), scan_m);
  if (need > c->tmp_bytes) {
    SB_TRY(dev_alloc((unsigned char**)&c->d_tmp, need));
    c->tmp_bytes = need;
  }
  SB_CUDA(cub::DeviceScan::ExclusiveSum(c->d_tmp, need, c->d_packed, c->d_packed_scan, (int)(C + 1), st));
  c->launches += 2;
  uint64_t tot = 0;
  SB_CUDA(cudaMemcpyAsync(&tot, c->d_packed_scan + C, 8, cudaMemcpyDeviceToHost, st));
  SB_CUDA(cudaMemcpyAsync(&c->n_degenerate, d_ndeg, 8, cudaMemcpyDeviceToHost, st));
  SB_CUDA(cudaStreamSynchronize(st));
  const uint32_t Cm = (uint32_t)(tot >> 32);
  const uint32_t nnzm = (uint32_t)(tot & 0xffffffffu);
  c->cm.n_rows = Cm; c->cm.nnz = nnzm;

  SB_TRY(dev_alloc(&c->cm.off, (size_t)Cm + 1));
  SB_TRY(dev_alloc(&c->cm.idx, (size_t)nnzm + 16));
  SB_TRY(dev_alloc(&c->cm.w, (size_t)nnzm + 16));
  SB_TRY(dev_alloc(&c->d_cnt, (size_t)Cm));
  SB_TRY(dev_alloc(&c->d_scale, (size_t)Cm));
  SB_TRY(dev_alloc(&c->d_ent_cls, (size_t)nnzm));
  SB_CUDA(cudaMemsetAsync(c->cm.idx, 0, ((size_t)nnzm + 16) * 4, st));
  SB_CUDA(cudaMemsetAsync(c->cm.w, 0, ((size_t)nnzm + 16) * 8, st));
  SB_CUDA(cudaMemcpyAsync(c->cm.off + Cm, &nnzm, 4, cudaMemcpyHostToDevice, st));
  if (C) {
    k_compact<<<nblk(C, 128), 128, 0, st>>>(C, c->d_off, c->d_tids, c->d_cw, c->d_counts,
                                             c->d_packed, c->d_packed_scan, c->cm.off, c->cm.idx,
                                             c->cm.w, c->d_cnt, c->d_ent_cls, c->d_tcnt);
    c->launches++;
  }
  // transcript-major copy: stable radix sort of (tid, entry) pairs
  SB_TRY(dev_alloc(&c->d_sort_keys, (size_t)nnzm));
  SB_TRY(dev_alloc(&c->d_sort_vals, (size_t)nnzm));
  SB_TRY(dev_alloc(&c->d_sort_keys2, (size_t)nnzm));
  SB_TRY(dev_alloc(&c->d_sort_vals2, (size_t)nnzm));
  if (nnzm) {
    SB_CUDA(cudaMemcpyAsync(c->d_sort_keys, c->cm.idx, (size_t)nnzm * 4, cudaMemcpyDeviceToDevice, st));
    k_iota<<<nblk(nnzm, 256), 256, 0, st>>>(nnzm, c->d_sort_vals);
    int end_bit = 1;
    while (end_bit < 32 && (1ull << end_bit) < (uint64_t)M) ++end_bit;
    size_t tb = c->tmp_bytes;
    SB_CUDA(cub::DeviceRadixSort::SortPairs(c->d_tmp, tb, c->d_sort_keys, c->d_sort_keys2,
                                            c->d_sort_vals, c->d_sort_vals2, (int)nnzm, 0, end_bit, st));
    c->launches += 2 + (end_bit + 7) / 8 * 3;
  }
  k_row_pack<<<nblk(M, 256), 256, 0, st>>>(M, c->d_tcnt, c->d_packed);
  SB_CUDA(cudaMemsetAsync(c->d_packed + M, 0, 8, st));
  {
    size_t tb = c->tmp_bytes;
    SB_CUDA(cub::DeviceScan::ExclusiveSum(c->d_tmp, tb, c->d_packed, c->d_packed_scan, (int)(M + 1), st));
  }
  c->launches += 3;
  uint64_t tot2 = 0;
  SB_CUDA(cudaMemcpyAsync(&tot2, c->d_packed_scan + M, 8, cudaMemcpyDeviceToHost, st));
  SB_CUDA(cudaStreamSynchronize(st));
  const uint32_t R = (uint32_t)(tot2 >> 32);
  if ((uint32_t)(tot2 & 0xffffffffu) != nnzm) {
    set_error("internal: transcript-major entry count mismatch");
    return SB_ERR_STATE;
  }
  c->tm.n_rows = R; c->tm.nnz = nnzm;
  SB_TRY(dev_alloc(&c->tm.off, (size_t)R + 1));
  SB_TRY(dev_alloc(&c->tm.idx, (size_t)nnzm + 16));
  SB_TRY(dev_alloc(&c->tm.w, (size_t)nnzm + 16));
  SB_TRY(dev_alloc(&c->d_row_tid, (size_t)R));
  SB_CUDA(cudaMemsetAsync(c->tm.idx, 0, ((size_t)nnzm + 16) * 4, st));
  SB_CUDA(cudaMemsetAsync(c->tm.w, 0, ((size_t)nnzm + 16) * 8, st));
  SB_CUDA(cudaMemcpyAsync(c->tm.off + R, &nnzm, 4, cudaMemcpyHostToDevice, st));
  k_row_fill<<<nblk(M, 256), 256, 0, st>>>(M, c->d_tcnt, c->d_packed_scan, c->tm.off, c->d_row_tid,
                                            c->d_tid_row);
  if (nnzm)
    k_gather_csc<<<nblk(nnzm, 256), 256, 0, st>>>(nnzm, c->d_sort_vals2, c->d_ent_cls, c->cm.w,
                                                   c->tm.idx, c->tm.w);
  c->launches += 2;
  uint32_t* d_nlong = (uint32_t*)(c->d_scalars + 8);
  SB_TRY(build_tiles(c, c->cm, d_nlong));
  SB_TRY(build_tiles(c, c->tm, d_nlong + 1));

  // iteration-0 state
  double* d_sum0 = c->d_scalars + 16;
  double* d_inact = c->d_scalars + 17;
  k_sum1<<<1, 1024, 0, st>>>(M, c->d_alpha0, c->d_prior, c->d_tid_row, 0, d_sum0);
  k_sum1<<<1, 1024, 0, st>>>(M, c->d_base, c->d_prior, c->d_tid_row, 1, d_inact);
  c->launches += 2;
  SB_CUDA(cudaMemcpyAsync(&c->sum0, d_sum0, 8, cudaMemcpyDeviceToHost, st));
  SB_CUDA(cudaMemcpyAsync(&c->inactive_sum, d_inact, 8, cudaMemcpyDeviceToHost, st));

  // launch geometry
  SB_CUDA(cudaFuncSetAttribute(k_em_persistent, cudaFuncAttributeMaxDynamicSharedMemorySize,
                               (int)SmemLayout::total));
  SB_CUDA(cudaFuncSetAttribute(k_em_p1, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)SmemLayout::total));
  SB_CUDA(cudaFuncSetAttribute(k_em_p2, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)SmemLayout::total));
  int occ = 0;
  SB_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, k_em_persistent, THREADS, SmemLayout::total));
  if (occ < 1) { set_error("persistent EM kernel does not fit on an SM"); return SB_ERR_CUDA; }
  if (c->blocks_per_sm > 0) occ = std::min(occ, c->blocks_per_sm);
  c->grid = (uint32_t)(occ * c->n_sm);
  SB_TRY(dev_alloc(&c->d_sum_partial, (size_t)2 * c->grid));
  SB_CUDA(cudaStreamSynchronize(st));
  SB_CUDA(cudaEventRecord(c->ev[1], st));
  SB_CUDA(cudaEventSynchronize(c->ev[1]));
  float ms = 0;
  cudaEventElapsedTime(&ms, c->ev[0], c->ev[1]);
  c->prepare_ms = ms;
  c->prepared = true;
  if (stats) {
    memset(stats, 0, sizeof(*stats));
    stats->n_degenerate = c->n_degenerate;
    stats->n_multi_classes = Cm;
    stats->nnz_multi = nnzm;
    stats->n_active_txps = R;
    stats->prepare_ms = ms;
    stats->gpu_launches = c->launches;
  }
  return SB_OK;
}

static void fill_args(sb_em_ctx* c, EmArgs& A) {
  memset(&A, 0, sizeof(A));
  A.c_off = c->cm.off; A.c_idx = c->cm.idx; A.c_w = c->cm.w; A.c_tiles = c->cm.tiles;
  A.c_long = c->cm.longs; A.c_cnt = c->d_cnt; A.scale = c->d_scale;
  A.c_ntiles = c->cm.n_tiles; A.c_nlong = c->cm.n_long; A.c_nrows = c->cm.n_rows;
  A.t_off = c->tm.off; A.t_idx = c->tm.idx; A.t_w = c->tm.w; A.t_tiles = c->tm.tiles;
  A.t_long = c->tm.longs; A.row_tid = c->d_row_tid;
  A.t_ntiles = c->tm.n_tiles; A.t_nlong = c->tm.n_long; A.t_nrows = c->tm.n_rows;
  A.alpha = c->d_alpha; A.theta = c->d_theta; A.prior = c->d_prior; A.base = c->d_base;
  A.sum_partial = c->d_sum_partial;
  A.maxrel = (unsigned long long*)(c->d_scalars + 24);
  A.inactive_sum = c->inactive_sum; A.sum0 = c->sum0;
  A.tol = c->params.tol; A.min_iter = c->params.min_iter; A.max_iter = c->params.max_iter;
  A.vbem = c->params.use_vbem;
  A.out = (uint32_t*)(c->d_scalars + 32);
}

// ---------------------------------------------------------------------------
// multi-GPU: NCCL is loaded at run time (dlopen) so the library has no link-time
// dependency on it; torch's bundled libnccl.so.2 is picked up when present.
// ---------------------------------------------------------------------------
#include <dlfcn.h>
namespace {
struct NcclUid { char b[128]; };  // ncclUniqueId (NCCL_UNIQUE_ID_BYTES = 128), passed by value
struct NcclFns {
  void* h = nullptr;
  int (*GetUniqueId)(void*) = nullptr;
  int (*CommInitRank)(void**, int, NcclUid, int) = nullptr;
  int (*AllReduce)(const void*, void*, size_t, int, int, void*, cudaStream_t) = nullptr;
  int (*CommDestroy)(void*) = nullptr;
  const char* (*GetErrorString)(int) = nullptr;
};
NcclFns g_nccl;
int nccl_load() {
  if (g_nccl.h) return SB_OK;
  const char* names[] = {"libnccl.so.2", "libnccl.so", nullptr};
  void* h = nullptr;
  for (int i = 0; names[i] && !h; ++i) h = dlopen(names[i], RTLD_NOW | RTLD_GLOBAL);
  if (!h) { set_error("cannot dlopen libnccl.so.2: %s", dlerror()); return SB_ERR_NCCL; }
  g_nccl.GetUniqueId = (decltype(g_nccl.GetUniqueId))dlsym(h, "ncclGetUniqueId");
  g_nccl.CommInitRank = (decltype(g_nccl.CommInitRank))dlsym(h, "ncclCommInitRank");
  g_nccl.AllReduce = (decltype(g_nccl.AllReduce))dlsym(h, "ncclAllReduce");
  g_nccl.CommDestroy = (decltype(g_nccl.CommDestroy))dlsym(h, "ncclCommDestroy");
  g_nccl.GetErrorString = (decltype(g_nccl.GetErrorString))dlsym(h, "ncclGetErrorString");
  if (!g_nccl.GetUniqueId || !g_nccl.CommInitRank || !g_nccl.AllReduce || !g_nccl.CommDestroy) {
    set_error("libnccl is missing required symbols");
    return SB_ERR_NCCL;
  }
  g_nccl.h = h;
  return SB_OK;
}
#define SB_NCCL(call)                                                            \
  do {                                                                           \
    int _r = (call);                                                             \
    if (_r != 0) {                                                               \
      set_error("%s:%d: %s -> nccl error %d (%s)", __FILE__, __LINE__, #call, _r, \
                g_nccl.GetErrorString ? g_nccl.GetErrorString(_r) : "?");        \
      return SB_ERR_NCCL;                                                        \
    }                                                                            \
  } while (0)
}  // namespace

extern "C" int sb_nccl_unique_id(void* out128) {
  if (!out128) { set_error("null argument"); return SB_ERR_INVALID; }
  SB_TRY(nccl_load());
  SB_NCCL(g_nccl.GetUniqueId(out128));
  return SB_OK;
}
extern "C" int sb_em_comm_init(sb_em_ctx* c, int rank, int nranks, const void* uid) {
  if (!c || !uid || nranks < 1 || rank < 0 || rank >= nranks) { set_error("bad argument"); return SB_ERR_INVALID; }
  SB_TRY(nccl_load());
  SB_CUDA(cudaSetDevice(c->device));
  NcclUid u;
  memcpy(u.b, uid, 128);
  void* comm = nullptr;
  SB_NCCL(g_nccl.CommInitRank(&comm, nranks, u, rank));
  c->nccl_comm = comm;
  c->rank = rank;
  c->nranks = nranks;
  return SB_OK;
}
extern "C" int sb_em_comm_destroy(sb_em_ctx* c) {
  if (!c) return SB_OK;
  if (c->nccl_comm && g_nccl.CommDestroy) {
    g_nccl.CommDestroy(c->nccl_comm);
    c->nccl_comm = nullptr;
  }
  c->nranks = 1;
  c->rank = 0;
  return SB_OK;
}

// One iteration = P1, P2-partial, all-reduce(alpha'), update.  Classes stay sharded.
static int em_run_multi_gpu(sb_em_ctx* c, EmArgs& A, uint32_t* out, uint32_t* launches,
                            uint32_t* loop_launches, float* loop_ms) {
  cudaStream_t st = c->stream;
  const uint32_t M = c->M;
  SB_TRY(dev_alloc(&c->d_part, (size_t)M));
  SB_TRY(dev_alloc(&c->d_part_red, (size_t)M));
  // locally inactive transcripts contribute their (constant) folded singleton mass
  SB_CUDA(cudaMemcpyAsync(c->d_part, c->d_base, (size_t)M * 8, cudaMemcpyDeviceToDevice, st));
  A.part_out = c->d_part;
  const uint32_t ugrid = (uint32_t)std::min<uint32_t>(c->grid, (M + THREADS - 1) / THREADS);
  // the update kernel writes its own partials: re-size for its grid
  uint32_t it = 0;
  bool converged = false;
  SB_CUDA(cudaEventRecord(c->ev[2], st));
  while (it < A.min_iter || (it < A.max_iter && !converged)) {
    k_reset_maxrel<<<1, 1, 0, st>>>(A.maxrel, it & 1u);
    k_em_p1<<<c->grid, THREADS, SmemLayout::total, st>>>(A);
    k_em_p2_partial<<<c->grid, THREADS, SmemLayout::total, st>>>(A);
    SB_NCCL(g_nccl.AllReduce(c->d_part, c->d_part_red, (size_t)M, /*ncclFloat64*/ 8, /*ncclSum*/ 0,
                             c->nccl_comm, st));
    k_em_update<<<ugrid, THREADS, 0, st>>>(A, c->d_part_red, M, it);
    *launches += 5; *loop_launches += 4;
    ++it;
    if (it >= A.min_iter) {
      unsigned long long mr = 0;
      SB_CUDA(cudaMemcpyAsync(&mr, A.maxrel + ((it - 1) & 1u), 8, cudaMemcpyDeviceToHost, st));
      SB_CUDA(cudaStreamSynchronize(st));
      double d;
      memcpy(&d, &mr, 8);
      converged = !(d > A.tol);
    }
  }
  SB_CUDA(cudaEventRecord(c->ev[3], st));
  SB_CUDA(cudaStreamSynchronize(st));
  cudaEventElapsedTime(loop_ms, c->ev[2], c->ev[3]);
  out[0] = it; out[1] = converged; out[2] = (it - 1) & 1u;
  return SB_OK;
}

extern "C" int sb_em_run(sb_em_ctx* c, sb_em_stats* stats) {
  if (!c) { set_error("null argument"); return SB_ERR_INVALID; }
  if (!c->prepared) { set_error("sb_em_run before sb_em_prepare"); return SB_ERR_STATE; }
  SB_CUDA(cudaSetDevice(c->device));
  cudaStream_t st = c->stream;
  const uint32_t M = c->M;
  uint32_t launches = 0;
  SB_CUDA(cudaEventRecord(c->ev[0], st));
  // restart from the prepared state
  k_theta0<<<nblk(M, 256), 256, 0, st>>>(M, c->params.use_vbem, c->d_alpha0, c->d_prior,
                                          c->d_scalars + 16, c->d_alpha, c->d_theta);
  ++launches;
  SB_CUDA(cudaMemsetAsync(c->d_scalars + 24, 0, 16 * 8, st));
  EmArgs A;
  fill_args(c, A);
  uint32_t out[4] = {0, 0, 0, 0};
  float loop_ms = 0;
  uint32_t loop_launches = 0;
  const bool multi_gpu = c->nranks > 1;
  if (c->params.max_iter == 0 && c->params.min_iter == 0) {
    // nothing to iterate
  } else if (c->variant == 1 && !multi_gpu) {
    void* args[] = {(void*)&A};
    SB_CUDA(cudaEventRecord(c->ev[2], st));
    SB_CUDA(cudaLaunchCooperativeKernel((void*)k_em_persistent, dim3(c->grid), dim3(THREADS), args,
                                        SmemLayout::total, st));
    SB_CUDA(cudaEventRecord(c->ev[3], st));
    ++launches; ++loop_launches;
    SB_CUDA(cudaMemcpyAsync(out, A.out, 16, cudaMemcpyDeviceToHost, st));
    SB_CUDA(cudaStreamSynchronize(st));
    cudaEventElapsedTime(&loop_ms, c->ev[2], c->ev[3]);
  } else if (!multi_gpu) {
    // multi-kernel variant: host checks convergence every `check_every` iterations only
    // where the reference's loop condition can actually change (it >= min_iter).
    uint32_t it = 0;
    bool converged = false;
    SB_CUDA(cudaEventRecord(c->ev[2], st));
    while (it < A.min_iter || (it < A.max_iter && !converged)) {
      k_reset_maxrel<<<1, 1, 0, st>>>(A.maxrel, it & 1u);
      k_em_p1<<<c->grid, THREADS, SmemLayout::total, st>>>(A);
      k_em_p2<<<c->grid, THREADS, SmemLayout::total, st>>>(A, it, nullptr);
      launches += 3; loop_launches += 2;
      ++it;
      if (it >= A.min_iter) {
        unsigned long long mr = 0;
        SB_CUDA(cudaMemcpyAsync(&mr, A.maxrel + ((it - 1) & 1u), 8, cudaMemcpyDeviceToHost, st));
        SB_CUDA(cudaStreamSynchronize(st));
        double d;
        memcpy(&d, &mr, 8);
        converged = !(d > A.tol);
      }
    }
    SB_CUDA(cudaEventRecord(c->ev[3], st));
    SB_CUDA(cudaStreamSynchronize(st));
    cudaEventElapsedTime(&loop_ms, c->ev[2], c->ev[3]);
    out[0] = it; out[1] = converged; out[2] = (it - 1) & 1u;
  } else {
    int r = em_run_multi_gpu(c, A, out, &launches, &loop_launches, &loop_ms);
    if (r != SB_OK) return r;
  }
  c->iters = out[0];
  c->converged = out[1];
  if (out[0] > 0) {
    unsigned long long mr = 0;
    SB_CUDA(cudaMemcpy(&mr, A.maxrel + out[2], 8, cudaMemcpyDeviceToHost));
    memcpy(&c->max_rel_diff, &mr, 8);
  } else {
    c->max_rel_diff = -DBL_MAX;
  }
  // inactive transcripts
  if (out[0] > 0 && !multi_gpu) {
    double bias = (!c->params.use_vbem && out[0] == 1) ? 1.0 : 0.0;
    k_finalize_inactive<<<nblk(M, 256), 256, 0, st>>>(M, c->d_tid_row, c->d_base, bias, c->d_alpha);
    ++launches;
  }
  SB_CUDA(cudaEventRecord(c->ev[1], st));
  SB_CUDA(cudaEventSynchronize(c->ev[1]));
  float ms = 0;
  cudaEventElapsedTime(&ms, c->ev[0], c->ev[1]);
  c->run_ms = ms;
  c->launches += launches;
  if (stats) {
    stats->iters = c->iters;
    stats->converged = c->converged;
    stats->max_rel_diff = c->max_rel_diff;
    stats->n_degenerate = c->n_degenerate;
    stats->n_multi_classes = c->cm.n_rows;
    stats->nnz_multi = c->cm.nnz;
    stats->n_active_txps = c->tm.n_rows;
    stats->prepare_ms = c->prepare_ms;
    stats->run_ms = ms;
    stats->loop_kernel_ms = loop_ms;
    stats->loop_kernel_launches = loop_launches;
    stats->gpu_launches = launches;
  }
  return SB_OK;
}

extern "C" int sb_em_download(sb_em_ctx* c, double* alpha_out, sb_em_stats* stats) {
  if (!c || !alpha_out) { set_error("null argument"); return SB_ERR_INVALID; }
  if (!c->prepared) { set_error("sb_em_download before sb_em_prepare"); return SB_ERR_STATE; }
  SB_CUDA(cudaSetDevice(c->device));
  SB_CUDA(cudaMemcpyAsync(alpha_out, c->d_alpha, (size_t)c->M * 8, cudaMemcpyDeviceToHost, c->stream));
  SB_CUDA(cudaStreamSynchronize(c->stream));
  // truncation + alphaSum, serial in reference order (:1004-1014, EMUtils.cpp:55-67)
  double alphaSum = 0.0;
  for (uint32_t i = 0; i < c->M; ++i) {
    if (alpha_out[i] <= 1e-8) alpha_out[i] = 0.0;
    alphaSum += alpha_out[i];
  }
  if (stats) stats->alpha_sum = alphaSum;
  return (alphaSum < DBL_MIN) ? 1 : SB_OK;   // :1016-1020 -> false
}

extern "C" int sb_em_get_combined(sb_em_ctx* c, double* combined_out, uint8_t* valid_out) {
  if (!c) { set_error("null argument"); return SB_ERR_INVALID; }
  if (!c->prepared) { set_error("not prepared"); return SB_ERR_STATE; }
  SB_CUDA(cudaSetDevice(c->device));
  if (combined_out && c->nnz)
    SB_CUDA(cudaMemcpy(combined_out, c->d_cw, c->nnz * 8, cudaMemcpyDeviceToHost));
  if (valid_out && c->C) SB_CUDA(cudaMemcpy(valid_out, c->d_valid, c->C, cudaMemcpyDeviceToHost));
  return SB_OK;
}

extern "C" int sb_em_optimize(sb_em_ctx* c, const sb_eq_csr* eq, const sb_em_params* p,
                              const double* projected, const double* eff_len,
                              const uint64_t* unique, double* alpha_out, sb_em_stats* stats) {
  sb_em_stats local;
  sb_em_stats* s = stats ? stats : &local;
  SB_TRY(sb_em_upload(c, eq, projected, eff_len, unique));
  SB_TRY(sb_em_prepare(c, p, s));
  uint32_t prep_launches = s->gpu_launches;
  SB_TRY(sb_em_run(c, s));
  s->gpu_launches += prep_launches;
  return sb_em_download(c, alpha_out, s);
}

extern "C" int sb_flush_l2(sb_em_ctx* c) {
  if (!c) { set_error("null argument"); return SB_ERR_INVALID; }
  SB_CUDA(cudaSetDevice(c->device));
  size_t bytes = std::max<size_t>(c->l2_bytes * 2, (size_t)256 << 20);
  if (!c->d_flush || c->flush_bytes < bytes) {
    SB_TRY(dev_alloc((unsigned char**)&c->d_flush, bytes));
    c->flush_bytes = bytes;
  }
  SB_CUDA(cudaMemsetAsync(c->d_flush, (int)(c->flush_ctr++ & 0xff), bytes, c->stream));
  SB_CUDA(cudaStreamSynchronize(c->stream));
  return SB_OK;
}
