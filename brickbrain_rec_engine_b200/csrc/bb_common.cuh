// Shared device-side pieces of the scoring engine: candidate keys, the warp
// bitonic sorter, per-query candidate buffers with a running threshold, and the
// hard-constraint predicate.  Used by both scan kernels (tcgen05 bf16 and FFMA
// fp32) and by the merge / re-score kernel.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "../../include/brickscore.h"

namespace bb {

constexpr int kCap = 256;          // candidate slots per (catalog split, query)
constexpr int kEpl = kCap / 32;    // keys per lane in the warp sorter
constexpr int kMaxSel = 128;       // largest k' kept per buffer after a compaction
constexpr unsigned kFull = 0xffffffffu;

// ---------------------------------------------------------------- candidate keys
// 64-bit key = (order-preserving score bits << 32) | ~id : larger key == better
// (higher score, then LOWER id).  Key 0 is below every real candidate.
__host__ __device__ __forceinline__ uint32_t f2ord(float f) {
#ifdef __CUDA_ARCH__
  uint32_t u = __float_as_uint(f);
#else
  union { float f; uint32_t u; } c; c.f = f; uint32_t u = c.u;
#endif
  return (u & 0x80000000u) ? ~u : (u | 0x80000000u);
}
__host__ __device__ __forceinline__ float ord2f(uint32_t o) {
  uint32_t u = (o & 0x80000000u) ? (o & 0x7fffffffu) : ~o;
#ifdef __CUDA_ARCH__
  return __uint_as_float(u);
#else
  union { float f; uint32_t u; } c; c.u = u; return c.f;
#endif
}
constexpr uint32_t kOrdNegInf = 0x007fffffu;   // f2ord(-inf)
__device__ __forceinline__ uint64_t make_key(float score, uint32_t id) {
  return (static_cast<uint64_t>(f2ord(score)) << 32) | static_cast<uint32_t>(~id);
}
__device__ __forceinline__ float key_score(uint64_t k) { return ord2f(static_cast<uint32_t>(k >> 32)); }
__device__ __forceinline__ int32_t key_id(uint64_t k) { return static_cast<int32_t>(~static_cast<uint32_t>(k)); }

// ---------------------------------------------------------------- warp bitonic sort (descending)
// 32*EPL keys distributed as element e = i*32 + lane  (i = register index).
template <int EPL>
__device__ __forceinline__ void warp_sort_desc(uint64_t (&v)[EPL], int lane) {
#pragma unroll
  for (int k = 2; k <= 32 * EPL; k <<= 1) {
#pragma unroll
    for (int j = k >> 1; j > 0; j >>= 1) {
      if (j >= 32) {
        const int ji = j >> 5;
#pragma unroll
        for (int i = 0; i < EPL; ++i) {
          const int ip = i ^ ji;
          if (ip > i) {
            const bool desc = (((i << 5) & k) == 0);   // k >= 64 here: the bit lives in i
            const uint64_t a = v[i], b = v[ip];
            const bool sw = desc ? (a < b) : (a > b);
            v[i] = sw ? b : a;
            v[ip] = sw ? a : b;
          }
        }
      } else {
#pragma unroll
        for (int i = 0; i < EPL; ++i) {
          const int e = (i << 5) | lane;
          const uint64_t a = v[i];
          const uint64_t b = __shfl_xor_sync(kFull, a, j);
          const bool lower = (lane & j) == 0;
          const bool desc = ((e & k) == 0);
          const bool keep_max = (lower == desc);
          const uint64_t mx = a > b ? a : b, mn = a > b ? b : a;
          v[i] = keep_max ? mx : mn;
        }
      }
    }
  }
}

// ---------------------------------------------------------------- scan parameters
struct ScanParams {
  // catalog shard
  int64_t n_items;        // local rows
  int64_t id_offset;      // global id of local row 0
  const int2* meta;       // {num_parts, (uint16)year | theme_id << 16}
  // queries
  int32_t n_queries;      // B (rows beyond it are padding)
  int32_t b_pad;          // padded row count (candidate buffer stride)
  int32_t k_pad;          // operand width (multiple of 64)
  int32_t k_sel;          // k' kept at compaction (multiple of 32, <= kMaxSel)
  const bb_constraint* cons;
  int32_t cons_stride;    // 0 = one constraint shared by all rows, 1 = per row
  const int32_t* self_id; // global ids
  const uint32_t* theme_sets;
  int32_t theme_words;
  const uint32_t* allow_bitmaps;
  int64_t allow_words;
  const int64_t* excl_off;
  const int32_t* excl_ids;
  const uint32_t* shared_valid;   // validity bitmap of the single shared constraint (built by the mask kernel) or nullptr
  // candidate buffers: slot = (split * n_sub + sub) * b_pad + row; cand[slot * kCap + i], cnt[slot]
  int32_t n_sub;          // buffers per (split, row): the tcgen05 epilogue uses one per column half
  uint64_t* cand;
  int32_t* cand_cnt;
  uint32_t* thr_g;        // per-row shared running threshold (ordered bits), max over splits
  // schedule
  int32_t n_qblocks;      // query blocks
  int32_t n_splits;       // catalog splits (units = n_qblocks * n_splits)
  int32_t n_tiles;        // catalog tiles of tile_items rows
  int32_t tile_items;
  // debug
  float* dump;            // dense [b x n_items] score dump (bb_debug_scores) or nullptr
  unsigned long long* stat_candidates;
  int32_t* err_flag;      // set by watchdogs before trapping
  int32_t debug_flags;    // perf triage (env BB_DEBUG_FLAGS): 1 = epilogue never inserts, 2 = epilogue skips TMEM loads
};

__device__ __forceinline__ void split_tile_range(const ScanParams& P, int split, int& t0, int& t1) {
  t0 = static_cast<int>((static_cast<int64_t>(P.n_tiles) * split) / P.n_splits);
  t1 = static_cast<int>((static_cast<int64_t>(P.n_tiles) * (split + 1)) / P.n_splits);
}

// ---------------------------------------------------------------- hard-constraint predicate
__device__ __forceinline__ bool bit_test(const uint32_t* words, int64_t bit) {
  return (__ldg(words + (bit >> 5)) >> (bit & 31)) & 1u;
}

__device__ __forceinline__ bool list_contains(const int32_t* ids, int64_t lo, int64_t hi, int32_t gid) {
  while (lo < hi) {
    const int64_t mid = (lo + hi) >> 1;
    const int32_t v = __ldg(ids + mid);
    if (v == gid) return true;
    if (v < gid) lo = mid + 1; else hi = mid;
  }
  return false;
}

// Integer restatement of the reference's constraint SQL (hard_constraint_filter.py:318-480)
// evaluated for one (query row, local item).
__device__ __forceinline__ bool constraint_ok(const ScanParams& P, const bb_constraint& c, int64_t item) {
  if (c.flags & BB_C_ALWAYS_FALSE) return false;
  const int2 m = __ldg(P.meta + item);
  const int32_t parts = m.x;
  const int32_t year = static_cast<int16_t>(m.y & 0xffff);
  const int32_t theme = m.y >> 16;                 // arithmetic shift: -1 stays -1 (SQL NULL)
  bool ok = (parts >= c.parts_min) & (parts <= c.parts_max) & (year >= c.year_min) & (year <= c.year_max);
  if (!ok) return false;
  if (c.theme_set >= 0) {
    if (theme < 0 || theme >= P.theme_words * 32) return false;
    if (!bit_test(P.theme_sets + static_cast<int64_t>(c.theme_set) * P.theme_words, theme)) return false;
  }
  if (c.allow_bitmap >= 0) {
    if (!bit_test(P.allow_bitmaps + static_cast<int64_t>(c.allow_bitmap) * P.allow_words, item)) return false;
  }
  const int32_t gid = static_cast<int32_t>(item + P.id_offset);
#pragma unroll
  for (int e = 0; e < 3; ++e) {
    const int32_t l = c.excl_list[e];
    if (l >= 0 && list_contains(P.excl_ids, __ldg(P.excl_off + l), __ldg(P.excl_off + l + 1), gid)) return false;
  }
  return true;
}

// Validity of local `item` for query `row` (bounds, self exclusion, constraints).
static __device__ __noinline__ bool item_ok(const ScanParams& P, int row, int64_t item) {
  if (item >= P.n_items) return false;
  if (P.self_id != nullptr && __ldg(P.self_id + row) == static_cast<int32_t>(item + P.id_offset)) return false;
  if (P.cons == nullptr) return true;
  const bb_constraint c = P.cons[static_cast<int64_t>(row) * P.cons_stride];
  return constraint_ok(P, c, item);
}

// ---------------------------------------------------------------- per-row candidate buffer
// One thread owns a query row: `thr` is a lower bound of the row's k'-th best
// valid score, `cnt` the number of keys in its buffer.  A score passes when
// score >= thr; passing + valid candidates are appended; when a buffer could
// overflow the warp sorts it and keeps the best k_sel, which raises thr.
struct RowState {
  float thr;
  int cnt;
  uint64_t* buf;
};

// Warp-cooperative compaction of ONE row's buffer (all 32 lanes call with the
// same buf / cnt).  Returns the new count; *new_thr gets the k_sel-th best score
// (or -inf when fewer than k_sel candidates exist).
__device__ __forceinline__ int warp_compact(uint64_t* buf, int cnt, int k_sel, int lane, float* new_thr) {
  uint64_t v[kEpl];
#pragma unroll
  for (int i = 0; i < kEpl; ++i) {
    const int e = (i << 5) | lane;
    v[i] = (e < cnt) ? __ldcg(buf + e) : 0ull;
  }
  warp_sort_desc<kEpl>(v, lane);
  const int kw = k_sel >> 5;
  uint64_t kth = 0;
#pragma unroll
  for (int i = 0; i < kEpl; ++i) {
    if (i < kw) {
      __stcg(buf + ((i << 5) | lane), v[i]);
      if (i == kw - 1) kth = v[i];
    }
  }
  kth = __shfl_sync(kFull, kth, 31);
  *new_thr = (cnt >= k_sel) ? key_score(kth) : -INFINITY;
  __syncwarp();
  return cnt < k_sel ? cnt : k_sel;
}

// After a group of <= 32 appends per lane: compact every row of this warp whose
// buffer could overflow on the next group, and publish the raised threshold.
__device__ __forceinline__ void warp_compact_if_needed(const ScanParams& P, RowState& st, int row, int lane) {
  __syncwarp();
  unsigned need = __ballot_sync(kFull, st.cnt > kCap - 32);
  while (need) {
    const int src = __ffs(need) - 1;
    need &= need - 1;
    uint64_t* b = reinterpret_cast<uint64_t*>(__shfl_sync(kFull, reinterpret_cast<unsigned long long>(st.buf), src));
    const int c = __shfl_sync(kFull, st.cnt, src);
    float nt;
    const int nc = warp_compact(b, c, P.k_sel, lane, &nt);
    if (lane == src) {
      st.cnt = nc;
      if (nt > st.thr) {
        st.thr = nt;
        atomicMax(P.thr_g + row, f2ord(nt));
      }
    }
  }
}

__device__ __forceinline__ void row_append(const ScanParams& P, RowState& st, float score, int64_t item) {
  __stcg(st.buf + st.cnt, make_key(score, static_cast<uint32_t>(item + P.id_offset)));
  ++st.cnt;
}

}  // namespace bb
